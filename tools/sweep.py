#!/usr/bin/env python
"""Device-timed throughput sweep over every entry point and length (B200).

For each (op, N) a batch of about --mib MiB is transformed --reps times; time is taken with
CUDA events on the launch stream (shim timer), data resident in HBM and larger than L2.
Prints one row per kernel: ms, Gsamples/s, algorithmic GB/s (one read + one write of the
payload) and the fraction of the measured HBM peak.  Output also goes to --json.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cmsis-dsp_b200", "python"))
import torch  # noqa: E402
import cmsisdsp_b200 as cd  # noqa: E402


def timed(fn, reps, stream, warm=3):
    cu = cd.cuda()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t = C.c_void_p()
    cu.cmsisdsp_cuda_timer_begin(C.byref(t), stream)
    for _ in range(reps):
        fn()
    ms = C.c_float()
    cu.cmsisdsp_cuda_timer_end(t, stream, C.byref(ms))
    return ms.value / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--ops", default="cfft_f32,cfft_q31,cfft_q15,rfft_fwd,rfft_inv")
    ap.add_argument("--lens", default="16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    peak = 6536.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cd.cuda().cmsisdsp_cuda_set_device(0)
    st = torch.cuda.current_stream().cuda_stream
    rows = []
    for op in args.ops.split(","):
        for N in [int(v) for v in args.lens.split(",")]:
            if op.startswith("rfft") and N < 32:
                continue
            if op == "mfcc" and N not in (256, 512, 1024):
                continue
            nbytes = args.mib << 20
            if op.startswith("cfft") and op not in ("cfft_mag", "cfft_peak"):
                kind = op.split("_")[1]
                esz = {"f32": 8, "q31": 8, "q15": 4, "f64": 16}[kind]
                B = nbytes // (esz * N)
                cd.ensure_plans(kind, N)
                buf = torch.zeros(B * N * esz // 4, dtype=torch.int32, device=dev)
                if kind == "f32":
                    buf.view(torch.float32).normal_()
                elif kind == "f64":
                    buf.view(torch.float64).normal_()
                else:
                    buf.random_(-2**20, 2**20)
                fn = lambda: cd.cfft_device(kind, N, buf.data_ptr(), B, 0, 1, st)
                alg = 2 * B * N * esz
                samples = B * N
                info = cd.kernel_info({"f32": 0, "q31": 1, "q15": 2, "f64": 10}[kind], N)
            elif op == "mfcc":
                sys.path.insert(0, os.path.join(ROOT, "tests"))
                from oracle_lib import mfcc_config
                m = cd.Mfcc(mfcc_config(N))
                B = nbytes // (4 * N)
                a = torch.randn(B * N, device=dev)
                b = torch.empty(B, 13, device=dev)
                L = cd.lib()
                fn = lambda: L.arm_mfcc_batch_f32(C.byref(m.S), a.data_ptr(), N, b.data_ptr(), B)
                fn()                                   # creates the device plan
                alg = B * (4 * N + 4 * 13)
                samples = B * N
                info = dict(threads_per_cta=0, frames_per_cta=0, smem_bytes=0, regs_per_thread=0, ctas_per_sm=0)
            elif op in ("cfft_mag", "cfft_peak"):
                B = nbytes // (8 * N)
                cd.ensure_plans("f32", N)
                a = torch.randn(B, 2 * N, device=dev)
                b = torch.empty(B, N if op == "cfft_mag" else 2, device=dev)
                cu = cd.cuda()
                if op == "cfft_mag":
                    fn = lambda: cu.cmsisdsp_cuda_cfft_mag_f32(a.data_ptr(), b.data_ptr(), N, B, 0, 0, st)
                    alg = B * N * 12
                else:
                    fn = lambda: cu.cmsisdsp_cuda_cfft_peak_f32(a.data_ptr(), b.data_ptr(), b.data_ptr() + 4 * B, N, B, 0, st)
                    alg = B * (N * 8 + 8)
                samples = B * N
                info = cd.kernel_info(9, N)
            elif op.startswith("rfft64"):
                # arm_rfft_fast_f64 (fused kernels); algorithmic bytes as for the f32 real FFT: N in + N out doubles
                if N < 32:
                    continue
                B = nbytes // (8 * N)
                cd.ensure_rfft_f64_plans(N)
                a = torch.randn(B, N, device=dev, dtype=torch.float64)
                b = torch.empty_like(a)
                inv = int(op.endswith("inv"))
                fn = lambda: cd.rfft_f64_device(N, a.data_ptr(), b.data_ptr(), B, inv, st)
                alg = 2 * B * N * 8
                samples = B * N
                info = cd.kernel_info(12 if inv else 11, N)
            elif op.startswith("rfftq"):
                # rfftq31_fwd / rfftq31_inv / rfftq15_fwd / rfftq15_inv; N = real length; algorithmic bytes:
                # forward N in + 2N out scalars, inverse N+2 in (bins 0..N/2) + N out
                kind, inv = ("q31" if "31" in op else "q15"), int(op.endswith("inv"))
                if N < 32:
                    continue
                esz = 4 if kind == "q31" else 2
                B = nbytes // (2 * N * esz)
                cd.ensure_rfft_fix_plans(kind, N)
                tdt = torch.int32 if kind == "q31" else torch.int16
                a = torch.randint(-2**13, 2**13, (B, 2 * N if inv else N), device=dev, dtype=tdt)
                b = torch.empty(B, N if inv else 2 * N, device=dev, dtype=tdt)
                fn = lambda: cd.rfft_fix_device(kind, N, a.data_ptr(), b.data_ptr(), B, inv, st)
                alg = B * esz * ((N + 2 + N) if inv else 3 * N)
                samples = B * N
                info = cd.kernel_info((5 if kind == "q31" else 7) + inv, N)
            else:
                B = nbytes // (4 * N)
                cd.ensure_rfft_plans(N)
                a = torch.randn(B, N, device=dev)
                b = torch.empty_like(a)
                inv = int(op == "rfft_inv")
                fn = lambda: cd.rfft_device(N, a.data_ptr(), b.data_ptr(), B, inv, st)
                alg = 2 * B * N * 4
                samples = B * N
                info = cd.kernel_info(4 if inv else 3, N)
            ms = timed(fn, args.reps, st, args.warm)
            gbs = alg / ms / 1e6
            row = dict(op=op, N=N, frames=B, ms=ms, gsamples=samples / ms / 1e6, gbs=gbs, frac=gbs / peak, **info)
            rows.append(row)
            print(f"{op:9s} N={N:5d} B={B:9d} {ms:8.4f} ms {row['gsamples']:8.1f} GS/s {gbs:8.1f} GB/s {100*gbs/peak:5.1f}% "
                  f"regs={info['regs_per_thread']} cta/sm={info['ctas_per_sm']} thr={info['threads_per_cta']} smem={info['smem_bytes']}", flush=True)
            del fn
            torch.cuda.empty_cache()
    if args.json:
        json.dump(dict(peak_gbs=peak, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
