#!/usr/bin/env python
"""BASELINE config 5: arm_cfft_f32 length sweep 16..4096 x batch sweep 2^10..2^24 frames, STRONG-scaled over the
GPUs of one box, with the host-CPU reference timed beside it.

    python tools/sweep_config5.py [--json out.json]                                 # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29517 \
        tools/sweep_config5.py --json out.json                                      # G GPUs

For every (N, B) the frame range is block-partitioned over the ranks (cmsisdsp_b200.shard_frames, no collective on the
data path); each rank transforms its shard in place, resident in HBM; time = CUDA events on the launch stream between
two barriers, MAX over ranks (torch.distributed is used for the barrier and that max only).  A shard larger than
--cap-gib is processed as consecutive resident chunks of that size over the same buffer and the row is flagged
"chunked" (SURVEY.md section 8(d) config 5).  Inputs smaller than L2 are L2-resident between repetitions: those rows
are flagged "l2" and say nothing about HBM.
The CPU column is the compiled reference (oracle/_ref, gcc -O3 generic-C build) on all host cores over a bounded
sample of the same length, rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "cmsis-dsp_b200", "python"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def cpu_column(lengths, seconds=0.25):
    """Gsamples/s of the reference's arm_cfft_f32 on all host cores, per length (bounded sample)."""
    import numpy as np
    from oracle_lib import oracle, ref
    lib = ref(fast=True) or oracle()
    cores = os.cpu_count() or 1
    out = {}
    for N in lengths:
        frames = max(cores * 8, (1 << 22) // N)                    # ~32 MiB per sample
        x = np.random.default_rng(N).standard_normal((frames, 2 * N)).astype(np.float32)
        fn = lib._fn("cfft_f32_batch")
        best, t_end = float("inf"), time.perf_counter() + seconds
        while True:
            y = x.copy()
            t0 = time.perf_counter()
            fn(N, y.ctypes.data, frames, 0, 1, cores)
            best = min(best, time.perf_counter() - t0)
            if time.perf_counter() > t_end:
                break
        out[N] = dict(gsamples=frames * N / best / 1e9, frames=frames, cores=cores)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lens", default="16,32,64,128,256,512,1024,2048,4096")
    ap.add_argument("--log2-batches", default="10,12,14,16,18,20,22,24")
    ap.add_argument("--cap-gib", type=float, default=8.0, help="resident bytes per GPU")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    lengths = [int(v) for v in args.lens.split(",")]
    cpu = cpu_column(lengths) if rank == 0 else {}     # before CUDA/torch are initialised in this process (thread affinity)

    import torch
    import cmsisdsp_b200 as cd
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    cu = cd.cuda()
    cu.cmsisdsp_cuda_set_device(local)
    st = torch.cuda.current_stream().cuda_stream
    peak = 6536.7
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk))["hbm_gbs"]
    cap = int(args.cap_gib * (1 << 30))
    rows = []
    for N in lengths:
        cd.ensure_plans("f32", N)
        for lb in [int(v) for v in args.log2_batches.split(",")]:
            B = 1 << lb
            lo, hi = cd.shard_frames(B, world, rank)
            mine = hi - lo
            shard_bytes = mine * N * 8
            resident = min(mine, max(1, cap // (N * 8)))
            chunks = -(-mine // resident) if mine else 0
            buf = torch.randn(max(resident, 1) * 2 * N, device=dev, dtype=torch.float32)

            calls = [0]

            def run():
                left = mine
                while left > 0:
                    n = min(left, resident)
                    cd.cfft_device("f32", N, buf.data_ptr(), n, calls[0] & 1, 1, st)   # forward / inverse alternate: values stay bounded
                    calls[0] += 1
                    left -= n

            for _ in range(3):
                run()
                buf.normal_()                       # keep values bounded over repeated in-place transforms
            torch.cuda.synchronize()
            if dist:
                dist.barrier()
            t = C.c_void_p()
            cu.cmsisdsp_cuda_timer_begin(C.byref(t), st)
            for _ in range(args.reps):
                run()
            ms = C.c_float()
            cu.cmsisdsp_cuda_timer_end(t, st, C.byref(ms))
            tm = torch.tensor([ms.value / args.reps], device=dev)
            if dist:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            ms_step = float(tm.item())
            if rank == 0:
                gs = B * N / ms_step / 1e6
                gbs = 2 * B * N * 8 / ms_step / 1e6
                flags = [f for f, on in (("chunked", chunks > 1), ("l2", world * 0 + shard_bytes <= 100 << 20)) if on]
                rows.append(dict(N=N, log2_frames=lb, frames=B, gpus=world, ms=ms_step, gsamples=gs, gbs=gbs,
                                 frac_of_peak_all_gpus=gbs / (peak * world), flags=flags,
                                 cpu_gsamples=cpu[N]["gsamples"], speedup_vs_cpu=gs / cpu[N]["gsamples"]))
                print(f"N={N:5d} B=2^{lb:<2d} gpus={world} {ms_step:9.4f} ms {gs:9.1f} GS/s {gbs:9.1f} GB/s "
                      f"{100 * gbs / (peak * world):6.1f}% of {world}x peak  cpu {cpu[N]['gsamples']:6.2f} GS/s  x{gs / cpu[N]['gsamples']:7.1f} "
                      f"{','.join(flags)}", flush=True)
            del buf
            torch.cuda.empty_cache()
    if rank == 0 and args.json:
        json.dump(dict(peak_gbs_per_gpu=peak, gpus=world, cap_gib=args.cap_gib, cpu=cpu, rows=rows), open(args.json, "w"), indent=1)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
