#!/usr/bin/env python
"""Host staging sweep of the end-to-end path: arm_rfft_fast_batch_f32 (N = 4096, forward + inverse, pinned host
buffers) for every (chunk MiB, streams) pair, beside the pinned-copy peak of the same bytes.

    python tools/sweep_staging.py [--frames 65536] [--json out.json]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cmsis-dsp_b200", "python"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cmsisdsp_b200 as cd  # noqa: E402
from bench import pinned_copy_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=65536)
    ap.add_argument("--chunks", default="4,8,16,32,64,128,256")
    ap.add_argument("--streams", default="2,3,4,6")
    ap.add_argument("--ramps", default="0", help="first / last chunk MiB of the geometric ramp (0: none)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    N, B = 4096, args.frames
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    cd.cuda().cmsisdsp_cuda_set_device(0)
    cd.set_devices([0])
    L, S = cd.lib(), cd.rfft_instance(N)
    hx = torch.randn(B, N).pin_memory()
    hs = torch.empty_like(hx).pin_memory()
    hy = torch.empty_like(hx).pin_memory()
    nbytes = B * N * 4
    peak = pinned_copy_peak(torch, dev, nbytes)
    print("pinned copy peak:", json.dumps(peak), flush=True)
    rows = []
    for chunk in [int(v) for v in args.chunks.split(",")]:
        for ns, ramp in [(int(v), int(r)) for v in args.streams.split(",") for r in args.ramps.split(",")]:
            assert L.arm_cuda_set_staging(chunk, ns) == 0
            assert L.arm_cuda_set_staging_ramp(ramp) == 0
            best = float("inf")
            for rep in range(4):
                t0 = time.perf_counter()
                a = L.arm_rfft_fast_batch_f32(C.byref(S), hx.data_ptr(), hs.data_ptr(), B, 0)
                b = L.arm_rfft_fast_batch_f32(C.byref(S), hs.data_ptr(), hy.data_ptr(), B, 1)
                dt = time.perf_counter() - t0
                assert a == 0 and b == 0, cd.last_error()
                if rep:
                    best = min(best, dt)
            gbs = 2 * nbytes / best / 1e9
            rows.append(dict(chunk_mib=chunk, streams=ns, ramp_mib=ramp, seconds=best, msamples=B * N / best / 1e6, pcie_gbs_per_direction=gbs,
                             frac_of_copy_peak=gbs / peak["both_gbs_per_direction"]))
            print(f"chunk {chunk:4d} MiB  streams {ns}  ramp {ramp:3d} MiB  {best * 1e3:8.1f} ms  {B * N / best / 1e6:8.0f} Msamples/s  {gbs:6.1f} GB/s per direction  "
                  f"{100 * gbs / peak['both_gbs_per_direction']:5.1f}% of the pinned-copy peak", flush=True)
    err = float(((hy[:64] - hx[:64]).double().pow(2).sum() / hx[:64].double().pow(2).sum()).sqrt())
    print("round trip rel-RMS", err)
    if args.json:
        json.dump(dict(peak=peak, rows=rows, roundtrip=err), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
