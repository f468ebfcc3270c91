#!/usr/bin/env python
"""BASELINE config 4: arm_mfcc_f32 front end (window + rfft_fast 1024 + mel + log + DCT) on synthetic 16 kHz audio,
the frame batch sharded across the GPUs of one box (weak scaling: --frames frames per GPU, no collective on the data
path), with the compiled reference's arm_mfcc_f32 timed on the host cores beside it.

    python tools/bench_config4.py [--json out.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29519 \
        tools/bench_config4.py --json out.json

Audio = 3 sines + 0.1 N(0,1), non-overlapping 1024-sample frames (--hop for
overlap), the reference's own 20-mel / 13-DCT / Hamming configuration (Testing/Source/Tests/mfccdata.c via
tests/golden/mfcc_patterns.npz).  Time = CUDA events on the launch stream between barriers, MAX over ranks.
Algorithmic bytes per frame = 4*hop (new samples read) + 4*13 written; with hop = fftLen that is 4148 B.
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


def cpu_reference(cfg, n, hop, seconds=1.0):
    import numpy as np
    from oracle_lib import oracle, ref
    lib = ref(fast=True) or oracle()
    cores = os.cpu_count() or 1
    frames = 4096 * max(1, cores // 4)
    t = np.arange((frames - 1) * hop + n) / 16000.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1300 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t)
         + 0.1 * np.random.default_rng(4).standard_normal(t.size)).astype(np.float32)
    best, t_end = float("inf"), time.perf_counter() + seconds
    while True:
        t0 = time.perf_counter()
        lib.mfcc(cfg, x, stride=hop, frames=frames, threads=cores)
        best = min(best, time.perf_counter() - t0)
        if time.perf_counter() > t_end:
            break
    return dict(frames_per_s=frames / best, cores=cores, sample_frames=frames, kind="reference" if ref(fast=True) else "port")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fft-len", type=int, default=1024)
    ap.add_argument("--hop", type=int, default=0, help="0 = fftLen (non-overlapping)")
    ap.add_argument("--frames", type=int, default=1 << 18, help="frames per GPU")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    from oracle_lib import mfcc_config
    n, hop = args.fft_len, args.hop or args.fft_len
    cfg = mfcc_config(n)
    cpu = cpu_reference(cfg, n, hop) if rank == 0 else None       # before torch / CUDA are initialised (thread affinity)

    import numpy as np
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
    B = args.frames
    lo, hi = cd.shard_frames(B * world, world, rank)
    assert hi - lo == B
    m = cd.Mfcc(cfg)
    nsamp = (B - 1) * hop + n
    t = (torch.arange(nsamp, device=dev, dtype=torch.float64) + lo * hop) / 16000.0
    x = (0.5 * torch.sin(2 * np.pi * 440 * t) + 0.3 * torch.sin(2 * np.pi * 1300 * t) + 0.2 * torch.sin(2 * np.pi * 3100 * t)).float()
    x += 0.1 * torch.randn(nsamp, device=dev, generator=torch.Generator(device=dev).manual_seed(4 + rank))
    del t
    out = torch.empty(B, 13, device=dev)
    L = cd.lib()

    def step():
        assert L.arm_mfcc_batch_f32(C.byref(m.S), x.data_ptr(), hop, out.data_ptr(), B) == 0, cd.last_error()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # arm_mfcc_batch_f32 runs on the library's own stream and returns after completion: time it on the host too
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / args.reps
    tm = torch.tensor([ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    # parity of a stratified subsample against the oracle with the reference's own MFCC thresholds
    if rank == 0:
        from oracle_lib import oracle
        idx = np.arange(0, B, max(1, B // 256))
        frames = torch.stack([x[i * hop:i * hop + n] for i in idx]).cpu().numpy()
        want = oracle().mfcc(cfg, frames.reshape(-1), threads=os.cpu_count() or 1)
        got = out[torch.from_numpy(idx).to(dev)].cpu().numpy()
        ok = bool(np.all(np.abs(got - want) <= 1e-5 + 1.2e-3 * np.abs(want)))
        peak = 6536.7
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peak = json.load(open(pk))["hbm_gbs"]
        fps = world * B / (ms * 1e-3)
        gbs = fps * (4 * hop + 4 * 13) / 1e9
        res = dict(metric="arm_mfcc_f32 frames/s (BASELINE config 4)", fft_len=n, hop=hop, gpus=world, frames_per_gpu=B,
                   ms_per_step=ms, frames_per_s=fps, msamples_per_s=fps * hop / 1e6, algorithmic_gbs=gbs,
                   frac_of_hbm_peak=gbs / (peak * world), parity_ok=ok, cpu_reference=cpu,
                   speedup_vs_cpu=fps / cpu["frames_per_s"], scaling="weak",
                   timing="host clock around reps synchronous library calls (each returns after stream completion), max over ranks")
        print(json.dumps(res))
        if args.json:
            json.dump(res, open(args.json, "w"), indent=1)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
