#!/usr/bin/env python
"""Probe: how fast can a chunked H2D -> kernel -> D2H pipeline over pinned buffers go on this host link, by stream layout?
   A  chunk k entirely on stream k % S (the library's layout)
   B  three role streams (all H2D on one, all kernels on one, all D2H on one) tied by events, S staging slots
   C  chunked copies without dependencies (H2D chunks on one stream, D2H chunks on another): bound for chunked duplex
   D  one 1 GiB copy per direction at once (bench.py's pinned_copy_peak "both")
The "kernel" is a device-to-device copy of the chunk (same order of time as the FFT kernel: ~10 us per 32 MiB)."""
import sys
import time
import torch

dev = torch.device("cuda", 0)
TOTAL = 1 << 30
h_in = torch.empty(TOTAL // 4, dtype=torch.float32).pin_memory()
h_out = torch.empty(TOTAL // 4, dtype=torch.float32).pin_memory()
h_in.normal_()


def run(fn, reps=4):
    best = 1e9
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if r:
            best = min(best, dt)
    return best


def layout_a(chunk, S):
    n = chunk // 4
    din = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(S)]
    dout = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(S)]
    st = [torch.cuda.Stream(device=dev) for _ in range(S)]

    def fn():
        for k in range(TOTAL // chunk):
            s = k % S
            with torch.cuda.stream(st[s]):
                din[s].copy_(h_in[k * n:(k + 1) * n], non_blocking=True)
                dout[s].copy_(din[s])
                h_out[k * n:(k + 1) * n].copy_(dout[s], non_blocking=True)
    return fn


def layout_b(chunk, S):
    n = chunk // 4
    din = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(S)]
    dout = [torch.empty(n, dtype=torch.float32, device=dev) for _ in range(S)]
    sh, sk, sd = (torch.cuda.Stream(device=dev) for _ in range(3))

    def fn():
        evk, evd = {}, {}
        for k in range(TOTAL // chunk):
            s = k % S
            with torch.cuda.stream(sh):
                if k - S in evk:
                    sh.wait_event(evk[k - S])          # the kernel of the slot's previous chunk has read din[s]
                din[s].copy_(h_in[k * n:(k + 1) * n], non_blocking=True)
                eh = torch.cuda.Event()
                eh.record(sh)
            with torch.cuda.stream(sk):
                sk.wait_event(eh)
                if k - S in evd:
                    sk.wait_event(evd[k - S])          # dout[s] has gone home
                dout[s].copy_(din[s])
                evk[k] = torch.cuda.Event()
                evk[k].record(sk)
            with torch.cuda.stream(sd):
                sd.wait_event(evk[k])
                h_out[k * n:(k + 1) * n].copy_(dout[s], non_blocking=True)
                evd[k] = torch.cuda.Event()
                evd[k].record(sd)
    return fn


def layout_c(chunk):
    n = chunk // 4
    d1 = torch.empty(TOTAL // 4, dtype=torch.float32, device=dev)
    d2 = torch.empty(TOTAL // 4, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def fn():
        for k in range(TOTAL // chunk):
            with torch.cuda.stream(s1):
                d1[k * n:(k + 1) * n].copy_(h_in[k * n:(k + 1) * n], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[k * n:(k + 1) * n].copy_(d2[k * n:(k + 1) * n], non_blocking=True)
    return fn


def report(name, dt):
    print(f"{name:40s} {dt * 1e3:7.2f} ms  {TOTAL / dt / 1e9:6.2f} GB/s per direction", flush=True)


report("D  one copy per direction", run(layout_c(TOTAL)))
for chunk_mib in (8, 32, 64):
    report(f"C  independent chunks {chunk_mib} MiB", run(layout_c(chunk_mib << 20)))
for chunk_mib in (8, 16, 32, 64):
    for S in (2, 3, 4):
        report(f"A  chunk {chunk_mib} MiB, {S} streams", run(layout_a(chunk_mib << 20, S)))
        report(f"B  chunk {chunk_mib} MiB, {S} slots, role streams", run(layout_b(chunk_mib << 20, S)))
