#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 CMSIS-DSP FFT hot path.

Workload (BASELINE.json configs[1]): arm_rfft_fast_f32, N = 4096, forward + inverse,
65536 frames of synthetic real noise per GPU.  One "step" = forward over the whole batch
followed by inverse over the whole batch (two kernel launches).  Metric: Msamples/s
(real samples: frames * N per forward+inverse pair).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Rank 0 prints ONE JSON line.  Frames are independent, so ranks share nothing: the batch is
block-partitioned (weak scaling: 65536 frames per rank), no collective on the data path;
torch.distributed is used only for the barrier and the max-over-ranks of the device time.

Beside the contract's keys the line carries
  roofline.sustained   the same two kernels back to back for >= 3 s (clock / power samples inside): the burst figure of
                       the timed region and the figure the chip holds once power management settles, side by side
  secondary            live-timed rooflines of the other kernels the north_star and BASELINE configs name:
                       arm_cfft_f32 N=1024, arm_cfft_q31 / q15 N=256/1024/4096 at 2^20 frames (config 3), MFCC (config 4),
                       each checked against the oracle on a stratified sample of its frames
  parity               every rank's stratified oracle check of ITS shard (forward and inverse), max over ranks
  e2e                  through the C API with pinned host buffers, each rank on its own device; pcie_gbs_per_direction
                       beside a measured pinned-copy peak of the same bytes (the copy roofline of e2e); at --gpus N > 1
                       also one process driving all N devices through the C dispatcher (arm_cuda_set_devices)

`--impl reference` times the reference's own generic-C CPU implementation
(oracle/_ref/libcmsisdsp_ref_fast.so, built from /root/reference's sources) on all host
cores, same workload; rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "cmsis-dsp_b200", "python"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_REAL = 4096
FRAMES_PER_GPU = 65536
METRIC = "batched FFT Msamples/s (arm_rfft_fast_f32 N=4096 forward+inverse)"
UNIT = "Msamples/s"
WORKLOAD = "arm_rfft_fast_f32 N=4096 forward+inverse, batch 65536 frames per GPU"
F32_TOL = 2e-6
# identical in both arms (the driver compares the two lines' `config`)
CONFIG = {"workload": WORKLOAD, "frames_per_gpu": FRAMES_PER_GPU, "fft_len": N_REAL,
          "l2_policy": "inputs larger than L2 (1 GiB per buffer per launch vs 126 MB L2)",
          "timing": "cuda arm: CUDA events on the launch stream, max over ranks; reference arm: wall clock around the threaded batch call"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every ~20 ms during the timed region
    (falls back to polling nvidia-smi when NVML is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.sm_max, self.reasons = [], None, set()
        self.mem, self.power = [], []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        try:                                          # HBM clock and board power: the bench is HBM-bound
            self.mem.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_MEM))
            self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for name, bit in (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown),
                          ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown),
                          ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        t = [v.strip() for v in out.split(",")]
        if len(t) >= 6 and t[0].isdigit():
            self.sm.append(int(t[0]))
            self.sm_max = int(t[1]) if t[1].isdigit() else self.sm_max
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.02 if self.nvml else 0.2)

    def finish(self):
        self.stop_flag.set()
        self.join(timeout=2)
        return self.summary()

    def summary(self):
        sm = sorted(self.sm)
        mem = sorted(self.mem)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi",
                "mem_mhz": mem[len(mem) // 2] if mem else None, "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------- the CPU arm

def cpu_baseline_subprocess(frames, steps, warmup):
    """cpu_reference_run in a fresh interpreter: the threads of a process that has initialised
    torch/CUDA/NVML inherit a CPU affinity of one core on some boxes, which would make the
    all-cores CPU baseline read many times too low."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only", str(frames), "--steps", str(steps),
                          "--warmup", str(warmup)], capture_output=True, text=True, timeout=900)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu baseline subprocess failed: " + out.stderr[-400:])


def cpu_reference_run(frames, steps, warmup, fast=True):
    """The reference's generic-C arm_rfft_fast_f32 forward+inverse over `frames` frames on all host cores,
    `warmup` untimed + `steps` timed repetitions on the SAME input (generated once).  Used by both arms, so
    cpu_baseline (cuda arm) and the reference arm's line are the same statistic of the same code."""
    import numpy as np
    from oracle_lib import oracle, ref
    try:
        os.sched_setaffinity(0, set(range(os.cpu_count() or 1)))
    except (AttributeError, OSError):
        pass
    lib, kind = ref(fast=fast), "reference"
    if lib is None:
        lib, kind = oracle(), "port"
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(1)
    x = rng.standard_normal((frames, N_REAL), dtype=np.float32)
    xin, spec, y = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    fn = lib._fn("rfft_fast_f32_batch")
    times = []
    for k in range(warmup + steps):
        np.copyto(xin, x)                                # the forward transform destroys its input (untimed refill)
        t0 = time.perf_counter()
        fn(N_REAL, xin.ctypes.data, spec.ctypes.data, frames, 0, cores)
        fn(N_REAL, spec.ctypes.data, y.ctypes.data, frames, 1, cores)
        dt = time.perf_counter() - t0
        if k >= warmup:
            times.append(dt)
    err = float(np.sqrt(((y.astype(np.float64) - x) ** 2).sum() / (x.astype(np.float64) ** 2).sum()))
    mean = sum(times) / len(times)
    srt = sorted(times)
    rate = lambda s: frames * N_REAL / s / 1e6
    return {"value": rate(mean), "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{frames} frames fwd+inv per step, {warmup} warm-up + {steps} timed steps on one input, {cores} pthreads, "
                      f"gcc -O3 generic-C build of the reference's own sources",
            "statistic": "mean over the timed steps (value); best and median beside it",
            "value_best": rate(srt[0]), "value_median": rate(srt[len(srt) // 2]),
            "seconds_per_step": mean, "step_seconds": times, "roundtrip_relrms": err}


def run_reference(args, rank):
    if rank != 0:
        return
    frames = FRAMES_PER_GPU if (os.cpu_count() or 1) >= 16 else FRAMES_PER_GPU // 4
    cb = cpu_reference_run(frames, args.steps, max(args.warmup, 1))
    value, sec = cb["value"], cb["seconds_per_step"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": CONFIG, "frames_timed_per_step": frames,
        "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- the CUDA arm

def _relrms(a, b):
    import numpy as np
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def timed_launches(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def secondary_rooflines(torch, cd, dev, stream, peak, quick):
    """Live-timed rooflines of the other kernels on the path (same schema as `roofline`), each with an oracle check of a
    stratified sample of its frames.  Algorithmic bytes per frame as in DESIGN.md section 4 / SURVEY.md 8(d)."""
    import numpy as np
    from oracle_lib import mfcc_config, oracle
    rows = []
    B = (1 << 18) if quick else (1 << 20)
    big = torch.empty(B * 4096 * 2, dtype=torch.int32, device=dev)          # 32 GiB at 2^20 frames: cfft_q31 N=4096 in place

    def row(name, workload, frames, alg_bytes, ms, dtype, ok, err):
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        rows.append({"kernel": name, "workload": workload, "frames": frames, "dtype": dtype, "bound": "hbm", "ms": ms,
                     "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_launch": alg_bytes,
                     "gsamples_per_s": None, "oracle_check": {"ok": bool(ok), "detail": err}})

    # arm_cfft_f32 N=1024 (north_star target kernel #1), 2^20 frames in place
    N, frames = 1024, B
    cd.ensure_plans("f32", N)
    buf = big[:frames * 2 * N].view(torch.float32)
    buf.normal_(generator=torch.Generator(device=dev).manual_seed(7))
    idx = torch.arange(0, frames, frames // 64, device=dev)
    x0 = buf.view(frames, 2 * N)[idx].cpu().numpy()
    cd.cfft_device("f32", N, buf.data_ptr(), frames, 0, 1, stream)
    torch.cuda.synchronize()
    err = _relrms(buf.view(frames, 2 * N)[idx].cpu().numpy(), oracle().cfft("f32", N, x0, 0, 1))
    ms = timed_launches(torch, lambda: cd.cfft_device("f32", N, buf.data_ptr(), frames, 0, 1, stream), 10)
    row("cfft_f32", f"arm_cfft_f32 N={N} forward, bitReverseFlag=1, batch {frames} frames in place", frames, 16 * N * frames, ms, "f32",
        err <= F32_TOL, f"rel-RMS vs oracle on 64 frames {err:.2e}")
    rows[-1]["gsamples_per_s"] = frames * N / (ms * 1e-3) / 1e9

    # BASELINE config 3: arm_cfft_q31 / arm_cfft_q15 N=256/1024/4096, 2^20 frames, bit-exact
    for kind, esz, tdt, ndt in (("q31", 8, torch.int32, np.int32), ("q15", 4, torch.int16, np.int16)):
        for N in (256, 1024, 4096):
            frames = B
            cd.ensure_plans(kind, N)
            words = frames * N * esz // 4
            raw = big[:words]
            raw.random_(-2 ** 31, 2 ** 31 - 1, generator=torch.Generator(device=dev).manual_seed(N))
            view = raw.view(tdt).view(frames, 2 * N)
            idx = torch.arange(0, frames, frames // 64, device=dev)
            x0 = view[idx].cpu().numpy()
            cd.cfft_device(kind, N, raw.data_ptr(), frames, 0, 1, stream)
            torch.cuda.synchronize()
            ok = np.array_equal(view[idx].cpu().numpy(), oracle().cfft(kind, N, x0.astype(ndt), 0, 1))
            ms = timed_launches(torch, lambda: cd.cfft_device(kind, N, raw.data_ptr(), frames, 0, 1, stream), 6 if N == 4096 else 10)
            row(f"cfft_{kind}", f"arm_cfft_{kind} N={N} forward, bitReverseFlag=1, batch {frames} frames in place (full-range uniform input)",
                frames, 2 * esz * N * frames, ms, kind, ok, "bit-exact vs oracle on 64 frames" if ok else "MISMATCH vs oracle")
            rows[-1]["gsamples_per_s"] = frames * N / (ms * 1e-3) / 1e9
    del big
    torch.cuda.empty_cache()

    # BASELINE config 4: arm_mfcc_f32 front end, N=1024, 20 mel, 13 DCT, non-overlapping frames of synthetic 16 kHz audio
    n, frames = 1024, (1 << 16) if quick else (1 << 18)
    cfg = mfcc_config(n)
    m = cd.Mfcc(cfg)
    t = torch.arange(frames * n, device=dev, dtype=torch.float64) / 16000.0
    x = (0.5 * torch.sin(2 * np.pi * 440 * t) + 0.3 * torch.sin(2 * np.pi * 1300 * t) + 0.2 * torch.sin(2 * np.pi * 3100 * t)).float()
    x += 0.1 * torch.randn(frames * n, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
    del t
    out = torch.empty(frames, 13, device=dev)
    L = cd.lib()
    torch.cuda.synchronize()
    assert L.arm_mfcc_batch_f32(C.byref(m.S), x.data_ptr(), n, out.data_ptr(), frames) == 0, cd.last_error()   # creates the plan
    idx = torch.arange(0, frames, frames // 64, device=dev)
    want = oracle().mfcc(cfg, x.view(frames, n)[idx].cpu().numpy().reshape(-1))
    got = out[idx].cpu().numpy()
    ok = bool(np.all(np.abs(got - want) <= 1e-5 + 1.2e-3 * np.abs(want)))
    ms = timed_launches(torch, lambda: L.arm_mfcc_batch_f32(C.byref(m.S), x.data_ptr(), n, out.data_ptr(), frames), 10)
    row("mfcc_f32", f"arm_mfcc_f32 N={n}, 20 mel filters, 13 DCT outputs, hop {n}, batch {frames} frames (device buffers, C API call incl. its stream sync)",
        frames, frames * (4 * n + 52), ms, "f32", ok, f"max |err| vs oracle on 64 frames {float(np.abs(got - want).max()):.2e} (reference thresholds)")
    rows[-1]["mframes_per_s"] = frames / (ms * 1e-3) / 1e6
    rows[-1]["note"] = "SM-bound by the real FFT, not HBM-bound (DESIGN.md section 4): frac is quoted against HBM for comparability only"
    # not measured live: the kernel's own floor from the stage ablation recorded under profiles/ (timing-only builds without the
    # mel / log / DCT stages), so that the live `ms` can be read against it
    fft_only_ms = 0.3655 * (frames * 4 * n) / 2**30
    rows[-1]["floor"] = {"fft_only_ms": fft_only_ms, "frac_of_floor": fft_only_ms / ms,
                         "source": "profiles/r2_ap_mfcc_ablation.txt (window + real FFT + magnitudes alone: 0.3655 ms per GiB of samples, "
                                   "45.5 % of the HBM figure -- more than arm_rfft_fast_f32 forward needs with its output, 0.348 ms)"}
    return rows


def pinned_copy_peak(torch, dev, nbytes, barrier=None, allmax=None):
    """The copy roofline of the end-to-end path: pinned host <-> device copies of the same byte count, both
    directions at once on two streams (what a perfectly overlapped H2D / kernel / D2H pipeline is bounded by), and each
    direction alone.  With several ranks every repetition starts at a barrier and the slowest rank's time counts, so the
    figure is what ONE rank gets while ALL ranks copy (the host's link is shared)."""
    h_in = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    h_out = torch.empty(nbytes // 4, dtype=torch.float32).pin_memory()
    d_in = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    d_out = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        best = float("inf")
        for _ in range(3):
            torch.cuda.synchronize()
            if barrier:
                barrier()
            t0 = time.perf_counter()
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = min(best, allmax(dt) if allmax else dt)
        res[mode + "_gbs_per_direction"] = nbytes / best / 1e9
    return res


def run_cuda(args, rank, world, local_rank):
    import numpy as np
    import torch
    import cmsisdsp_b200 as cd
    from oracle_lib import oracle

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # keep this rank's host thread (and so its pinned buffers, first touch) on the NUMA node of its GPU
    # (matters on multi-socket hosts when several ranks stream host buffers at once; a no-op on one node)
    numa_bound = False
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa_bound = True
    except Exception:
        pass
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        # NCCL_DEBUG is left as the launcher set it: with VERSION / WARN / INFO the library writes its own banner to stdout
        # before rank 0's JSON line, which stays the LAST line of stdout
        dist.init_process_group("nccl", device_id=dev)
    cu = cd.cuda()
    cu.cmsisdsp_cuda_set_device(local_rank)
    cd.ensure_rfft_plans(N_REAL)
    warmup = max(args.warmup, 3)

    def allmax(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = FRAMES_PER_GPU
    lo, hi = cd.shard_frames(B * world, world, rank)       # weak scaling: B frames per rank
    assert hi - lo == B
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    x = torch.randn(B, N_REAL, device=dev, dtype=torch.float32, generator=gen)
    spec = torch.empty_like(x)
    y = torch.empty_like(x)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        cd.rfft_device(N_REAL, x.data_ptr(), spec.data_ptr(), B, 0, stream)
        cd.rfft_device(N_REAL, spec.data_ptr(), y.data_ptr(), B, 1, stream)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()

    # parity of THIS rank's shard (outside the timed region): a stratified sample of its frames, forward spectrum and
    # inverse result against the oracle, plus the round trip; reduced with MAX over the ranks
    idx = torch.arange(0, B, B // 256, device=dev)
    xs = x[idx].cpu().numpy()
    want_spec = oracle().rfft(N_REAL, xs, 0, threads=min(16, os.cpu_count() or 1))
    got_spec = spec[idx].cpu().numpy()
    e_fwd = _relrms(got_spec, want_spec)
    e_inv = _relrms(y[idx].cpu().numpy(), oracle().rfft(N_REAL, got_spec, 1, threads=min(16, os.cpu_count() or 1)))
    rt = float(((y[idx] - x[idx]).double().pow(2).sum() / x[idx].double().pow(2).sum()).sqrt())
    parity = {"frames_checked_per_rank": int(idx.numel()), "forward_relrms_max_over_ranks": allmax(e_fwd),
              "inverse_relrms_max_over_ranks": allmax(e_inv), "roundtrip_relrms_max_over_ranks": allmax(rt), "tolerance": F32_TOL,
              "how": "every rank: 256 frames spread over its own shard, CUDA result vs the CPU oracle on the same input"}
    parity["ok"] = max(parity["forward_relrms_max_over_ranks"], parity["inverse_relrms_max_over_ranks"],
                       parity["roundtrip_relrms_max_over_ranks"]) <= F32_TOL

    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = cu.cmsisdsp_cuda_launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record()
        cd.rfft_device(N_REAL, x.data_ptr(), spec.data_ptr(), B, 0, stream)
        ev[k][1].record()
        cd.rfft_device(N_REAL, spec.data_ptr(), y.data_ptr(), B, 1, stream)
        ev[k][2].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = cu.cmsisdsp_cuda_launch_count() - launches0
    clocks = sampler.finish()

    total_ms = allmax(ev[0][0].elapsed_time(ev[-1][2]))
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    ms_per_step = total_ms / args.steps
    value = world * B * N_REAL / (ms_per_step * 1e-3) / 1e6

    # roofline of the dominant kernel: algorithmic bytes = one read + one write of the payload
    peak, peak_src = load_peaks()
    bytes_per_launch = B * N_REAL * 4 * 2                 # 8N bytes per frame
    dom_name, dom_ms = ("rfft_fwd", fwd_ms) if fwd_ms >= inv_ms else ("rfft_inv", inv_ms)
    achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": dom_name, "peak_source": peak_src,
                "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                "fwd_gbs": bytes_per_launch / (fwd_ms * 1e-3) / 1e9, "inv_gbs": bytes_per_launch / (inv_ms * 1e-3) / 1e9,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "regime": f"burst: {args.steps} steps = {total_ms:.0f} ms of device time (see `sustained` for the long-run figure)"}
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        with open(tr) as f:
            roofline["traffic"] = json.load(f).get(dom_name)

    # sustained: the same two kernels back to back for >= 3 s, clocks and power sampled inside; then a plain device copy
    # of the same size for the same time (what the memory system itself holds over such a region)
    if not args.no_sustained:
        sus_s = 1.0 if args.quick else 3.0
        chunk = 200
        s2 = ClockSampler(local_rank)
        s2.start()
        seg = []
        barrier()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < sus_s:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(chunk):
                step()
            e1.record()
            torch.cuda.synchronize()
            seg.append(e0.elapsed_time(e1) / chunk)
        sus_clocks = s2.finish()
        last = seg[len(seg) // 2:]                          # the settled half
        sus_ms = sum(last) / len(last)
        sus_gbs = 2 * bytes_per_launch / (sus_ms * 1e-3) / 1e9
        ca, cb = x.view(-1), y.view(-1)
        for _ in range(3):
            cb.copy_(ca)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ncopy = max(50, int(0.5 * sus_s / 0.00033))
        c0.record()
        for _ in range(ncopy):
            cb.copy_(ca)
        c1.record()
        torch.cuda.synchronize()
        copy_gbs = ncopy * 2 * ca.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        roofline["sustained"] = {
            "seconds": time.perf_counter() - t0, "steps": chunk * len(seg), "ms_per_step_first_segment": seg[0], "ms_per_step_settled": sus_ms,
            "achieved": sus_gbs, "frac": sus_gbs / peak, "frac_of_sustained_copy": sus_gbs / copy_gbs, "sustained_copy_gbs": copy_gbs,
            "value_msamples": world * B * N_REAL / (sus_ms * 1e-3) / 1e6, "clocks": sus_clocks,
            "how": f"forward+inverse launches back to back in segments of {chunk} steps for >= {sus_s:.0f} s, CUDA events per segment; settled = mean "
                   f"of the second half of the segments; achieved = 2 * algorithmic bytes per launch / settled step time (average of both kernels); "
                   f"copy = {ncopy} back-to-back 1 GiB torch copy_ launches"}

    secondary = None
    if rank == 0 and not args.no_secondary:
        try:
            secondary = secondary_rooflines(torch, cd, dev, stream, peak, args.quick)
        except Exception as e:                              # diagnostic lines must not take the headline down
            secondary = [{"error": repr(e)[:300]}]
    barrier()

    # end to end through the public C API with pinned HOST buffers (H2D + kernels + D2H timed); every rank drives ITS
    # device (arm_cuda_set_devices), as a caller with one process per GPU would
    e2e = None
    if not args.no_e2e:
        del spec, y
        torch.cuda.empty_cache()
        Be = B // 4 if args.quick else B
        S = cd.rfft_instance(N_REAL)
        hx = torch.empty(Be, N_REAL, dtype=torch.float32).pin_memory()
        hs = torch.empty_like(hx).pin_memory()
        hy = torch.empty_like(hx).pin_memory()
        hx.normal_(generator=torch.Generator().manual_seed(99 + rank))
        L = cd.lib()
        cd.set_devices([local_rank])

        def e2e_step():
            a = L.arm_rfft_fast_batch_f32(C.byref(S), hx.data_ptr(), hs.data_ptr(), Be, 0)
            b = L.arm_rfft_fast_batch_f32(C.byref(S), hs.data_ptr(), hy.data_ptr(), Be, 1)
            assert a == 0 and b == 0, cd.last_error()

        def timed_e2e(nrep):
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(nrep):
                e2e_step()
            barrier()
            return allmax((time.perf_counter() - t0) / nrep)

        nrep = max(1, min(args.steps, 3))
        dt = timed_e2e(nrep)
        ert = float(((hy[:64] - hx[:64]).double().pow(2).sum() / hx[:64].double().pow(2).sum()).sqrt())
        nbytes = Be * N_REAL * 4
        cp = pinned_copy_peak(torch, dev, nbytes, barrier, allmax)
        barrier()
        # per step and direction 2 * nbytes cross the link (forward + inverse call); the two directions overlap
        pcie = 2 * nbytes / dt / 1e9
        e2e = {"value": world * Be * N_REAL / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": 2 * nbytes, "d2h_bytes_per_step": 2 * nbytes,
               "frames": Be, "api": "arm_rfft_fast_batch_f32 (host pointers, pinned), forward then inverse; arm_cuda_set_devices([this rank's device])",
               "host_thread_bound_to_gpu_numa_node": numa_bound, "roundtrip_relrms": ert,
               "pcie_gbs_per_direction": pcie, "pinned_copy_peak": cp,
               "frac_of_pinned_copy_peak": pcie / cp["both_gbs_per_direction"],
               "limiter": "host link: the call moves 1 GiB each way per transform and overlaps copy-in, kernel and copy-out; "
                          "compare pcie_gbs_per_direction with pinned_copy_peak.both_gbs_per_direction (same bytes, same ranks active)",
               "staging": {"chunk_mib": int(os.environ.get("CMSISDSP_CUDA_CHUNK_MIB", "32")), "streams": int(os.environ.get("CMSISDSP_CUDA_NSTREAMS", "3"))}}
        if world > 1:
            # one process, all devices, through the C dispatcher: rank 0 alone drives every GPU of the box with ONE
            # batch of world * Be frames (block-partitioned inside the call); the other ranks wait at the barrier
            if rank == 0:
                try:
                    Ball = world * Be
                    gx = torch.empty(Ball, N_REAL, dtype=torch.float32).pin_memory()
                    gs = torch.empty_like(gx).pin_memory()
                    gx.normal_(generator=torch.Generator().manual_seed(5))
                    cd.set_devices(list(range(world)))
                    for k in range(3):
                        t0 = time.perf_counter()
                        assert L.arm_rfft_fast_batch_f32(C.byref(S), gx.data_ptr(), gs.data_ptr(), Ball, 0) == 0, cd.last_error()
                        dt1 = time.perf_counter() - t0
                    sub = torch.arange(0, Ball, Ball // 64)
                    err = _relrms(gs[sub].numpy(), oracle().rfft(N_REAL, gx[sub].numpy(), 0))
                    e2e["one_process_all_devices"] = {
                        "devices": world, "frames": Ball, "direction": "forward only", "seconds": dt1,
                        "value_msamples": Ball * N_REAL / dt1 / 1e6, "pcie_gbs_per_direction_total": Ball * N_REAL * 4 / dt1 / 1e9,
                        "forward_relrms_vs_oracle": err,
                        "api": "arm_rfft_fast_batch_f32, one call, arm_cuda_set_devices(all): one host thread per device inside the library"}
                    del gx, gs
                except Exception as e:
                    e2e["one_process_all_devices"] = {"error": repr(e)[:300]}
                cd.set_devices([local_rank])
            barrier()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = cpu_baseline_subprocess(FRAMES_PER_GPU // (4 if args.quick else 1), 3, 1)

    if rank == 0:
        if not parity["ok"]:
            print(json.dumps({"error": "parity check failed", "parity": parity}), file=sys.stderr)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG, "wall_s": t_wall,
            "roofline": roofline, "secondary": secondary, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "parity": parity, "roundtrip_relrms": parity["roundtrip_relrms_max_over_ranks"],
            "kernel_info": {"fwd": cd.kernel_info(3, N_REAL), "inv": cd.kernel_info(4, N_REAL)},
        }))
    if dist:
        dist.destroy_process_group()
    if not parity["ok"]:
        raise SystemExit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller e2e / cpu / secondary samples (profiling runs)")
    ap.add_argument("--cpu-baseline-only", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_only:
        print(json.dumps(cpu_reference_run(args.cpu_baseline_only, args.steps, args.warmup)))
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
