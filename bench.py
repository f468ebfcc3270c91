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

    def summary(self):
        sm = sorted(self.sm)
        mem = sorted(self.mem)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi",
                "mem_mhz": mem[len(mem) // 2] if mem else None, "power_w_max": max(self.power) if self.power else None}


def cpu_baseline_subprocess(frames):
    """cpu_baseline_run in a fresh interpreter: the threads of a process that has initialised
    torch/CUDA/NVML inherit a CPU affinity of one core on some boxes, which would make the
    16-thread CPU baseline read 16x too low."""
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only", str(frames)],
                         capture_output=True, text=True, timeout=900)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("cpu baseline subprocess failed: " + out.stderr[-400:])


def cpu_baseline_run(frames, repeats=3, fast=True):
    """The reference's generic-C arm_rfft_fast_f32 forward+inverse on all host cores."""
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
    x = rng.standard_normal((frames, N_REAL)).astype(np.float32)
    spec, y = np.empty_like(x), np.empty_like(x)
    fn = lib._fn("rfft_fast_f32_batch")
    best = float("inf")
    for _ in range(repeats + 1):                         # first pass is the warm-up
        xin = x.copy()                                   # forward destroys its input
        t0 = time.perf_counter()
        fn(N_REAL, xin.ctypes.data, spec.ctypes.data, frames, 0, cores)
        fn(N_REAL, spec.ctypes.data, y.ctypes.data, frames, 1, cores)
        dt = time.perf_counter() - t0
        best = min(best, dt) if _ else best
    err = float(np.sqrt(((y - x) ** 2).sum() / (x ** 2).sum()))
    return {"value": frames * N_REAL / best / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{frames} frames fwd+inv, best of {repeats}, {cores} pthreads, gcc -O3 generic-C build",
            "seconds": best, "roundtrip_relrms": err}


def run_reference(args, rank):
    if rank != 0:
        return
    frames = FRAMES_PER_GPU if (os.cpu_count() or 1) >= 16 else FRAMES_PER_GPU // 4
    vals = []
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_baseline_run(frames, repeats=1))
    timed = vals[args.warmup:]
    sec = sum(v["seconds"] for v in timed) / len(timed)
    value = frames * N_REAL / sec / 1e6
    cb = dict(timed[-1], value=value)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "frames_timed_per_step": frames},
        "cpu_baseline": cb, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_cuda(args, rank, world, local_rank):
    import numpy as np
    import torch
    import cmsisdsp_b200 as cd

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
        dist.init_process_group("nccl", device_id=dev)
    cu = cd.cuda()
    cu.cmsisdsp_cuda_set_device(local_rank)
    cd.ensure_rfft_plans(N_REAL)

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

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    # sanity: round trip and oracle parity on a few frames (outside the timed region)
    rt = float(((y[:64] - x[:64]).double().pow(2).sum() / x[:64].double().pow(2).sum()).sqrt())

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
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
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
                "algorithmic_bytes_per_launch": bytes_per_launch}
    # the same number of back-to-back launches of a plain device copy of the same size, timed the same way: what the
    # memory system sustains over a region this long (MEASURED_PEAKS.json's figure is a best-of-10 burst)
    try:
        ca, cb = x.view(-1), y.view(-1)
        for _ in range(3):
            cb.copy_(ca)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2 * args.steps):
            cb.copy_(ca)
        c1.record()
        torch.cuda.synchronize()
        copy_gbs = 2 * args.steps * 2 * ca.numel() * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        roofline["sustained_copy_gbs"] = copy_gbs
        roofline["frac_of_sustained_copy"] = achieved / copy_gbs
        roofline["sustained_copy_how"] = f"{2 * args.steps} back-to-back torch copy_ launches of 1 GiB f32 (read + write bytes), CUDA events"
    except Exception as e:                                # diagnostic only
        roofline["sustained_copy_gbs"] = None
        roofline["sustained_copy_how"] = f"failed: {e}"
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        with open(tr) as f:
            roofline["traffic"] = json.load(f).get(dom_name)

    # end to end through the public C API with pinned HOST buffers (H2D + kernels + D2H timed)
    e2e = None
    if not args.no_e2e:
        Be = B // 4 if args.quick else B
        S = cd.rfft_instance(N_REAL)
        hx = torch.empty(Be, N_REAL, dtype=torch.float32).pin_memory()
        hs = torch.empty_like(hx).pin_memory()
        hy = torch.empty_like(hx).pin_memory()
        hx.normal_(generator=torch.Generator().manual_seed(99 + rank))
        L = cd.lib()

        def e2e_step():
            a = L.arm_rfft_fast_batch_f32(C.byref(S), hx.data_ptr(), hs.data_ptr(), Be, 0)
            b = L.arm_rfft_fast_batch_f32(C.byref(S), hs.data_ptr(), hy.data_ptr(), Be, 1)
            assert a == 0 and b == 0, cd.last_error()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        nrep = max(1, min(args.steps, 3))
        for _ in range(nrep):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / nrep], device=dev, dtype=torch.float64)
        if dist:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ert = float(((hy[:64] - hx[:64]).double().pow(2).sum() / hx[:64].double().pow(2).sum()).sqrt())
        e2e = {"value": world * Be * N_REAL / float(dt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": 2 * Be * N_REAL * 4, "d2h_bytes_per_step": 2 * Be * N_REAL * 4,
               "frames": Be, "api": "arm_rfft_fast_batch_f32 (host pointers, pinned), forward then inverse",
               "host_thread_bound_to_gpu_numa_node": numa_bound,
               "roundtrip_relrms": ert}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = cpu_baseline_subprocess(FRAMES_PER_GPU // (4 if args.quick else 1))

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": B, "fft_len": N_REAL,
                       "l2_policy": "inputs larger than L2 (1 GiB per buffer per launch vs 126 MB L2)",
                       "timing": "CUDA events on the launch stream, max over ranks", "wall_s": t_wall},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "roundtrip_relrms": rt,
            "kernel_info": {"fwd": cd.kernel_info(3, N_REAL), "inv": cd.kernel_info(4, N_REAL)},
        }))
    if dist:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="smaller e2e / cpu samples (profiling runs)")
    ap.add_argument("--cpu-baseline-only", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_only:
        print(json.dumps(cpu_baseline_run(args.cpu_baseline_only)))
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
