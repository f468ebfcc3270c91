"""Deterministic synthetic inputs shared by the golden-digest generator, the
oracle tests and the GPU parity tests (numpy PCG64 streams keyed by seed)."""
import numpy as np


def cfft_input(kind, N, frames, seed):
    """[frames, 2N] interleaved (re, im). Integer kinds mix full-range uniform,
    clipped gaussian and hand-built extremes."""
    rng = np.random.default_rng([seed, {"f32": 1, "q31": 2, "q15": 3}[kind]])
    if kind == "f32":
        return rng.standard_normal((frames, 2 * N)).astype(np.float32)
    dt = np.int32 if kind == "q31" else np.int16
    info = np.iinfo(dt)
    x = rng.integers(info.min, info.max, size=(frames, 2 * N), endpoint=True).astype(dt)
    if frames >= 6:
        x[0] = info.min
        x[1] = info.max
        x[2, 0::2] = info.min
        x[2, 1::2] = info.max
        x[3] = np.where(np.arange(2 * N) % 4 < 2, info.max, info.min).astype(dt)
        x[4] = (rng.standard_normal(2 * N) * 0.25 * info.max).clip(info.min, info.max).astype(dt)
    return x


def rfft_input(N, frames, seed):
    rng = np.random.default_rng([seed, 4])
    return rng.standard_normal((frames, N)).astype(np.float32)
