"""The kernel bodies, compiled for the CPU (tests/emu/emu.cpp runs the very same
fft_*.cuh templates phase by phase), against the oracle: index math, twiddle indexing and
the bit-exact fixed-point arithmetic are checked here without a GPU.  Also asserts the
bank-conflict degree of every shared-memory exchange from the emulator's access trace."""
import ctypes as C
import os
import sys
import subprocess

import numpy as np
import pytest

import cmsisdsp_b200 as cd
from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, oracle, perm_from_swaps
from seeded_inputs import cfft_input, rfft_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "cmsis-dsp_b200", "csrc", "cuda")


@pytest.fixture(scope="module")
def emu():
    sys.path.insert(0, EMU_DIR)
    import build as emu_build
    so = emu_build.build()                     # parallel parts; only when a kernel header is newer than the library
    L = C.CDLL(so)
    L.emu_cfft.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.emu_rfft.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.emu_trace_stats.argtypes = [C.c_void_p, C.c_int]
    L.emu_cfft_mag.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    L.emu_rfft64.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
    L.emu_rfft_fix.argtypes = [C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def relrms(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


def product_tables(kind, N):
    """twiddles + permutation exactly as the front library would upload them"""
    tw, br = cd.instance_tables(cd.cfft_instance(kind, N), kind)
    return tw, perm_from_swaps(N, br).astype(np.uint16)


@pytest.mark.parametrize("kind", ["f32", "q31", "q15"])
@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_bodies(emu, kind, N):
    tw, perm = product_tables(kind, N)
    frames = 2 * {16: 128, 32: 64, 64: 32, 128: 16, 256: 8, 512: 4, 1024: 2}.get(N, 1) + 3   # ragged last CTA
    x = cfft_input(kind, N, frames=max(frames, 6), seed=N)
    for ifft in (0, 1):
        for bitrev in (0, 1):
            want = oracle().cfft(kind, N, x, ifft, bitrev)
            got = x.copy()
            assert emu.emu_cfft(cd.TYPE_ID[kind], N, got.ctypes.data, got.shape[0], ifft, bitrev, tw.ctypes.data, perm.ctypes.data) == 0
            if kind == "f32":
                assert relrms(got, want) <= 2e-6, (N, ifft, bitrev)       # north_star tolerance
                for f in range(got.shape[0]):
                    assert relrms(got[f], want[f]) <= 2e-6
            else:
                assert np.array_equal(got, want), (kind, N, ifft, bitrev)  # bit-exact


@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_f64_bodies(emu, N):
    """the f64 butterflies run in the reference's operation order: bit-identical to the oracle on the same table"""
    tw, perm = product_tables("f64", N)
    assert np.array_equal(tw, oracle().twiddle_f64(N))
    rng = np.random.default_rng(N)
    x = rng.standard_normal((2 * {16: 32, 32: 32, 64: 32, 128: 16, 256: 8, 512: 4, 1024: 2}.get(N, 1) + 3, 2 * N))
    for ifft in (0, 1):
        for bitrev in (0, 1):
            want = oracle().cfft_f64(N, x, ifft, bitrev)
            got = x.copy()
            assert emu.emu_cfft(cd.TYPE_ID["f64"], N, got.ctypes.data, got.shape[0], ifft, bitrev, tw.ctypes.data, perm.ctypes.data) == 0
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (N, ifft, bitrev)


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_bodies(emu, N):
    tw, _ = cd.instance_tables(cd.cfft_instance("f32", N // 2), "f32")
    S = cd.rfft_instance(N)
    twr = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
    x = rfft_input(N, frames=67, seed=N)
    for ifft in (0, 1):
        want = oracle().rfft(N, x, ifft)
        got = np.zeros_like(x)
        xin = x.copy()
        assert emu.emu_rfft(N, xin.ctypes.data, got.ctypes.data, x.shape[0], ifft, tw.ctypes.data, twr.ctypes.data) == 0
        assert np.array_equal(xin, x)                       # input left untouched
        assert relrms(got, want) <= 2e-6, (N, ifft)


def _stats(emu):
    buf = np.zeros((64, 6), dtype=np.int64)
    n = emu.emu_trace_stats(buf.ctypes.data, 64)
    return buf[:n]


def _trace(emu, run):
    emu.emu_trace_begin()
    run()
    return _stats(emu)


@pytest.mark.parametrize("kind,limit", [("f32", 1.0), ("q31", 1.0), ("q15", 1.0)])
@pytest.mark.parametrize("N", [128, 256, 512, 1024, 2048, 4096])
def test_exchange_bank_conflicts(emu, kind, N, limit):
    """wavefronts / ideal wavefronts of every exchange load/store: all three types exchange
    8-byte elements (q15 travels sign-extended) and must be conflict-free for N >= 128 (N = 128, eight threads per
    frame: the fixed-point plan pads after every 8 elements -- with 16 the first pass's stores of the two frames of a
    half-warp met in the same banks, ncu: 17 % excess wavefronts in the real FFT of N = 256, which is bound by that pipe)."""
    tw, _ = product_tables(kind, N)
    y = np.zeros((8, 2 * N), dtype=cd.NP_DTYPE[kind])
    rows = _trace(emu, lambda: emu.emu_cfft(cd.TYPE_ID[kind], N, y.ctypes.data, 8, 0, 1, tw.ctypes.data, None))
    assert len(rows) > 0
    for ph, st, nbytes, req, ideal, wf in rows:
        assert wf <= limit * ideal, (kind, N, int(ph), "store" if st else "load", wf / ideal)


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_fast_f64_bodies(emu, N):
    """fused f64 real FFT (CFFT + split epilogue / merge prologue): bit-identical to the oracle, input untouched,
    every exchange (the natural-order one of the split stage included) conflict-free"""
    tw, _ = cd.instance_tables(cd.cfft_instance("f64", N // 2), "f64")
    S = cd.rfft_f64_instance(N)
    twr = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
    x = np.random.default_rng(N).standard_normal((67, N))
    for ifft in (0, 1):
        want = oracle().rfft_f64(N, x, ifft)
        got, xin = np.zeros_like(x), x.copy()
        rows = _trace(emu, lambda: emu.emu_rfft64(N, xin.ctypes.data, got.ctypes.data, x.shape[0], ifft, tw.ctypes.data, twr.ctypes.data))
        assert np.array_equal(xin, x)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (N, ifft)
        for ph, st, nbytes, req, ideal, wf in rows:
            assert wf == ideal, (N, ifft, int(ph), "store" if st else "load", wf / ideal)


@pytest.mark.parametrize("N", LENGTHS)
def test_f64_exchange_bank_conflicts(emu, N):
    """16-byte exchange elements (a quarter warp per wavefront): every f64 plan is conflict-free"""
    tw, _ = product_tables("f64", N)
    y = np.zeros((64, 2 * N))
    rows = _trace(emu, lambda: emu.emu_cfft(cd.TYPE_ID["f64"], N, y.ctypes.data, 64, 0, 1, tw.ctypes.data, None))
    assert len(rows) > 0
    for ph, st, nbytes, req, ideal, wf in rows:
        assert nbytes == 16 and wf == ideal, (N, int(ph), "store" if st else "load", wf / ideal)


@pytest.mark.parametrize("N", [256, 512, 1024, 2048, 4096])
@pytest.mark.parametrize("ifft", [0, 1])
def test_rfft_exchange_bank_conflicts(emu, N, ifft):
    """same for the rfft plans (mirror passes read / write the exchange in reversed lane order);
    the scratch traffic of the 2R special bins (a few scattered 8-byte accesses) is exempt"""
    tw, _ = cd.instance_tables(cd.cfft_instance("f32", N // 2), "f32")
    S = cd.rfft_instance(N)
    twr = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
    x = np.zeros((8, N), dtype=np.float32)
    y = np.zeros_like(x)
    rows = _trace(emu, lambda: emu.emu_rfft(N, x.ctypes.data, y.ctypes.data, 8, ifft, tw.ctypes.data, twr.ctypes.data))
    big = [r for r in rows if r[3] >= 8]          # the exchanges proper: many requests per phase
    assert len(big) >= 2
    for ph, st, nbytes, req, ideal, wf in big:
        if (ifft == 0 and ph == 1 and st) or (ifft == 0 and ph == 2) or (ifft == 1 and ph == 0) or (ifft == 1 and ph == 1 and not st):
            continue                              # phases that touch the scratch area
        assert wf <= 1.0 * ideal, (N, ifft, int(ph), "store" if st else "load", wf / ideal)


@pytest.mark.parametrize("kind", ["q31", "q15"])
@pytest.mark.parametrize("N", RFIX_LENGTHS)
def test_rfft_fixed_point_bodies(emu, kind, N):
    """RfftFixFwdBody (CFFT + split stage fused) and CfftBody<.., RIFFT> (merge stage fused into the load, final
    saturating << 1) against the oracle, bit for bit, with the product's own tables; ragged last CTA; full-scale
    frames (wrap-around sums, saturation); no bank conflicts in the extra exchange of the forward body"""
    S = cd.rfft_fix_instance(kind, N)
    A = np.ctypeslib.as_array(S.pTwiddleAReal, shape=(8192,))
    B = np.ctypeslib.as_array(S.pTwiddleBReal, shape=(8192,))
    tw, _ = cd.instance_tables(S.pCfft.contents, kind)
    frames = 2 * {16: 128, 32: 64, 64: 32, 128: 16, 256: 8, 512: 4, 1024: 2}.get(N // 2, 1) + 3
    x = cfft_input(kind, N // 2, frames=max(frames, 6), seed=7 * N)                 # [frames, N] real samples
    want = oracle().rfft_fix(kind, N, x, 0, 1)
    got = np.zeros_like(want)
    emu.emu_trace_begin()
    assert emu.emu_rfft_fix(cd.TYPE_ID[kind], N, x.ctypes.data, got.ctypes.data, x.shape[0], 0, tw.ctypes.data,
                            A.ctypes.data, B.ctypes.data) == 0
    assert np.array_equal(got, want), (kind, N, "forward")
    rows = np.zeros((64, 6), dtype=np.int64)
    n = emu.emu_trace_stats(rows.ctypes.data, 64)
    first_new = 2 if N // 2 <= 256 else 4          # the phases the fused split adds behind the CFFT's own (2- / 3-pass plans)
    for ph, st, nbytes, req, ideal, wf in rows[:n]:
        if (N // 2 >= 256 and ph >= first_new) or N // 2 == 128:
            assert wf == ideal, f"{kind} N={N} phase {ph} {'store' if st else 'load'}: {wf} wavefronts for {ideal} ideal"
    spec = np.concatenate([want, cfft_input(kind, N, frames=6, seed=11 * N)])       # genuine spectra + arbitrary bins
    want = oracle().rfft_fix(kind, N, spec, 1, 1)
    got = np.zeros_like(want)
    assert emu.emu_rfft_fix(cd.TYPE_ID[kind], N, spec.ctypes.data, got.ctypes.data, spec.shape[0], 1, tw.ctypes.data,
                            A.ctypes.data, B.ctypes.data) == 0
    assert np.array_equal(got, want), (kind, N, "inverse")


@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_magnitude_epilogue_body(emu, N):
    """CfftMagBody (magnitude / squared magnitude fused behind the last pass) against the oracle's
    arm_cfft_f32 + arm_cmplx_mag[_squared]_f32, north_star tolerance"""
    tw, _ = product_tables("f32", N)
    frames = 2 * {16: 128, 32: 64, 64: 32, 128: 16, 256: 8, 512: 4, 1024: 2}.get(N, 1) + 3
    x = cfft_input("f32", N, frames=frames, seed=13 * N)
    for ifft in (0, 1):
        for sq in (0, 1):
            want = oracle().cfft_mag(N, x, ifft, bool(sq))
            got = np.zeros_like(want)
            xin = x.copy()
            assert emu.emu_cfft_mag(N, xin.ctypes.data, got.ctypes.data, frames, ifft, sq, tw.ctypes.data) == 0
            assert np.array_equal(xin, x)
            assert relrms(got, want) <= (4e-6 if sq else 2e-6), (N, ifft, sq)

def test_every_plan_every_exchange_is_conflict_free(emu):
    """the whole plan table at once: every traced shared-memory exchange of every (type, length, direction) -- complex
    f32 / q31 / q15 / f64, real f32, real q31 / q15 -- costs exactly its ideal number of wavefronts (the plans with 8 threads
    per frame were outside the per-family assertions above until ncu showed their conflicts on the GPU)"""
    frames = 256

    def bad_rows():
        return [(int(ph), "store" if st else "load", int(ideal), int(wf)) for ph, st, nb, req, ideal, wf in _stats(emu) if wf > ideal]

    for kind in ("f32", "q31", "q15", "f64"):
        for N in LENGTHS:
            tw, _ = product_tables(kind, N)
            y = np.zeros((frames, 2 * N), dtype=cd.NP_DTYPE[kind])
            emu.emu_trace_begin()
            emu.emu_cfft(cd.TYPE_ID[kind], N, y.ctypes.data, frames, 0, 1, tw.ctypes.data, None)
            assert not bad_rows(), ("cfft", kind, N, bad_rows())
    for N in RLENGTHS:
        tw, _ = cd.instance_tables(cd.cfft_instance("f32", N // 2), "f32")
        S = cd.rfft_instance(N)
        twr = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
        for ifft in (0, 1):
            x = np.zeros((frames, N), dtype=np.float32)
            y = np.zeros_like(x)
            emu.emu_trace_begin()
            emu.emu_rfft(N, x.ctypes.data, y.ctypes.data, frames, ifft, tw.ctypes.data, twr.ctypes.data)
            assert not bad_rows(), ("rfft f32", N, ifft, bad_rows())       # (the special-bin scratch accesses are not part of the trace)
    for kind in ("q31", "q15"):
        for N in RFIX_LENGTHS:
            S = cd.rfft_fix_instance(kind, N)
            A = np.ctypeslib.as_array(S.pTwiddleAReal, shape=(8192,))
            B = np.ctypeslib.as_array(S.pTwiddleBReal, shape=(8192,))
            tw, _ = cd.instance_tables(S.pCfft.contents, kind)
            for ifft in (0, 1):
                x = np.zeros((frames, 2 * N if ifft else N), dtype=cd.NP_DTYPE[kind])
                got = np.zeros((frames, N if ifft else 2 * N), dtype=x.dtype)
                emu.emu_trace_begin()
                assert emu.emu_rfft_fix(cd.TYPE_ID[kind], N, x.ctypes.data, got.ctypes.data, frames, ifft, tw.ctypes.data,
                                        A.ctypes.data, B.ctypes.data) == 0
                assert not bad_rows(), ("rfft", kind, N, ifft, bad_rows())
