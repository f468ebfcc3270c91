"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C ABI
(libcmsisdsp_b200.so front library and libcmsisdsp_cuda.so shim), against the CPU oracle on
the same seeded inputs, against the reference's golden vectors, and -- at BASELINE.json's
full sizes -- through size-independent properties.

Bar (BASELINE.json north_star): q15/q31 bit-exact; f32 relative RMS error <= 2e-6.
"""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import cmsisdsp_b200 as cd
from golden_checks import (assert_like_reference, assert_rfft_fix_like_reference, golden_cases, golden_rfft_fix_cases,
                           ref_digests, rfft_fix_threshold_applies)
from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, oracle
from seeded_inputs import cfft_input, rfft_input

pytestmark = pytest.mark.gpu
F32_TOL = 2e-6          # relative RMS, north_star
NT = min(32, os.cpu_count() or 1)
HERE = os.path.dirname(os.path.abspath(__file__))


def relrms(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.fixture(scope="module", autouse=True)
def device():
    cu = cd.cuda()
    assert cu.cmsisdsp_cuda_device_count() >= 1, "no CUDA device: these tests must run on a GPU box"
    assert cu.cmsisdsp_cuda_set_device(0) == 0
    return cu


# ------------------------------------------------------------------ every (type, N, direction, order)

@pytest.mark.parametrize("kind", ["f32", "q31", "q15"])
@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_all_modes_host_buffers(kind, N):
    x = cfft_input(kind, N, frames=301, seed=N)          # ragged: not a multiple of any frames-per-CTA
    for ifft in (0, 1):
        for bitrev in (0, 1):
            want = oracle().cfft(kind, N, x, ifft, bitrev, threads=NT)
            got = cd.cfft_batch(kind, N, x, ifft, bitrev)
            if kind == "f32":
                assert relrms(got, want) <= F32_TOL, (N, ifft, bitrev)
                per_frame = np.sqrt(((got.astype(np.float64) - want) ** 2).sum(1) / (want.astype(np.float64) ** 2).sum(1))
                assert per_frame.max() <= F32_TOL, (N, ifft, bitrev, per_frame.max())
            else:
                assert np.array_equal(got, want), (kind, N, ifft, bitrev)


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_both_directions(N):
    x = rfft_input(N, frames=203, seed=N)
    xin = x.copy()
    spec = cd.rfft_batch(N, xin, 0)
    assert np.array_equal(xin, x)                          # the batched call leaves p untouched
    assert relrms(spec, oracle().rfft(N, x, 0, threads=NT)) <= F32_TOL
    want_spec = oracle().rfft(N, x, 0, threads=NT)
    back = cd.rfft_batch(N, want_spec, 1)
    assert relrms(back, oracle().rfft(N, want_spec, 1, threads=NT)) <= F32_TOL
    assert relrms(back, x) <= F32_TOL                      # round trip


def test_digests_of_compiled_reference_fixed_point():
    """bit-level pin against outputs of the compiled reference (tests/golden/ref_digests.json)"""
    dig = ref_digests()
    for kind in ("q31", "q15"):
        for N in LENGTHS:
            x = cfft_input(kind, N, frames=8, seed=N)
            for ifft in (0, 1):
                for bitrev in (0, 1):
                    y = cd.cfft_batch(kind, N, x, ifft, bitrev)
                    assert hashlib.sha256(y.tobytes()).hexdigest() == dig[f"cfft_{kind}/{N}/{ifft}/{bitrev}"], (kind, N, ifft, bitrev)


# ------------------------------------------------------------------ the reference's own tests

@pytest.mark.parametrize("kind", ["f32", "q31", "q15"])
def test_reference_cfft_patterns(kind):
    n = 0
    for N, sig, ifft, x, ref in golden_cases(kind, "c"):
        out = cd.cfft_batch(kind, N, x, ifft, 1).reshape(-1)
        assert_like_reference(kind, "c", out, ref, N, ifft)
        n += 1
    assert n == 36


def test_reference_rfft_patterns():
    n = 0
    for N, sig, ifft, x, ref in golden_cases("f32", "r"):
        out = cd.rfft_batch(N, x, ifft).reshape(-1)
        assert_like_reference("f32", "r", out, ref, N, ifft)
        n += 1
    assert n == 32


def test_fft_bin_example_known_answer():
    d = np.load(os.path.join(HERE, "golden", "fft_bin_example.npz"))
    S = cd.preset("f32", 1024)                               # arm_cfft_sR_f32_len1024, as the example uses
    buf = d["input"].copy()
    cd.lib().arm_cfft_f32(C.byref(S), buf.ctypes.data, 0, 1)
    assert cd.lib().arm_cuda_last_status() == 0
    mag = np.hypot(*buf.reshape(-1, 2).astype(np.float64).T)
    # the input is real, so bins 213 and 1024-213 carry the same magnitude up to rounding;
    # the example's arm_max_f32 reports the first one (arm_fft_bin_example_f32.c:145-155)
    assert int(np.argmax(mag[:512])) == 213
    assert int(np.argmax(mag)) in (213, 1024 - 213)
    assert abs(mag[213] - mag[811]) <= 1e-5 * mag[213]


# ------------------------------------------------------------------ API behaviour

# ------------------------------------------------------------------ arm_cfft_f64
F64_TOL = 1e-15         # relative RMS against the compiled reference (its twiddle literals differ from ours by <= 1 ulp)


@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_f64_all_modes(N):
    """Bit-identical to the oracle (same operation order, no FMA contraction, same table); relative RMS <= 1e-15 per
    frame against the reference's own build, whose table differs in the last place."""
    from oracle_lib import ref
    rng = np.random.default_rng(7000 + N)
    x = rng.standard_normal((301, 2 * N))
    x[1] *= 1e-6
    x[2] *= 1e9
    for ifft in (0, 1):
        for bitrev in (0, 1):
            want = oracle().cfft_f64(N, x, ifft, bitrev, threads=NT)
            got = cd.cfft_batch("f64", N, x, ifft, bitrev)
            assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), (N, ifft, bitrev)
            if ref() is not None:
                r = ref().cfft_f64(N, x[:40], ifft, bitrev)
                per_frame = np.sqrt(((got[:40] - r) ** 2).sum(1) / (r ** 2).sum(1))
                assert per_frame.max() <= F64_TOL, (N, ifft, bitrev, per_frame.max())


def test_cfft_f64_reference_patterns_legacy_call_and_device_pointers():
    torch = pytest.importorskip("torch")
    L = cd.lib()
    n = 0
    for N, sig, ifft, x, refv in golden_cases("f64", "c"):          # TransformCF64.cpp thresholds (SNR 250 dB)
        buf = np.ascontiguousarray(x, dtype=np.float64).copy()
        S = cd.cfft_instance("f64", N)
        L.arm_cfft_f64(C.byref(S), buf.ctypes.data, ifft, 1)         # the reference's single-frame signature
        assert L.arm_cuda_last_status() == 0
        assert_like_reference("f64", "c", buf, refv, N, ifft)
        n += 1
    assert n == 36
    dev = torch.device("cuda", 0)
    for N, frames in ((16, 100003), (64, 4099), (1024, 2049), (4096, 1025)):
        x = np.random.default_rng(N).standard_normal((frames, 2 * N))
        cd.ensure_plans("f64", N)
        t = torch.from_numpy(x).to(dev)
        cd.cfft_device("f64", N, t.data_ptr(), frames, 0, 1, torch.cuda.current_stream().cuda_stream)
        spec = t.cpu().numpy()
        z = np.fft.fft(x[:, 0::2] + 1j * x[:, 1::2], axis=1)         # an independent DFT
        zz = np.stack([z.real, z.imag], axis=2).reshape(frames, 2 * N)
        assert relrms(spec, zz) <= 1e-15 * np.log2(N)
        sub = slice(0, 64)
        assert np.array_equal(spec[sub], oracle().cfft_f64(N, x[sub], 0, 1))
        cd.cfft_device("f64", N, t.data_ptr(), frames, 1, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert relrms(t.cpu().numpy(), x) <= 1e-15 * np.log2(N)      # round trip
        # misaligned device pointer: refused, not mangled
        assert cd.cuda().cmsisdsp_cuda_cfft_f64(t.data_ptr() + 8, N, 1, 0, 1, 0) == -1
    S = cd.cfft_instance("f64", 64)
    assert L.arm_cfft_batch_f64(None, None, 1, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    assert L.arm_cfft_batch_f64(C.byref(S), None, 1, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    one = np.zeros(128)
    assert L.arm_cfft_batch_f64(C.byref(S), one.ctypes.data, 0, 0, 1) == 0


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_fast_f64_both_directions(N):
    """arm_rfft_fast_f64 (adapter: complex kernel + split / merge stage): bit-identical to the oracle on the product's
    tables, <= 1e-15 against the compiled reference, the reference's golden vectors, the legacy call's side effect."""
    from oracle_lib import ref
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8000 + N)
    x = rng.standard_normal((301, N))
    want = oracle().rfft_f64(N, x, 0)
    spec = cd.rfft_f64_batch(N, x, 0)
    assert np.array_equal(spec.view(np.uint64), want.view(np.uint64))
    back = cd.rfft_f64_batch(N, spec, 1)
    assert np.array_equal(back.view(np.uint64), oracle().rfft_f64(N, spec, 1).view(np.uint64))
    assert relrms(back, x) <= 1e-15 * np.log2(N)
    if ref() is not None:
        assert relrms(spec[:40], ref().rfft_f64(N, x[:40], 0)) <= F64_TOL
        assert relrms(back[:40], ref().rfft_f64(N, spec[:40], 1)) <= F64_TOL
    # numpy's rfft as an independent check of the packing {DC, Nyquist, Re1, Im1, ...}
    z = np.fft.rfft(x, axis=1)
    packed = np.empty_like(x)
    packed[:, 0], packed[:, 1] = z[:, 0].real, z[:, N // 2].real
    packed[:, 2::2], packed[:, 3::2] = z[:, 1:N // 2].real, z[:, 1:N // 2].imag
    assert relrms(spec, packed) <= 1e-15 * np.log2(N)
    # the reference's single-frame signature: forward leaves the N/2-point CFFT in p, inverse leaves p alone
    L = cd.lib()
    S = cd.rfft_f64_instance(N)
    p, out = x[:1].copy(), np.zeros((1, N))
    L.arm_rfft_fast_f64(C.byref(S), p.ctypes.data, out.ctypes.data, 0)
    assert L.arm_cuda_last_status() == 0
    assert np.array_equal(out, want[:1]) and np.array_equal(p, oracle().cfft_f64(N // 2, x[:1], 0, 1))
    p2, out2 = want[:1].copy(), np.zeros((1, N))
    L.arm_rfft_fast_f64(C.byref(S), p2.ctypes.data, out2.ctypes.data, 1)
    assert np.array_equal(p2, want[:1]) and relrms(out2, x[:1]) <= 1e-15 * np.log2(N)
    # device pointers: source left untouched
    dev = torch.device("cuda", 0)
    cd.ensure_rfft_f64_plans(N)
    a = torch.from_numpy(x).to(dev)
    b = torch.empty_like(a)
    cd.rfft_f64_device(N, a.data_ptr(), b.data_ptr(), x.shape[0], 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(b.cpu().numpy(), spec) and np.array_equal(a.cpu().numpy(), x)
    assert L.arm_rfft_fast_batch_f64(C.byref(S), a.data_ptr(), a.data_ptr(), 1, 0) == cd.ARM_MATH_ARGUMENT_ERROR    # aliasing


def test_rfft_fast_f64_reference_patterns():
    n = 0
    for N, sig, ifft, x, refv in golden_cases("f64", "r"):           # TransformRF64.cpp thresholds
        out = cd.rfft_f64_batch(N, x, ifft).reshape(-1)
        assert_like_reference("f64", "r", out, refv, N, ifft)
        n += 1
    assert n == 32


def test_legacy_single_frame_signatures():
    L = cd.lib()
    for kind in ("f32", "q31", "q15"):
        x = cfft_input(kind, 256, frames=1, seed=5)
        buf = x.copy()
        S = cd.cfft_instance(kind, 256)
        getattr(L, f"arm_cfft_{kind}")(C.byref(S), buf.ctypes.data, 0, 1)
        assert L.arm_cuda_last_status() == 0
        want = oracle().cfft(kind, 256, x, 0, 1)
        assert relrms(buf, want) <= F32_TOL if kind == "f32" else np.array_equal(buf, want)
    # rfft forward: pOut = spectrum AND p is left holding the N/2-point CFFT, like the reference
    x = rfft_input(512, frames=1, seed=6)
    p, out = x.copy(), np.zeros_like(x)
    S = cd.rfft_instance(512)
    L.arm_rfft_fast_f32(C.byref(S), p.ctypes.data, out.ctypes.data, 0)
    assert L.arm_cuda_last_status() == 0
    want_out, want_p = oracle().rfft(512, x, 0, return_clobbered=True)
    assert relrms(out, want_out) <= F32_TOL and relrms(p, want_p) <= F32_TOL
    # inverse: p untouched
    p2, back = want_out.copy(), np.zeros_like(x)
    L.arm_rfft_fast_f32(C.byref(S), p2.ctypes.data, back.ctypes.data, 1)
    assert np.array_equal(p2, want_out) and relrms(back, x) <= F32_TOL


def test_edge_cases():
    L = cd.lib()
    S = cd.cfft_instance("f32", 64)
    x = cfft_input("f32", 64, frames=3, seed=1)
    buf = x.copy()
    assert L.arm_cfft_batch_f32(C.byref(S), buf.ctypes.data, 0, 0, 1) == 0          # empty batch: no-op
    assert np.array_equal(buf, x)
    assert L.arm_cfft_batch_f32(None, buf.ctypes.data, 3, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    assert L.arm_cfft_batch_f32(C.byref(S), None, 3, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    bad = cd.arm_cfft_instance_f32()
    bad.fftLen, bad.pTwiddle, bad.pBitRevTable, bad.bitRevLength = 24, S.pTwiddle, S.pBitRevTable, S.bitRevLength
    assert L.arm_cfft_batch_f32(C.byref(bad), buf.ctypes.data, 3, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    L.arm_cfft_f32(C.byref(bad), buf.ctypes.data, 0, 1)                             # unsupported length: silent no-op
    assert np.array_equal(buf, x)
    R = cd.rfft_instance(64)
    r = rfft_input(64, frames=2, seed=2)
    assert L.arm_rfft_fast_batch_f32(C.byref(R), r.ctypes.data, r.ctypes.data, 2, 0) == cd.ARM_MATH_ARGUMENT_ERROR   # aliasing
    for frames in (1, 2, 31, 33, 127, 129):                                          # every tail shape of the CTA packing
        for kind, N in (("f32", 16), ("q15", 32), ("q31", 128), ("f32", 1024)):
            xi = cfft_input(kind, N, frames=frames, seed=frames)
            got, want = cd.cfft_batch(kind, N, xi), oracle().cfft(kind, N, xi)
            assert relrms(got, want) <= F32_TOL if kind == "f32" else np.array_equal(got, want)


@pytest.mark.parametrize("N", [64, 1024, 4096])
def test_f32_extreme_amplitudes_and_denormals(N):
    """arm_cfft_f32 / arm_rfft_fast_f32 far from unit scale: 1e30 and 1e-30 ... 1e-36 keep the 2e-6 bar; at 1e-41 inputs,
    intermediates and results are denormal -- the kernels compute with gradual underflow like the reference's generic-C build
    on the host (no flush to zero: the same values are exactly zero in both), within what 7-bit significands allow"""
    rng = np.random.default_rng(N)
    for scale in (1e30, 1e-30, 1e-36):
        x = (rng.standard_normal((6, 2 * N)) * scale).astype(np.float32)
        r = (rng.standard_normal((6, N)) * scale).astype(np.float32)
        for ifft in (0, 1):
            assert relrms(cd.cfft_batch("f32", N, x, ifft, 1), oracle().cfft("f32", N, x, ifft, 1)) <= F32_TOL, (scale, ifft)
        spec = oracle().rfft(N, r, 0)
        assert relrms(cd.rfft_batch(N, r, 0), spec) <= F32_TOL, scale
        assert relrms(cd.rfft_batch(N, spec, 1), oracle().rfft(N, spec, 1)) <= F32_TOL, scale
    x = (rng.standard_normal((6, 2 * N)) * 1e-41).astype(np.float32)
    assert np.count_nonzero(x) > N and np.abs(x).max() < 1.2e-38          # denormal inputs
    for ifft in (0, 1):
        got, want = cd.cfft_batch("f32", N, x, ifft, 1), oracle().cfft("f32", N, x, ifft, 1)
        assert np.count_nonzero(want) > 0 and relrms(got, want) <= 2e-3, ifft
        assert np.count_nonzero(got) >= 0.9 * np.count_nonzero(want), ifft   # not flushed


@pytest.mark.parametrize("N", [64, 1024, 4096])
def test_non_finite_frames_stay_in_their_frame(N):
    """a frame with a NaN or an Inf comes out non-finite -- and its neighbours in the batch, which share shared-memory
    buffers, CTAs and warps with it, do not change by a bit"""
    rng = np.random.default_rng(3 * N)
    frames, bad = 37, (5, 20, 36)
    x = rng.standard_normal((frames, 2 * N)).astype(np.float32)
    r = rng.standard_normal((frames, N)).astype(np.float32)
    xp, rp = x.copy(), r.copy()
    xp[5, 7], xp[20, 2 * N - 1], xp[36, 0] = np.nan, np.inf, -np.inf
    rp[5, 7], rp[20, N - 1], rp[36, 0] = np.nan, np.inf, -np.inf
    good = np.array([k for k in range(frames) if k not in bad])
    for ifft in (0, 1):
        y0, y1 = cd.cfft_batch("f32", N, x, ifft, 1), cd.cfft_batch("f32", N, xp, ifft, 1)
        assert np.array_equal(y0[good], y1[good]), ("cfft", ifft)
        for k in bad:                                        # (X[0] sums the real and the imaginary parts separately, trivial
            assert np.mean(~np.isfinite(y1[k])) >= 0.45, ("cfft", ifft, k)     # twiddles are not multiplied: not EVERY output is hit)
        z0, z1 = cd.rfft_batch(N, r, ifft), cd.rfft_batch(N, rp, ifft)
        assert np.array_equal(z0[good], z1[good]), ("rfft", ifft)
        for k in bad:
            assert np.mean(~np.isfinite(z1[k])) >= 0.45, ("rfft", ifft, k)
    if N == 1024:
        from oracle_lib import mfcc_config
        m = cd.Mfcc(mfcc_config(N))
        m0, m1 = m.batch(r.reshape(-1), hop=N), m.batch(rp.reshape(-1), hop=N)
        assert np.array_equal(m0[good], m1[good])
        assert np.any(~np.isfinite(m1[5]))


def test_device_pointer_path_matches_host_path():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    for kind, N in (("f32", 1024), ("q31", 512), ("q15", 2048)):
        x = cfft_input(kind, N, frames=77, seed=3)
        cd.ensure_plans(kind, N)
        t = torch.from_numpy(x).to(dev)
        cd.cfft_device(kind, N, t.data_ptr(), 77, 0, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy(), cd.cfft_batch(kind, N, x, 0, 1))
        # front library accepts device pointers too
        t2 = torch.from_numpy(x).to(dev)
        S = cd.cfft_instance(kind, N)
        assert getattr(cd.lib(), f"arm_cfft_batch_{kind}")(C.byref(S), t2.data_ptr(), 77, 0, 1) == 0
        assert np.array_equal(t2.cpu().numpy(), t.cpu().numpy())


# ------------------------------------------------------------------ BASELINE.json full sizes

def test_config1_cfft_f32_1024_x4096_full_compare():
    x = cfft_input("f32", 1024, frames=4096, seed=1)
    want = oracle().cfft("f32", 1024, x, 0, 1, threads=NT)
    got = cd.cfft_batch("f32", 1024, x, 0, 1)
    assert relrms(got, want) <= F32_TOL
    per_frame = np.sqrt(((got.astype(np.float64) - want) ** 2).sum(1) / (want.astype(np.float64) ** 2).sum(1))
    assert per_frame.max() <= F32_TOL


def test_config2_rfft_4096_x65536_properties():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    N, B = 4096, 65536
    cd.ensure_rfft_plans(N)
    st = torch.cuda.current_stream().cuda_stream
    x = torch.randn(B, N, device=dev, generator=torch.Generator(device=dev).manual_seed(7))
    spec, y = torch.empty_like(x), torch.empty_like(x)
    cd.rfft_device(N, x.data_ptr(), spec.data_ptr(), B, 0, st)
    cd.rfft_device(N, spec.data_ptr(), y.data_ptr(), B, 1, st)
    torch.cuda.synchronize()
    rt = ((y - x).double().pow(2).sum() / x.double().pow(2).sum()).sqrt().item()
    assert rt <= F32_TOL                                           # forward -> inverse round trip over the full batch
    # Parseval on the packed spectrum: sum x^2 = (DC^2 + Nyq^2 + 2 sum |X_k|^2) / N
    e_t = x.double().pow(2).sum(1)
    s = spec.double()
    e_f = (s[:, 0] ** 2 + s[:, 1] ** 2 + 2 * s[:, 2:].pow(2).sum(1)) / N
    assert ((e_f - e_t).abs() / e_t).max().item() < 1e-5
    # linearity: F(a + 2b) = F(a) + 2 F(b) on a slice
    a, b = x[:256], x[256:512]
    lin = torch.empty_like(a)
    cd.rfft_device(N, (a + 2 * b).contiguous().data_ptr(), lin.data_ptr(), 256, 0, st)
    torch.cuda.synchronize()
    ref_lin = spec[:256] + 2 * spec[256:512]
    assert ((lin - ref_lin).double().pow(2).sum() / ref_lin.double().pow(2).sum()).sqrt().item() <= 2 * F32_TOL
    # oracle parity on a stratified subsample of 1024 frames
    idx = torch.arange(0, B, B // 1024, device=dev)
    xs = x[idx].cpu().numpy()
    assert relrms(spec[idx].cpu().numpy(), oracle().rfft(N, xs, 0, threads=NT)) <= F32_TOL
    assert relrms(y[idx].cpu().numpy(), oracle().rfft(N, spec[idx].cpu().numpy(), 1, threads=NT)) <= F32_TOL


@pytest.mark.parametrize("kind", ["q15", "q31"])
@pytest.mark.parametrize("N", [256, 1024, 4096])
def test_config3_fixed_point_1M_frames_bit_exact(kind, N):
    """2^20 frames built by tiling 4096 distinct seeded frames 256 times: the first tile must be
    memcmp-identical to the oracle and every other tile identical to the first (frames are
    independent, so this pins all 2^20 outputs bit for bit)."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    B, U = 1 << 20, 4096
    base = cfft_input(kind, N, frames=U, seed=100 + N)
    cd.ensure_plans(kind, N)
    st = torch.cuda.current_stream().cuda_stream
    for ifft, bitrev in ((0, 1), (1, 1), (0, 0)):
        want = oracle().cfft(kind, N, base, ifft, bitrev, threads=NT)
        t = torch.from_numpy(base).to(dev).repeat(B // U, 1)
        assert t.shape == (B, 2 * N)
        cd.cfft_device(kind, N, t.data_ptr(), B, ifft, bitrev, st)
        torch.cuda.synchronize()
        tiles = t.view(B // U, U, 2 * N)
        assert np.array_equal(tiles[0].cpu().numpy(), want), (kind, N, ifft, bitrev)
        assert bool((tiles == tiles[0:1]).all().item()), (kind, N, ifft, bitrev)
        del t, tiles
        torch.cuda.empty_cache()


def test_config5_length_sweep_f32_large_batches():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    for N in LENGTHS:
        B = (1 << 28) // (8 * N)                 # 256 MiB of complex f32 per length
        cd.ensure_plans("f32", N)
        x = torch.randn(B, 2 * N, device=dev, generator=torch.Generator(device=dev).manual_seed(N))
        y = x.clone()
        cd.cfft_device("f32", N, y.data_ptr(), B, 0, 1, st)
        fwd = y.clone()
        cd.cfft_device("f32", N, y.data_ptr(), B, 1, 1, st)
        torch.cuda.synchronize()
        rt = ((y - x).double().pow(2).sum() / x.double().pow(2).sum()).sqrt().item()
        assert rt <= F32_TOL, (N, rt)
        idx = torch.arange(0, B, max(1, B // 512), device=dev)
        want = oracle().cfft("f32", N, x[idx].cpu().numpy(), 0, 1, threads=NT)
        assert relrms(fwd[idx].cpu().numpy(), want) <= F32_TOL, N


# ------------------------------------------------------------------ kernel flavours

def test_direct_and_pipelined_flavours_agree_bit_for_bit():
    """The persistent TMA-fed kernels (two-pass f32 plans) run the same arithmetic as the direct
    kernels: outputs must be identical, for ragged batch sizes too (partial last group, fewer
    groups than resident CTAs)."""
    cu = cd.cuda()
    try:
        for frames in (1, 3, 257, 5000):
            for N in (512, 1024, 2048, 4096):
                x = cfft_input("f32", N, frames=frames, seed=N + frames)
                for ifft, bitrev in ((0, 1), (1, 0)):
                    out = []
                    for flavour in (0, 1):
                        assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
                        out.append(cd.cfft_batch("f32", N, x, ifft, bitrev))
                    assert np.array_equal(out[0], out[1]), (N, frames, ifft, bitrev)
                    assert relrms(out[1], oracle().cfft("f32", N, x, ifft, bitrev, threads=NT)) <= F32_TOL
            for kind in ("q31", "q15"):                      # fixed point: persistent TMA-fed flavour of the multi-pass plans
                for N in (128, 256, 512, 1024, 2048, 4096):
                    xi = cfft_input(kind, N, frames=max(frames, 6), seed=N + frames)[:max(frames, 1)] if frames >= 6 else cfft_input(kind, N, frames=6, seed=N)[:frames]
                    for ifft, bitrev in ((0, 1), (1, 0)):
                        want = oracle().cfft(kind, N, xi, ifft, bitrev, threads=NT)
                        for flavour in (0, 1):
                            assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
                            assert np.array_equal(cd.cfft_batch(kind, N, xi, ifft, bitrev), want), (kind, N, frames, ifft, bitrev, flavour)
            for N in (512, 1024, 2048, 4096):
                xr = rfft_input(N, frames=frames, seed=N + frames)
                for ifft in (0, 1):
                    src = xr if not ifft else oracle().rfft(N, xr, 0, threads=NT)
                    out = []
                    for flavour in (0, 1):
                        assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
                        out.append(cd.rfft_batch(N, src, ifft))
                    assert np.array_equal(out[0], out[1]), (N, frames, ifft)
                    assert relrms(out[1], oracle().rfft(N, src, ifft, threads=NT)) <= F32_TOL
    finally:
        cu.cmsisdsp_cuda_set_kernel_flavour(-1)
    assert cu.cmsisdsp_cuda_set_kernel_flavour(7) != 0


# ------------------------------------------------------------------ call-compatible Python surface

def test_python_wrapper_compatible_surface():
    """`import cmsisdsp_b200.compat as dsp` takes the call sequences written for the reference's PythonWrapper
    (cmsisdsp_transform.c:2074-2545): instance(), init(S, ...), exec(S, x, ...) -> array"""
    import cmsisdsp_b200.compat as dsp
    from oracle_lib import mfcc_config
    x = cfft_input("f32", 256, frames=3, seed=1)
    S = dsp.arm_cfft_instance_f32()
    assert dsp.arm_cfft_init_f32(S, 256) == 0 and dsp.arm_cfft_init_f32(dsp.arm_cfft_instance_f32(), 100) == cd.ARM_MATH_ARGUMENT_ERROR
    assert relrms(dsp.arm_cfft_f32(S, x[0], 0, 1), oracle().cfft("f32", 256, x[0], 0, 1).reshape(-1)) <= F32_TOL   # one frame
    assert relrms(dsp.arm_cfft_f32(S, x, 1, 1), oracle().cfft("f32", 256, x, 1, 1).reshape(-1)) <= F32_TOL    # three frames, one call
    for kind in ("q31", "q15"):
        Si = getattr(dsp, f"arm_cfft_instance_{kind}")()
        assert getattr(dsp, f"arm_cfft_init_{kind}")(Si, 1024) == 0
        xi = cfft_input(kind, 1024, frames=7, seed=2)
        assert np.array_equal(getattr(dsp, f"arm_cfft_{kind}")(Si, xi, 0, 1), oracle().cfft(kind, 1024, xi, 0, 1).reshape(-1))
        Sr = getattr(dsp, f"arm_rfft_instance_{kind}")()
        assert getattr(dsp, f"arm_rfft_init_{kind}")(Sr, 512, 0, 1) == 0
        xr = cfft_input(kind, 256, frames=6, seed=3)
        spec = getattr(dsp, f"arm_rfft_{kind}")(Sr, xr)
        assert np.array_equal(spec, oracle().rfft_fix(kind, 512, xr, 0, 1).reshape(-1))
        Sv = getattr(dsp, f"arm_rfft_instance_{kind}")()
        assert getattr(dsp, f"arm_rfft_init_{kind}")(Sv, 512, 1, 1) == 0
        one = spec[:2 * 512]
        assert np.array_equal(getattr(dsp, f"arm_rfft_{kind}")(Sv, one[:512 + 2]), oracle().rfft_fix(kind, 512, one, 1, 1).reshape(-1))
    R = dsp.arm_rfft_fast_instance_f32()
    assert dsp.arm_rfft_fast_init_f32(R, 1024) == 0
    xr = rfft_input(1024, frames=5, seed=4)
    assert relrms(dsp.arm_rfft_fast_f32(R, xr, 0), oracle().rfft(1024, xr, 0).reshape(-1)) <= F32_TOL
    D = dsp.arm_cfft_instance_f64()
    assert dsp.arm_cfft_init_f64(D, 256) == 0 and dsp.arm_cfft_init_f64(D, 100) == cd.ARM_MATH_ARGUMENT_ERROR
    assert dsp.arm_cfft_init_f64(D, 256) == 0
    xd = np.random.default_rng(9).standard_normal(3 * 512)
    assert np.array_equal(dsp.arm_cfft_f64(D, xd, 0, 1), oracle().cfft_f64(256, xd, 0, 1))
    RD = dsp.arm_rfft_fast_instance_f64()
    assert dsp.arm_rfft_fast_init_f64(RD, 512) == 0
    assert np.array_equal(dsp.arm_rfft_fast_f64(RD, xd, 0), oracle().rfft_f64(512, xd, 0))
    cfg = mfcc_config(512)
    M = dsp.arm_mfcc_instance_f32()
    assert dsp.arm_mfcc_init_f32(M, 512, 20, 13, cfg["dct"], cfg["pos"], cfg["len"], cfg["coefs"], cfg["window"]) == 0
    sig = rfft_input(512, frames=9, seed=5).reshape(-1)
    _mfcc_ok(dsp.arm_mfcc_f32(M, sig, None).reshape(9, 13), oracle().mfcc(cfg, sig))


# ------------------------------------------------------------------ deprecated radix-4 / radix-2 instance API

def test_deprecated_radix_api_against_the_compiled_reference():
    """arm_cfft_radix4_{q31,q15}: memcmp-identical to the reference's own deprecated functions (both flags, both
    directions); arm_cfft_radix4_f32 / arm_cfft_radix2_f32: relative RMS <= 2e-6 (bitReverseFlag = 0: tests/test_gpu_boundary.py)"""
    from oracle_lib import ref
    if ref() is None:
        pytest.skip("oracle/_ref not built")
    for kind in ("q31", "q15"):
        for N in (16, 64, 256, 1024, 4096):
            x = cfft_input(kind, N, frames=41, seed=N)
            for ifft in (0, 1):
                for bitrev in (1, 0):
                    assert np.array_equal(cd.cfft_radix_batch(kind, 4, N, x, ifft, bitrev), ref().cfft_radix(kind, 4, N, x, ifft, bitrev)), (kind, N, ifft, bitrev)
    for radix, lens in ((4, (16, 64, 256, 1024, 4096)), (2, LENGTHS)):
        for N in lens:
            x = cfft_input("f32", N, frames=41, seed=N)
            for ifft in (0, 1):
                assert relrms(cd.cfft_radix_batch("f32", radix, N, x, ifft, 1), ref().cfft_radix("f32", radix, N, x, ifft, 1)) <= F32_TOL, (radix, N, ifft)
    # legacy single-frame call
    S = cd.arm_cfft_radix4_instance_q15()
    L = cd.lib()
    assert L.arm_cfft_radix4_init_q15(C.byref(S), 256, 0, 1) == 0
    x = cfft_input("q15", 256, frames=1, seed=5).reshape(-1)
    y = x.copy()
    L.arm_cfft_radix4_q15(C.byref(S), y.ctypes.data)
    assert L.arm_cuda_last_status() == 0 and np.array_equal(y, oracle().cfft("q15", 256, x, 0, 1).reshape(-1))


# ------------------------------------------------------------------ fused spectrum epilogues

@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_magnitude_and_peak_epilogues(N):
    """arm_cfft_mag[_squared]_batch_f32 / arm_cfft_peak_batch_f32 against arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 +
    arm_max_f32 of the oracle; both kernel flavours; ragged batches; the source is left untouched"""
    cu = cd.cuda()
    rng = np.random.default_rng(17 * N)
    frames = 301
    k0 = rng.integers(1, N - 1, size=frames)
    n = np.arange(N)
    tone = np.exp(2j * np.pi * k0[:, None] * n[None, :] / N) * (0.5 + rng.random((frames, 1)))
    z = tone + 0.05 * (rng.standard_normal((frames, N)) + 1j * rng.standard_normal((frames, N)))
    x = np.empty((frames, 2 * N), dtype=np.float32)
    x[:, 0::2], x[:, 1::2] = z.real, z.imag
    x[7] = 0.0                                                       # all-zero frame: every bin ties, index 0 must win
    try:
        for flavour in (0, 1):
            assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
            for ifft in (0, 1):
                xin = x.copy()
                for sq in (False, True):
                    want = oracle().cfft_mag(N, x, ifft, sq)
                    got = cd.cfft_mag_batch(N, xin, ifft, sq)
                    assert np.array_equal(xin, x)
                    assert relrms(got, want) <= (2 * F32_TOL if sq else F32_TOL), (N, flavour, ifft, sq)
                wv, wi = oracle().cfft_mag(N, x, ifft, peak=True)
                gv, gi = cd.cfft_peak_batch(N, xin, ifft)
                assert gi[7] == 0 and gv[7] == 0.0
                # a clear tone: the peak bin is unambiguous (forward: k0; inverse: N - k0)
                assert np.array_equal(gi, wi), (N, flavour, ifft, np.flatnonzero(gi != wi)[:5])
                assert np.all(np.abs(gv - wv) <= 4e-6 * np.abs(wv)), (N, flavour, ifft)
                if not ifft:
                    keep = np.arange(frames) != 7
                    assert np.array_equal(gi[keep], k0[keep].astype(np.uint32))
    finally:
        cu.cmsisdsp_cuda_set_kernel_flavour(-1)


def test_fft_bin_example_fused_peak():
    """Examples/ARM/arm_fft_bin_example: the 10 kHz tone peaks at bin 213, now in ONE fused call.  The example's input
    is a real signal (imaginary parts zero), so bins 213 and 1024 - 213 tie mathematically and only rounding separates
    them: the reference's own arithmetic happens to favour 213, any other correctly rounded FFT may favour either."""
    d = np.load(os.path.join(HERE, "golden", "fft_bin_example.npz"))
    val, idx = cd.cfft_peak_batch(1024, d["input"], 0)
    assert int(d["ref_index"]) == 213 and int(idx[0]) in (213, 1024 - 213)
    wv, wi = oracle().cfft_mag(1024, d["input"], 0, peak=True)
    assert int(wi[0]) == 213 and abs(float(val[0]) - float(wv[0])) <= 4e-6 * float(wv[0])
    mag = cd.cfft_mag_batch(1024, d["input"], 0)[0]
    assert int(np.argmax(mag[:512])) == 213                        # the known answer on the non-redundant half


def test_spectrum_epilogues_device_pointers_large_batch():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    N, B = 1024, 1 << 16
    g = torch.Generator(device=dev).manual_seed(21)
    x = torch.randn(B, 2 * N, device=dev, generator=g)
    x0 = x.clone()
    mag = torch.empty(B, N, device=dev)
    val = torch.empty(B, device=dev)
    idx = torch.empty(B, device=dev, dtype=torch.int32)
    S = cd.cfft_instance("f32", N)
    L = cd.lib()
    assert L.arm_cfft_mag_batch_f32(C.byref(S), x.data_ptr(), mag.data_ptr(), B, 0) == 0, cd.last_error()
    assert L.arm_cfft_peak_batch_f32(C.byref(S), x.data_ptr(), val.data_ptr(), idx.data_ptr(), B, 0) == 0, cd.last_error()
    torch.cuda.synchronize()
    assert bool((x == x0).all().item())
    # size-independent properties: the peak is the maximum of the magnitudes, at the first index that holds it;
    # Parseval: sum |X|^2 = N sum |x|^2
    mx, am = mag.max(dim=1)
    assert bool((val == mx).all().item()) and bool((idx.long() == am).all().item() or (mag.gather(1, idx.long()[:, None])[:, 0] == mx).all().item())
    e_t = x.double().pow(2).sum(1) * N
    e_f = mag.double().pow(2).sum(1)
    assert float(((e_f - e_t).abs() / e_t).max()) < 1e-5
    sel = torch.arange(0, B, B // 256, device=dev)
    want = oracle().cfft_mag(N, x[sel].cpu().numpy(), 0, False)
    assert relrms(mag[sel].cpu().numpy(), want) <= F32_TOL


# ------------------------------------------------------------------ arm_rfft_q31 / arm_rfft_q15

@pytest.mark.parametrize("kind", ["q31", "q15"])
@pytest.mark.parametrize("N", RFIX_LENGTHS)
def test_rfft_fixed_point_bit_exact(kind, N):
    """forward and inverse through arm_rfft_batch_*, host buffers, ragged batch, full-scale frames: memcmp-identical
    to the oracle; digests of the compiled reference; pSrc left untouched by the batched call"""
    x = cfft_input(kind, N // 2, frames=203, seed=N)                # [203, N] real samples
    xin = x.copy()
    spec = cd.rfft_fix_batch(kind, N, xin, 0)
    assert np.array_equal(xin, x)
    assert spec.shape == (203, 2 * N)
    assert np.array_equal(spec, oracle().rfft_fix(kind, N, x, 0, 1, threads=NT)), (kind, N, "forward")
    src = np.concatenate([spec, cfft_input(kind, N, frames=37, seed=2 * N)])
    back = cd.rfft_fix_batch(kind, N, src, 1)
    assert np.array_equal(back, oracle().rfft_fix(kind, N, src, 1, 1, threads=NT)), (kind, N, "inverse")
    dig = ref_digests()
    x8 = cfft_input(kind, N // 2, frames=8, seed=3 * N)
    y = cd.rfft_fix_batch(kind, N, x8, 0)
    assert hashlib.sha256(y.tobytes()).hexdigest() == dig[f"rfft_{kind}/{N}/0"]
    z = cd.rfft_fix_batch(kind, N, np.concatenate([y, cfft_input(kind, N, frames=4, seed=5 * N)]), 1)
    assert hashlib.sha256(z.tobytes()).hexdigest() == dig[f"rfft_{kind}/{N}/1"]


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_rfft_fixed_point_reference_patterns_and_legacy_call(kind):
    n = 0
    for N, sig, ifft, x, ref in golden_rfft_fix_cases(kind):
        if not rfft_fix_threshold_applies(kind, N, sig, ifft):
            continue
        snr, want = assert_rfft_fix_like_reference(kind, N, ifft, cd.rfft_fix_batch(kind, N, x, ifft), ref)
        assert want is None or snr >= want, (kind, N, sig, ifft, snr)
        n += 1
    assert n == {"q31": 24, "q15": 22}[kind]
    # legacy single-frame call (arm_rfft_q31.c:145-181): same bits, and pSrc ends up holding the N/2-point CFFT
    N = 256
    L = cd.lib()
    x = cfft_input(kind, N // 2, frames=1, seed=77).reshape(-1)
    p, out = x.copy(), np.zeros(2 * N, dtype=x.dtype)
    S = cd.rfft_fix_instance(kind, N, 0, 1)
    getattr(L, f"arm_rfft_{kind}")(C.byref(S), p.ctypes.data, out.ctypes.data)
    assert L.arm_cuda_last_status() == 0, cd.last_error()
    assert np.array_equal(out, oracle().rfft_fix(kind, N, x, 0, 1)[0])
    assert np.array_equal(p, oracle().cfft(kind, N // 2, x, 0, 1).reshape(-1))
    Si = cd.rfft_fix_instance(kind, N, 1, 1)
    back = np.zeros(N, dtype=x.dtype)
    spec = out.copy()
    getattr(L, f"arm_rfft_{kind}")(C.byref(Si), spec.ctypes.data, back.ctypes.data)
    assert L.arm_cuda_last_status() == 0
    assert np.array_equal(back, oracle().rfft_fix(kind, N, out, 1, 1)[0]) and np.array_equal(spec, out)
    # bad arguments (bitReverseFlagR = 0 is a supported mode: tests/test_gpu_boundary.py)
    assert getattr(L, f"arm_rfft_batch_{kind}")(C.byref(S), p.ctypes.data, p.ctypes.data, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    assert getattr(L, f"arm_rfft_batch_{kind}")(C.byref(S), p.ctypes.data, out.ctypes.data, 0) == 0


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_rfft_fixed_point_large_batch_device_pointers(kind):
    """2^16 frames of N = 1024 resident on the device, forward then inverse through the shim; stratified oracle check"""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    N, B = 1024, 1 << 16
    tdt = torch.int32 if kind == "q31" else torch.int16
    info = np.iinfo(cd.NP_DTYPE[kind])
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randint(info.min // 2, info.max // 2, (B, N), device=dev, dtype=tdt, generator=g)
    spec = torch.empty(B, 2 * N, device=dev, dtype=tdt)
    back = torch.empty(B, N, device=dev, dtype=tdt)
    cd.ensure_rfft_fix_plans(kind, N)
    st = torch.cuda.current_stream().cuda_stream
    cd.rfft_fix_device(kind, N, x.data_ptr(), spec.data_ptr(), B, 0, st)
    cd.rfft_fix_device(kind, N, spec.data_ptr(), back.data_ptr(), B, 1, st)
    torch.cuda.synchronize()
    idx = torch.arange(0, B, B // 512, device=dev)
    xs = x[idx].cpu().numpy()
    want = oracle().rfft_fix(kind, N, xs, 0, 1, threads=NT)
    assert np.array_equal(spec[idx].cpu().numpy(), want)
    assert np.array_equal(back[idx].cpu().numpy(), oracle().rfft_fix(kind, N, want, 1, 1, threads=NT))
    # conjugate symmetry of every frame's spectrum (size-independent property): bin N-k == conj(bin k)
    sp = spec.view(B, N, 2)
    assert bool((sp[:, 1:N // 2, 0] == sp[:, N // 2 + 1:, 0].flip(1)).all().item())
    assert bool((sp[:, 1:N // 2, 1] == -sp[:, N // 2 + 1:, 1].flip(1)).all().item())


# ------------------------------------------------------------------ MFCC front end (BASELINE config 4)

def _mfcc_ok(got, want):
    """the reference's own MFCC thresholds (Testing/Source/Tests/MFCCF32.cpp:7-16)"""
    got, want = got.astype(np.float64), want.astype(np.float64)
    assert np.all(np.abs(got - want) <= 1e-5 + 1.2e-3 * np.abs(want)), float(np.abs(got - want).max())
    snr = 10 * np.log10((want ** 2).sum() / max(((want - got) ** 2).sum(), 1e-300))
    assert snr >= 115.0, snr


@pytest.mark.parametrize("n", [256, 512, 1024])
def test_mfcc_reference_patterns_and_oracle(n):
    from oracle_lib import mfcc_config
    d = np.load(os.path.join(HERE, "golden", "mfcc_patterns.npz"))
    cfg = mfcc_config(n)
    m = cd.Mfcc(cfg)
    for sig in ("noise", "sine"):                                   # the reference's own vectors
        _mfcc_ok(m.batch(d[f"{sig}/{n}/input"])[0], d[f"{sig}/{n}/ref"])
    rng = np.random.default_rng(n)
    t = np.arange(300 * n) / 16000.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1300 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t)
         + 0.1 * rng.standard_normal(t.size)).astype(np.float32)
    x[5 * n:6 * n] = 0.0                                            # an all-zero frame (maxValue == 0 branch)
    for hop in (n, n // 4, 2 * n, n // 2 + 2):                      # back to back, overlapping, gapped, not a multiple of 4 (direct kernel)
        want = oracle().mfcc(cfg, x, stride=hop, threads=NT)
        got = m.batch(x, hop=hop)
        assert got.shape == want.shape and got.shape[0] >= 149
        _mfcc_ok(got, want)
    # legacy single-frame call: same numbers, pSrc may be clobbered
    frame = x[:n].copy()
    one = np.zeros(13, dtype=np.float32)
    tmp = np.zeros(2 * n, dtype=np.float32)
    cd.lib().arm_mfcc_f32(C.byref(m.S), frame.ctypes.data, one.ctypes.data, tmp.ctypes.data)
    assert cd.lib().arm_cuda_last_status() == 0
    _mfcc_ok(one, oracle().mfcc(cfg, x[:n])[0])


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096])
def test_mfcc_both_kernels_ragged_batches(n):
    """TMA-fed and direct MFCC kernels against the oracle for ragged frame counts (partial last group,
    fewer groups than resident CTAs) and every supported fftLen (2048/4096 with a synthetic filter bank)."""
    from oracle_lib import mfcc_config
    cfg = mfcc_config(n)
    m = cd.Mfcc(cfg)
    cu = cd.cuda()
    rng = np.random.default_rng(7 * n)
    try:
        for frames in (1, 3, 130, 1031):
            x = rng.standard_normal((frames - 1) * (n // 2) + n).astype(np.float32)
            want = oracle().mfcc(cfg, x, stride=n // 2, threads=NT)
            for flavour in (0, 1):
                assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
                got = m.batch(x, hop=n // 2)
                assert got.shape == want.shape == (frames, 13)
                _mfcc_ok(got, want)
    finally:
        cu.cmsisdsp_cuda_set_kernel_flavour(-1)


@pytest.mark.parametrize("n", [256, 1024])
def test_mfcc_degenerate_frames(n):
    """frames that exercise the normalisation branches of arm_mfcc_f32.c:104-146: all zeros (absmax = 0: neither scaling is
    applied, every mel energy is 0 and the outputs are the DCT of log(1e-6)), one impulse (flat spectrum), very large and very
    small amplitudes (the scale by 1 / max keeps the squared magnitudes inside the f32 range) -- both kernels against the
    oracle, frame by frame.  (A frame whose upper bins are pure rounding noise, e.g. a constant, is NOT a parity case: the log
    of a noise-floor energy differs between any two correct implementations.)"""
    from oracle_lib import mfcc_config
    cfg = mfcc_config(n)
    m = cd.Mfcc(cfg)
    cu = cd.cuda()
    rng = np.random.default_rng(n)
    noise = rng.standard_normal(n).astype(np.float32)
    impulse = np.zeros(n, dtype=np.float32)
    impulse[n // 3] = -0.75
    names = ["zeros", "impulse", "noise * 1e30", "noise * 1e-30", "noise", "zeros again"]
    frames = np.stack([np.zeros(n, dtype=np.float32), impulse, noise * np.float32(1e30), noise * np.float32(1e-30), noise,
                       np.zeros(n, dtype=np.float32)])
    x = frames.reshape(-1)
    want = oracle().mfcc(cfg, x, threads=NT).astype(np.float64)
    assert np.all(np.isfinite(want))
    try:
        for flavour in (0, 1):
            assert cu.cmsisdsp_cuda_set_kernel_flavour(flavour) == 0
            got = m.batch(x, hop=n).astype(np.float64)
            assert np.all(np.isfinite(got)), flavour
            for k, name in enumerate(names):
                err = np.abs(got[k] - want[k])
                assert np.all(err <= 1e-5 + 1.2e-3 * np.abs(want[k])), (flavour, name, float(err.max()), got[k][:4].tolist(), want[k][:4].tolist())
            assert np.array_equal(got[0], got[-1])            # the two all-zero frames
    finally:
        cu.cmsisdsp_cuda_set_kernel_flavour(-1)


def test_mfcc_config4_synthetic_audio_large_batch():
    """BASELINE config 4 on one GPU: 16 kHz synthetic audio (3 sines + noise), 1024-sample frames, the
    reference's 20-mel / 13-DCT / Hamming configuration; device-resident, 2^17 frames, stratified oracle check"""
    torch = pytest.importorskip("torch")
    from oracle_lib import mfcc_config
    dev = torch.device("cuda", 0)
    n, B = 1024, 1 << 17
    cfg = mfcc_config(n)
    m = cd.Mfcc(cfg)
    g = torch.Generator(device=dev).manual_seed(4)
    t = torch.arange(B * n, device=dev, dtype=torch.float64) / 16000.0
    x = (0.5 * torch.sin(2 * np.pi * 440 * t) + 0.3 * torch.sin(2 * np.pi * 1300 * t) + 0.2 * torch.sin(2 * np.pi * 3100 * t)).float()
    x += 0.1 * torch.randn(B * n, device=dev, generator=g)
    out = torch.empty(B, 13, device=dev)
    assert cd.lib().arm_mfcc_batch_f32(C.byref(m.S), x.data_ptr(), n, out.data_ptr(), B) == 0, cd.last_error()
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all().item())
    idx = np.arange(0, B, B // 2048)
    xs = x.view(B, n)[torch.from_numpy(idx).to(dev)].cpu().numpy()
    want = oracle().mfcc(cfg, xs.reshape(-1), threads=NT)
    _mfcc_ok(out[torch.from_numpy(idx).to(dev)].cpu().numpy(), want)


def test_mfcc_argument_errors():
    from oracle_lib import mfcc_config
    m = cd.Mfcc(mfcc_config(256))
    x = np.zeros(1024, dtype=np.float32)
    out = np.zeros(13 * 4, dtype=np.float32)
    L = cd.lib()
    assert L.arm_mfcc_batch_f32(C.byref(m.S), x.ctypes.data, 255, out.ctypes.data, 2) == cd.ARM_MATH_ARGUMENT_ERROR   # odd hop
    assert L.arm_mfcc_batch_f32(C.byref(m.S), None, 256, out.ctypes.data, 2) == cd.ARM_MATH_ARGUMENT_ERROR
    assert L.arm_mfcc_batch_f32(C.byref(m.S), x.ctypes.data, 256, out.ctypes.data, 0) == 0                            # empty batch
    bad = cd.arm_mfcc_instance_f32()
    assert L.arm_mfcc_init_f32(C.byref(bad), 100, 20, 13, *[a.ctypes.data for a in m.arrs]) == cd.ARM_MATH_ARGUMENT_ERROR
