"""Pins the oracle restatement (oracle/liboracle.so) to the reference's own
generic-C sources (oracle/_ref/libcmsisdsp_ref.so): tables entry for entry and
every transform bit for bit, f32 included (both are built -ffp-contract=off)."""
import numpy as np
import pytest

from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, oracle, ref

pytestmark = pytest.mark.skipif(ref() is None, reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 2: np.uint16}[a.dtype.itemsize])


@pytest.mark.parametrize("N", LENGTHS)
def test_tables_identical(N):
    o, r = oracle(), ref()
    for name in ("twiddle_f32", "twiddle_q31", "twiddle_q15"):
        assert np.array_equal(bits(o.table(name, N)), bits(r.table(name, N))), name
    if N >= 32:
        assert np.array_equal(bits(o.table("twiddle_rfft_f32", N)), bits(r.table("twiddle_rfft_f32", N)))
    for which in ("f32", "fixed"):
        assert np.array_equal(o.bitrev(which, N), r.bitrev(which, N)), which


def _inputs(kind, N, rng, frames=6):
    if kind == "f32":
        x = rng.standard_normal((frames, 2 * N)).astype(np.float32)
        x[1] *= 1e-3
        x[2] *= 1e4
        return x
    info = np.iinfo(np.int32 if kind == "q31" else np.int16)
    x = rng.integers(info.min, info.max, size=(frames + 4, 2 * N), endpoint=True).astype(info.dtype)
    x[0] = info.min            # all 0x8000.. : saturation / wrap paths
    x[1] = info.max
    x[2, 0::2] = info.min
    x[2, 1::2] = info.max
    x[3] = np.where(np.arange(2 * N) % 4 < 2, info.max, info.min)
    x[4] = (rng.standard_normal(2 * N) * 0.25 * info.max).clip(info.min, info.max).astype(info.dtype)
    return x


@pytest.mark.parametrize("kind", ["f32", "q31", "q15"])
@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_bit_exact(kind, N):
    rng = np.random.default_rng(1000 + N)
    x = _inputs(kind, N, rng)
    for ifft in (0, 1):
        for bitrev in (0, 1):
            a = oracle().cfft(kind, N, x, ifft, bitrev)
            b = ref().cfft(kind, N, x, ifft, bitrev)
            assert np.array_equal(bits(a), bits(b)), (kind, N, ifft, bitrev)


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_bit_exact(N):
    rng = np.random.default_rng(2000 + N)
    x = rng.standard_normal((5, N)).astype(np.float32)
    for ifft in (0, 1):
        a, pa = oracle().rfft(N, x, ifft, return_clobbered=True)
        b, pb = ref().rfft(N, x, ifft, return_clobbered=True)
        assert np.array_equal(bits(a), bits(b)), (N, ifft)
        assert np.array_equal(bits(pa), bits(pb)), ("clobbered input", N, ifft)


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_real_coef_tables_identical(kind):
    for b in (0, 1):
        assert np.array_equal(oracle().real_coef(kind, b), ref().real_coef(kind, b)), (kind, b)


@pytest.mark.parametrize("kind", ["q31", "q15"])
@pytest.mark.parametrize("N", RFIX_LENGTHS)
def test_rfft_fixed_point_bit_exact(kind, N):
    """arm_rfft_q31 / arm_rfft_q15, forward and inverse, incl. full-scale frames (wrap / saturation paths)"""
    rng = np.random.default_rng(3000 + N)
    x = _inputs(kind, N // 2, rng)                                  # [frames, N] real samples
    a = oracle().rfft_fix(kind, N, x, 0, 1, threads=2)
    b = ref().rfft_fix(kind, N, x, 0, 1, threads=3)
    assert a.shape == (x.shape[0], 2 * N) and np.array_equal(a, b), (kind, N, "forward")
    spec = np.concatenate([a, _inputs(kind, N, rng)], axis=0)       # genuine spectra and arbitrary full-scale bins
    c = oracle().rfft_fix(kind, N, spec, 1, 1, threads=2)
    d = ref().rfft_fix(kind, N, spec, 1, 1, threads=3)
    assert c.shape == (spec.shape[0], N) and np.array_equal(c, d), (kind, N, "inverse")


@pytest.mark.parametrize("N", [16, 128, 1024, 4096])
def test_cfft_mag_and_peak_bit_exact(N):
    """arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 + arm_max_f32 restated == the compiled reference's own functions"""
    rng = np.random.default_rng(4000 + N)
    x = rng.standard_normal((7, 2 * N)).astype(np.float32)
    x[3] = 0.0
    x[4, 0::2] = 1.0; x[4, 1::2] = 0.0                               # constant frame: every bin but DC ties at ~0
    for ifft in (0, 1):
        for sq in (False, True):
            assert np.array_equal(bits(oracle().cfft_mag(N, x, ifft, sq)), bits(ref().cfft_mag(N, x, ifft, sq)))
        (v0, i0), (v1, i1) = oracle().cfft_mag(N, x, ifft, peak=True), ref().cfft_mag(N, x, ifft, peak=True)
        assert np.array_equal(bits(v0), bits(v1)) and np.array_equal(i0, i1)


def test_threads_agree():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((37, 2 * 256)).astype(np.float32)
    assert np.array_equal(oracle().cfft("f32", 256, x, threads=1), oracle().cfft("f32", 256, x, threads=5))
    assert np.array_equal(ref().cfft("f32", 256, x, threads=1), ref().cfft("f32", 256, x, threads=5))


@pytest.mark.parametrize("n", [256, 512, 1024])
def test_mfcc_bit_exact(n):
    """oracle restatement of arm_mfcc_f32 == compiled reference, bit for bit, incl. overlapping frames"""
    from oracle_lib import mfcc_config
    cfg = mfcc_config(n)
    rng = np.random.default_rng(n)
    t = np.arange(40 * n) / 16000.0
    x = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1300 * t) + 0.1 * rng.standard_normal(t.size)).astype(np.float32)
    x[3 * n:4 * n] = 0.0                                   # an all-zero frame: maxValue == 0 branch
    for stride in (n, n // 4):
        a = oracle().mfcc(cfg, x, stride=stride, threads=3)
        b = ref().mfcc(cfg, x, stride=stride, threads=2)
        assert a.shape == b.shape and a.shape[0] >= 37
        assert np.array_equal(a, b), (n, stride, np.abs(a - b).max())


# ------------------------------------------------------------------ arm_cfft_f64
def test_f64_tables_within_one_ulp_of_the_reference():
    """The reference's twiddleCoefF64_N literals follow no reproducible rule: the generated table may differ by one
    unit in the last place, never more; the bit-reversal lists are the fixed-point ones (identical)."""
    o, r = oracle(), ref()
    for N in LENGTHS:
        a, b = o.twiddle_f64(N), r.twiddle_f64(N)
        d = np.abs(a.view(np.int64) - b.view(np.int64))
        assert d.max() <= 1 and (d != 0).mean() <= 0.2, (N, int(d.max()), float((d != 0).mean()))
        assert not np.signbit(a[a == 0]).any()
        ln = np.zeros(1, dtype=np.uint16)
        fn = r._fn("bitrev_f64")
        import ctypes as C
        fn.argtypes, fn.restype = [C.c_uint32, C.c_void_p], C.POINTER(C.c_uint16)
        tab = np.ctypeslib.as_array(fn(N, ln.ctypes.data), shape=(int(ln[0]),))
        assert np.array_equal(tab, o.bitrev("fixed", N))


@pytest.mark.parametrize("N", LENGTHS)
def test_cfft_f64_bit_exact_on_the_reference_table_and_1e15_on_the_generated_one(N):
    rng = np.random.default_rng(3000 + N)
    x = rng.standard_normal((6, 2 * N))
    x[1] *= 1e-6
    x[2] *= 1e9
    tw = ref().twiddle_f64(N)
    for ifft in (0, 1):
        for bitrev in (0, 1):
            b = ref().cfft_f64(N, x, ifft, bitrev)
            a = oracle().cfft_f64(N, x, ifft, bitrev, twiddle=tw)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (N, ifft, bitrev)
            g = oracle().cfft_f64(N, x, ifft, bitrev)
            for f in range(x.shape[0]):
                assert np.sqrt(((g[f] - b[f]) ** 2).sum() / (b[f] ** 2).sum()) <= 1e-15, (N, ifft, bitrev, f)


@pytest.mark.parametrize("N", RLENGTHS)
def test_rfft_fast_f64_bit_exact_on_the_reference_tables(N):
    rng = np.random.default_rng(4000 + N)
    x = rng.standard_normal((5, N))
    tc, tr = ref().twiddle_f64(N // 2), ref().twiddle_rfft_f64(N)
    d = np.abs(oracle().twiddle_rfft_f64(N).view(np.int64) - tr.view(np.int64))
    assert d.max() <= 1 and (d != 0).mean() <= 0.2
    for ifft in (0, 1):
        b = ref().rfft_f64(N, x, ifft)
        assert np.array_equal(oracle().rfft_f64(N, x, ifft, tc, tr).view(np.uint64), b.view(np.uint64)), (N, ifft)
        g = oracle().rfft_f64(N, x, ifft)
        assert np.sqrt(((g - b) ** 2).sum() / (b ** 2).sum()) <= 1e-15, (N, ifft)


@pytest.mark.parametrize("kind", ["q31", "q15"])
@pytest.mark.parametrize("N", LENGTHS)
def test_deprecated_fixed_point_radix2_bit_exact(kind, N):
    """oracle/orc_cfft_radix2_fix.c == the reference's arm_cfft_radix2_q31 / _q15 (arm_cfft_radix2_q31.c:62-318,
    arm_cfft_radix2_q15.c:62-78,275-386,577-681), saturating / wrapping frames included; bitReverseFlag is ignored by
    the reference (it always bit-reverses), so both values must give the same bits"""
    rng = np.random.default_rng(3000 + N)
    x = _inputs(kind, N, rng)
    for ifft in (0, 1):
        a = oracle().cfft_radix2_fix(kind, N, x, ifft)
        for bitrev in (0, 1):
            assert np.array_equal(a, ref().cfft_radix(kind, 2, N, x, ifft, bitrev)), (kind, N, ifft, bitrev)
