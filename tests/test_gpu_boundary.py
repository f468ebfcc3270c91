"""GPU tests of the drop-in boundary (run with `-m gpu` on a B200), round 2 additions:

* device buffers that are only element-aligned (offset by 8 or 4 bytes from a 16-byte boundary) for every entry point,
  and a clean error -- not a fault -- for buffers that are not even element-aligned
  (the reference takes any scalar-aligned buffer: Include/dsp/transform_functions.h:846-849)
* the C multi-device dispatcher: a host-pointer batch call fans out over the device list (arm_cuda_set_devices;
  an ordinal may repeat, so the partition logic is exercised on a one-GPU box too); every device's block is checked
* modes the reference executes: bitReverseFlagR = 0 for arm_rfft_q31 / q15 (arm_rfft_q31.c:164,173),
  bitReverseFlag = 0 for the deprecated arm_cfft_radix4_f32 / arm_cfft_radix2_f32 (arm_cfft_radix4_f32.c:81)
* the plan cache is keyed by the content of the instance's tables
* the inverse fixed-point real FFT reads fftLenReal + 2 scalars of a single frame, not 2 * fftLenReal
"""
import ctypes as C
import mmap

import numpy as np
import pytest

import cmsisdsp_b200 as cd
from oracle_lib import LENGTHS, mfcc_config, oracle, ref
from seeded_inputs import cfft_input, rfft_input

pytestmark = pytest.mark.gpu
F32_TOL = 2e-6


def relrms(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))


@pytest.fixture(scope="module", autouse=True)
def device():
    cu = cd.cuda()
    assert cu.cmsisdsp_cuda_device_count() >= 1, "no CUDA device: these tests must run on a GPU box"
    assert cu.cmsisdsp_cuda_set_device(0) == 0
    yield cu
    cd.set_devices(None)


def _torch():
    import torch
    return torch, torch.device("cuda", 0)


def _dev_copy(x, offset_bytes):
    """x on the device at an address that is `offset_bytes` past a 256-byte boundary; returns (pointer, keep-alive)"""
    torch, dev = _torch()
    raw = torch.empty(x.nbytes + 256, dtype=torch.uint8, device=dev)
    base = raw.data_ptr()
    off = (-base) % 256 + offset_bytes
    view = raw[off:off + x.nbytes]
    view.copy_(torch.from_numpy(x.view(np.uint8).reshape(-1)).to(dev))
    return base + off, (raw, view)


def _dev_empty(nbytes, offset_bytes):
    torch, dev = _torch()
    raw = torch.zeros(nbytes + 256, dtype=torch.uint8, device=dev)
    off = (-raw.data_ptr()) % 256 + offset_bytes
    return raw.data_ptr() + off, (raw, raw[off:off + nbytes])


def _back(keep, dtype, shape):
    torch, _ = _torch()
    torch.cuda.synchronize()
    return keep[1].cpu().numpy().view(dtype).reshape(shape)


# ------------------------------------------------------------------ element-aligned (not 16-byte aligned) device buffers

@pytest.mark.parametrize("N", [16, 64, 256, 1024, 4096])
def test_cfft_device_buffers_offset_from_16_bytes(N):
    L = cd.lib()
    for kind, off in (("f32", 8), ("q31", 8), ("q15", 4), ("q15", 12)):
        x = cfft_input(kind, N, frames=77, seed=N + off)
        S = cd.cfft_instance(kind, N)
        for ifft, bitrev in ((0, 1), (1, 0)):
            p, keep = _dev_copy(x, off)
            assert getattr(L, f"arm_cfft_batch_{kind}")(C.byref(S), p, 77, ifft, bitrev) == 0, cd.last_error()
            got = _back(keep, x.dtype, x.shape)
            want = oracle().cfft(kind, N, x, ifft, bitrev)
            if kind == "f32":
                assert relrms(got, want) <= F32_TOL, (kind, N, ifft, bitrev)
            else:
                assert np.array_equal(got, want), (kind, N, ifft, bitrev)


@pytest.mark.parametrize("N", [32, 128, 512, 2048, 4096])
def test_rfft_fast_device_buffers_offset_from_16_bytes(N):
    """source / destination 8 bytes past a 16-byte boundary, in every combination: the TMA-fed kernels need 16-byte
    aligned bulk copies and must hand such calls to the direct kernels (a misaligned bulk store is a sticky fault)"""
    L = cd.lib()
    x = rfft_input(N, frames=53, seed=N)
    S = cd.rfft_instance(N)
    spec = oracle().rfft(N, x, 0)
    for oin, oout in ((8, 0), (0, 8), (8, 8)):
        for ifft, src, want in ((0, x, spec), (1, spec, oracle().rfft(N, spec, 1))):
            p, keep_in = _dev_copy(np.ascontiguousarray(src), oin)
            q, keep_out = _dev_empty(src.nbytes, oout)
            assert L.arm_rfft_fast_batch_f32(C.byref(S), p, q, 53, ifft) == 0, cd.last_error()
            assert relrms(_back(keep_out, np.float32, src.shape), want) <= F32_TOL, (N, oin, oout, ifft)
            assert np.array_equal(_back(keep_in, np.float32, src.shape), src)          # the batched call leaves p untouched


@pytest.mark.parametrize("N", [32, 256, 2048])
def test_spectrum_epilogues_device_buffers_offset(N):
    L = cd.lib()
    x = cfft_input("f32", N, frames=45, seed=3 * N)
    S = cd.cfft_instance("f32", N)
    p, keep_in = _dev_copy(x, 8)
    q, keep_out = _dev_empty(45 * N * 4, 4)
    assert L.arm_cfft_mag_batch_f32(C.byref(S), p, q, 45, 0) == 0, cd.last_error()
    assert relrms(_back(keep_out, np.float32, (45, N)), oracle().cfft_mag(N, x, 0)) <= F32_TOL
    v, keep_v = _dev_empty(45 * 4, 4)
    i, keep_i = _dev_empty(45 * 4, 12)
    assert L.arm_cfft_peak_batch_f32(C.byref(S), p, v, i, 45, 0) == 0, cd.last_error()
    wv, wi = oracle().cfft_mag(N, x, 0, peak=True)
    assert np.array_equal(_back(keep_i, np.uint32, (45,)), wi)
    assert np.all(np.abs(_back(keep_v, np.float32, (45,)) - wv) <= 4e-6 * np.abs(wv))


@pytest.mark.parametrize("kind,off", [("q31", 8), ("q15", 4), ("q15", 12)])
@pytest.mark.parametrize("N", [32, 128, 1024, 8192])
def test_rfft_fixed_point_device_buffers_offset(kind, off, N):
    L = cd.lib()
    x = cfft_input(kind, N // 2, frames=29, seed=N).reshape(29, N)
    for ifft in (0, 1):
        src = x if not ifft else oracle().rfft_fix(kind, N, x, 0, 1)
        want = oracle().rfft_fix(kind, N, src, ifft, 1)
        S = cd.rfft_fix_instance(kind, N, ifft, 1)
        p, keep_in = _dev_copy(np.ascontiguousarray(src), off)
        q, keep_out = _dev_empty(want.nbytes, off)
        assert getattr(L, f"arm_rfft_batch_{kind}")(C.byref(S), p, q, 29) == 0, cd.last_error()
        assert np.array_equal(_back(keep_out, want.dtype, want.shape), want), (kind, N, ifft)


def test_mfcc_device_buffers_offset():
    cfg = mfcc_config(1024)
    m = cd.Mfcc(cfg)
    sig = rfft_input(1024, frames=9, seed=11).reshape(-1)
    frames = (sig.size - 1024) // 256 + 1
    want = oracle().mfcc(cfg, sig, stride=256)
    p, keep_in = _dev_copy(sig, 8)                                    # 8-byte aligned source: the direct kernel
    q, keep_out = _dev_empty(frames * 13 * 4, 4)
    assert cd.lib().arm_mfcc_batch_f32(C.byref(m.S), p, 256, q, frames) == 0, cd.last_error()
    got = _back(keep_out, np.float32, (frames, 13))
    assert np.all(np.abs(got - want) <= 1e-5 + 1.2e-3 * np.abs(want))


def test_scalar_aligned_device_buffers_get_an_error_not_a_fault():
    """a device pointer that is not aligned to one complex element is refused (ARM_MATH_ARGUMENT_ERROR); the context
    survives: the next call works"""
    L = cd.lib()
    x = cfft_input("f32", 256, frames=5, seed=1)
    S = cd.cfft_instance("f32", 256)
    p, keep = _dev_copy(x, 4)
    assert L.arm_cfft_batch_f32(C.byref(S), p, 5, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    xq = cfft_input("q15", 256, frames=5, seed=1)
    p2, keep2 = _dev_copy(xq, 2)
    assert L.arm_cfft_batch_q15(C.byref(cd.cfft_instance("q15", 256)), p2, 5, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    R = cd.rfft_instance(512)
    r = rfft_input(512, frames=5, seed=2)
    p3, keep3 = _dev_copy(r, 4)
    q3, keep4 = _dev_empty(r.nbytes, 0)
    assert L.arm_rfft_fast_batch_f32(C.byref(R), p3, q3, 5, 0) == cd.ARM_MATH_ARGUMENT_ERROR
    assert relrms(cd.cfft_batch("f32", 256, x, 0, 1), oracle().cfft("f32", 256, x, 0, 1)) <= F32_TOL


# ------------------------------------------------------------------ the C multi-device dispatcher

def _device_lists():
    n = cd.cuda().cmsisdsp_cuda_device_count()
    lists = [[0], [0, 0], [0, 0, 0]]                 # repeated ordinals: several workers on one device
    if n > 1:
        lists += [list(range(n)), list(range(n - 1, -1, -1))]
    return lists


def test_host_batches_fan_out_over_the_device_list():
    """arm_*_batch_* with host pointers: the frame range is block-partitioned over the device list, one host thread per
    device; every block is compared with the oracle (bit-exact for fixed point)."""
    L = cd.lib()
    N, frames = 1024, 9001                           # 70 MiB of f32: enough for every worker (8 MiB each at least); ragged
    x = cfft_input("f32", N, frames=frames, seed=5)
    xq = cfft_input("q15", N, frames=frames, seed=6)
    r = rfft_input(2 * N, frames=frames // 2, seed=7)
    want = oracle().cfft("f32", N, x, 0, 1, threads=8)
    wantq = oracle().cfft("q15", N, xq, 1, 0, threads=8)
    wantr = oracle().rfft(2 * N, r, 0, threads=8)
    for devs in _device_lists():
        cd.set_devices(devs)
        assert cd.get_devices() == devs
        got = cd.cfft_batch("f32", N, x, 0, 1)
        G = len(devs)
        per = -(-frames // G)
        for g in range(G):                           # every device's block
            lo, hi = min(frames, g * per), min(frames, (g + 1) * per)
            assert relrms(got[lo:hi], want[lo:hi]) <= F32_TOL, (devs, g)
        assert np.array_equal(cd.cfft_batch("q15", N, xq, 1, 0), wantq), devs
        assert relrms(cd.rfft_batch(2 * N, r, 0), wantr) <= F32_TOL, devs
    # the staging knobs: tiny chunks over many streams, one big chunk on one stream; chunk sizes ramping up from and
    # down to 1 / 2 MiB (x[:3000] is 23 MiB: a worker's 11.7 MiB travel as 1 + 2 + 4 + 2.4 + 1.2 + 1 MiB, say), no ramp
    for chunk_mib, nstreams, ramp_mib in ((1, 6, 0), (512, 1, 4), (4, 3, 1), (8, 2, 2), (32, 3, 4)):
        assert L.arm_cuda_set_staging(chunk_mib, nstreams) == 0
        assert L.arm_cuda_set_staging_ramp(ramp_mib) == 0
        cd.set_devices([0, 0])
        assert relrms(cd.cfft_batch("f32", N, x[:3000], 0, 1), want[:3000]) <= F32_TOL, (chunk_mib, nstreams, ramp_mib)
        assert np.array_equal(cd.cfft_batch("q15", N, xq[:5001], 1, 0), wantq[:5001]), (chunk_mib, nstreams, ramp_mib)
    assert L.arm_cuda_set_staging_ramp(5000) == cd.ARM_MATH_ARGUMENT_ERROR
    cd.set_devices(None)
    assert len(cd.get_devices()) == cd.cuda().cmsisdsp_cuda_device_count()
    assert L.arm_cuda_set_devices((C.c_int32 * 1)(99), 1) == cd.ARM_MATH_ARGUMENT_ERROR
    L.arm_cuda_release()                             # streams and staging buffers of this thread; the next call re-creates them
    assert relrms(cd.cfft_batch("f32", N, x[:100], 0, 1), want[:100]) <= F32_TOL


def test_host_batches_from_several_host_threads():
    """the library keeps no shared mutable state between calling threads (streams, staging and table selection are
    per thread): concurrent calls on distinct buffers with a shared const instance, as the reference allows"""
    import threading
    N = 512
    xs = [cfft_input("q31", N, frames=700, seed=s) for s in range(4)]
    wants = [oracle().cfft("q31", N, x, 0, 1) for x in xs]
    out = [None] * 4

    def run(k):
        out[k] = cd.cfft_batch("q31", N, xs[k], 0, 1)

    ts = [threading.Thread(target=run, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for k in range(4):
        assert np.array_equal(out[k], wants[k]), k


# ------------------------------------------------------------------ modes the reference executes

@pytest.mark.parametrize("kind", ["q31", "q15"])
@pytest.mark.parametrize("N", [32, 64, 128, 256, 1024, 4096, 8192])
def test_rfft_fixed_point_bit_reverse_flag_zero(kind, N):
    """bitReverseFlagR = 0 is handed to the complex transform inside (arm_rfft_q31.c:164,173): bit-exact against the
    oracle, both directions, host and device buffers"""
    x = cfft_input(kind, N // 2, frames=37, seed=N + 1).reshape(37, N)
    spec = oracle().rfft_fix(kind, N, x, 0, 0)
    assert not np.array_equal(spec, oracle().rfft_fix(kind, N, x, 0, 1))
    assert np.array_equal(cd.rfft_fix_batch(kind, N, x, 0, bitrev=0), spec), (kind, N)
    src = oracle().rfft_fix(kind, N, x, 0, 1)
    back = oracle().rfft_fix(kind, N, src, 1, 0)
    assert np.array_equal(cd.rfft_fix_batch(kind, N, src, 1, bitrev=0), back), (kind, N)
    # legacy single-frame call: also leaves the (unordered) complex transform in pSrc
    S = cd.rfft_fix_instance(kind, N, 0, 0)
    p = x[3].copy()
    out = np.zeros(2 * N, dtype=x.dtype)
    getattr(cd.lib(), f"arm_rfft_{kind}")(C.byref(S), p.ctypes.data, out.ctypes.data)
    assert cd.lib().arm_cuda_last_status() == 0, cd.last_error()
    assert np.array_equal(out, spec[3]) and np.array_equal(p, oracle().cfft(kind, N // 2, x[3], 0, 0).reshape(-1))


def test_deprecated_f32_radix_api_with_bit_reverse_flag_zero():
    """arm_cfft_radix4_f32 / arm_cfft_radix2_f32 with bitReverseFlag = 0 leave the spectrum in binary bit-reversed
    order: against the compiled reference where it is built, and against the oracle's natural-order result permuted"""
    def bitrev(N):
        b = N.bit_length() - 1
        return np.array([int(format(k, f"0{b}b")[::-1], 2) for k in range(N)])

    for radix, lens in ((4, (16, 64, 256, 1024, 4096)), (2, LENGTHS)):
        for N in lens:
            x = cfft_input("f32", N, frames=23, seed=N)
            P = bitrev(N)
            for ifft in (0, 1):
                got = cd.cfft_radix_batch("f32", radix, N, x, ifft, 0).reshape(23, N, 2)
                nat = oracle().cfft("f32", N, x, ifft, 1).reshape(23, N, 2)
                want = np.empty_like(nat)
                want[:, P] = nat                                       # X[k] sits at position bitrev(k)
                assert relrms(got, want) <= F32_TOL, (radix, N, ifft)
                if ref() is not None:
                    assert relrms(got.reshape(23, -1), ref().cfft_radix("f32", radix, N, x, ifft, 0)) <= F32_TOL, (radix, N, ifft)


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_deprecated_fixed_point_radix2_api_bit_exact(kind):
    """arm_cfft_radix2_q31 / arm_cfft_radix2_q15 (their own algorithm and per-stage scaling: arm_cfft_radix2_q31.c:87-318,
    arm_cfft_radix2_q15.c:275-386,577-681): memcmp-identical to the oracle restatement -- itself bit-identical to the
    compiled reference (tests/test_oracle_vs_ref.py) -- for every length, both directions, both values of the (ignored)
    bitReverseFlag, full-scale frames included; and to the compiled reference directly where it is built"""
    for N in LENGTHS:
        x = cfft_input(kind, N, frames=67, seed=N + 5)
        info = np.iinfo(x.dtype)
        x[0], x[1] = info.min, info.max
        x[2, 0::2], x[2, 1::2] = info.min, info.max
        for ifft in (0, 1):
            want = oracle().cfft_radix2_fix(kind, N, x, ifft)
            for bitrev in (1, 0):
                got = cd.cfft_radix_batch(kind, 2, N, x, ifft, bitrev)
                assert np.array_equal(got, want), (kind, N, ifft, bitrev)
            if ref() is not None:
                assert np.array_equal(want, ref().cfft_radix(kind, 2, N, x, ifft, 1)), (kind, N, ifft)
    # legacy single-frame call and a device buffer
    L = cd.lib()
    S = cd.RADIX_INSTANCE[kind]()
    assert getattr(L, f"arm_cfft_radix2_init_{kind}")(C.byref(S), 512, 1, 1) == 0
    assert getattr(L, f"arm_cfft_radix2_init_{kind}")(C.byref(cd.RADIX_INSTANCE[kind]()), 100, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    x = cfft_input(kind, 512, frames=3, seed=2)
    y = x[1].copy()
    getattr(L, f"arm_cfft_radix2_{kind}")(C.byref(S), y.ctypes.data)
    assert L.arm_cuda_last_status() == 0, cd.last_error()
    assert np.array_equal(y, oracle().cfft_radix2_fix(kind, 512, x[1], 1).reshape(-1))
    p, keep = _dev_copy(x, 8 if kind == "q31" else 4)
    assert getattr(L, f"arm_cfft_radix2_batch_{kind}")(C.byref(S), p, 3, ) == 0, cd.last_error()
    assert np.array_equal(_back(keep, x.dtype, x.shape), oracle().cfft_radix2_fix(kind, 512, x, 1))


# ------------------------------------------------------------------ pre-FFT window multiply as a fused prologue

@pytest.mark.parametrize("N", [32, 64, 128, 256, 1024, 4096])
def test_window_multiply_fused_into_the_transforms(N):
    """arm_rfft_fast_window_batch_f32 = arm_mult_f32(p, window) + arm_rfft_fast_f32 (the front of arm_mfcc_f32.c:112,137),
    arm_cfft_window_batch_f32 = a real window over the complex samples + arm_cfft_f32: against the oracle run on frames
    windowed in numpy with the same single-precision multiply; host and device buffers; a second window of the same
    length gets its own device copy"""
    L = cd.lib()
    n = np.arange(N)
    hamming = (0.54 - 0.46 * np.cos(2 * np.pi * n / N)).astype(np.float32)       # arm_hamming_f32.c:72 (sym = False)
    hann = (0.5 - 0.5 * np.cos(2 * np.pi * n / N)).astype(np.float32)
    x = rfft_input(N, frames=61, seed=N)
    R = cd.rfft_instance(N)
    for win in (hamming, hann, hamming):
        want = oracle().rfft(N, x * win[None, :], 0)
        out = np.zeros_like(x)
        xin = x.copy()
        assert L.arm_rfft_fast_window_batch_f32(C.byref(R), win.ctypes.data, xin.ctypes.data, out.ctypes.data, 61) == 0, cd.last_error()
        assert relrms(out, want) <= F32_TOL and np.array_equal(xin, x), N
    p, keep_in = _dev_copy(x, 8)
    q, keep_out = _dev_empty(x.nbytes, 0)
    assert L.arm_rfft_fast_window_batch_f32(C.byref(R), hann.ctypes.data, p, q, 61) == 0, cd.last_error()
    assert relrms(_back(keep_out, np.float32, x.shape), oracle().rfft(N, x * hann[None, :], 0)) <= F32_TOL
    z = cfft_input("f32", N, frames=47, seed=N + 9)
    S = cd.cfft_instance("f32", N)
    zw = (z.reshape(47, N, 2) * hamming[None, :, None]).reshape(47, 2 * N)
    for ifft in (0, 1):
        y = z.copy()
        assert L.arm_cfft_window_batch_f32(C.byref(S), hamming.ctypes.data, y.ctypes.data, 47, ifft) == 0, cd.last_error()
        assert relrms(y, oracle().cfft("f32", N, zw, ifft, 1)) <= F32_TOL, (N, ifft)
    assert L.arm_cfft_window_batch_f32(C.byref(S), None, z.ctypes.data, 47, 0) == cd.ARM_MATH_ARGUMENT_ERROR
    assert L.arm_rfft_fast_window_batch_f32(C.byref(R), hann.ctypes.data, x.ctypes.data, x.ctypes.data, 61) == cd.ARM_MATH_ARGUMENT_ERROR


# ------------------------------------------------------------------ plan cache keyed by table content

def test_a_second_instance_with_other_tables_gets_its_own_plan():
    """two instances of the same length whose tables differ: each call uses ITS instance's tables, in any order, and a
    copy of a table at another address is recognised by content"""
    N = 256
    x = cfft_input("q15", N, frames=19, seed=3)
    S = cd.cfft_instance("q15", N)
    std = oracle().cfft("q15", N, x, 0, 1)
    n_tw = 3 * N // 2
    tw = np.ctypeslib.as_array(S.pTwiddle, shape=(n_tw,)).copy()
    other = (tw // 2).astype(np.int16)                                  # every twiddle halved: a different transform
    S2 = cd.arm_cfft_instance_q15()
    S2.fftLen, S2.pBitRevTable, S2.bitRevLength = N, S.pBitRevTable, S.bitRevLength
    S2.pTwiddle = other.ctypes.data_as(C.POINTER(C.c_int16))
    L = cd.lib()

    def run(inst):
        y = x.copy()
        assert L.arm_cfft_batch_q15(C.byref(inst), y.ctypes.data, 19, 0, 1) == 0, cd.last_error()
        return y

    a1 = run(S)
    b1 = run(S2)
    a2 = run(S)
    assert np.array_equal(a1, std) and np.array_equal(a2, std)
    assert not np.array_equal(b1, std)
    # the same values in place of the first table's address: content decides, not the pointer
    other[:] = tw
    assert np.array_equal(run(S2), std)
    other[:] = tw // 2
    assert np.array_equal(run(S2), b1)
    if hasattr(oracle(), "cfft_q15_tables"):
        assert np.array_equal(b1, oracle().cfft_q15_tables(N, x, tw // 2))

    # MFCC: same addresses, new coefficient values (numpy reuses buffers) -> new device copies
    cfg = mfcc_config(256)
    m = cd.Mfcc(cfg)
    sig = rfft_input(256, frames=6, seed=9).reshape(-1)
    first = m.batch(sig)
    m.arrs[4][:] = m.arrs[4][::-1].copy()                              # window reversed in place
    cfg2 = dict(cfg, window=m.arrs[4].copy())
    second = m.batch(sig)
    want2 = oracle().mfcc(cfg2, sig)
    assert np.all(np.abs(second - want2) <= 1e-5 + 1.2e-3 * np.abs(want2))
    assert not np.allclose(first, second)


# ------------------------------------------------------------------ buffer contract of the inverse fixed-point real FFT

@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_inverse_fixed_point_rfft_reads_only_n_plus_2_scalars(kind):
    """arm_rfft_q31 / q15 inverse: the reference reads bins 0..N/2 (arm_rifft_input_buffer_size = N + 2 scalars,
    arm_rfft_q31.c:406-478).  The source sits flush against an unreadable page: reading 2 N scalars would fault."""
    N = 1024
    libc = C.CDLL(None, use_errno=True)
    page = mmap.PAGESIZE
    dt = cd.NP_DTYPE[kind]
    x = cfft_input(kind, N // 2, frames=1, seed=8).reshape(1, N)
    spec = oracle().rfft_fix(kind, N, x, 0, 1).reshape(-1)
    nbytes = (N + 2) * np.dtype(dt).itemsize
    npages = (nbytes + page - 1) // page
    buf = mmap.mmap(-1, (npages + 1) * page)
    base = C.addressof(C.c_char.from_buffer(buf))
    libc.mprotect.argtypes, libc.mprotect.restype = [C.c_void_p, C.c_size_t, C.c_int], C.c_int
    assert libc.mprotect(base + npages * page, page, 0) == 0           # PROT_NONE guard page
    start = base + npages * page - nbytes
    C.memmove(start, spec[:N + 2].ctypes.data, nbytes)
    S = cd.rfft_fix_instance(kind, N, 1, 1)
    out = np.zeros(N, dtype=dt)
    getattr(cd.lib(), f"arm_rfft_{kind}")(C.byref(S), start, out.ctypes.data)
    assert cd.lib().arm_cuda_last_status() == 0, cd.last_error()
    assert np.array_equal(out, oracle().rfft_fix(kind, N, spec.reshape(1, -1), 1, 1).reshape(-1))
    assert libc.mprotect(base + npages * page, page, 3) == 0
