"""ctypes bindings of the B200 CMSIS-DSP FFT libraries (used by tests/, bench.py, smoke()).

Two shared objects, both built in-tree by `cmsis-dsp_b200/csrc/Makefile`:

  libcmsisdsp_b200.so   the C front library: arm_cfft_init_*/arm_cfft_* /arm_rfft_fast_* and the
                        batched extensions, i.e. the reference's operator API for the FFT path
                        (include/dsp/transform_functions.h)
  libcmsisdsp_cuda.so   the CUDA shim it calls (include/cmsisdsp_cuda.h)

This module is a binding, not an implementation: every transform goes through the C ABI.
It raises immediately if the libraries are missing -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# CMSISDSP_B200_LIBDIR: an alternative build of the two libraries (A/B measurements of compile-time choices)
LIBDIR = os.environ.get("CMSISDSP_B200_LIBDIR") or os.path.normpath(os.path.join(_HERE, "..", "..", "lib"))

LENGTHS = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
RLENGTHS = [32, 64, 128, 256, 512, 1024, 2048, 4096]
RFIX_LENGTHS = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]      # arm_rfft_q31 / arm_rfft_q15
ARM_MATH_SUCCESS = 0
ARM_MATH_ARGUMENT_ERROR = -1
ARM_MATH_CUDA_NO_DEVICE, ARM_MATH_CUDA_NO_PLAN, ARM_MATH_CUDA_RUNTIME_ERROR = -101, -102, -103
TYPE_ID = {"f32": 0, "q31": 1, "q15": 2, "f64": 3}
NP_DTYPE = {"f32": np.float32, "q31": np.int32, "q15": np.int16, "f64": np.float64}
C_SCALAR = {"f32": C.c_float, "q31": C.c_int32, "q15": C.c_int16, "f64": C.c_double}


def _mk_cfft_instance(scalar):
    class Inst(C.Structure):
        _fields_ = [("fftLen", C.c_uint16), ("pTwiddle", C.POINTER(scalar)),
                    ("pBitRevTable", C.POINTER(C.c_uint16)), ("bitRevLength", C.c_uint16)]
    return Inst


arm_cfft_instance_f32 = _mk_cfft_instance(C.c_float)
arm_cfft_instance_q31 = _mk_cfft_instance(C.c_int32)
arm_cfft_instance_q15 = _mk_cfft_instance(C.c_int16)
arm_cfft_instance_f64 = _mk_cfft_instance(C.c_double)
CFFT_INSTANCE = {"f32": arm_cfft_instance_f32, "q31": arm_cfft_instance_q31, "q15": arm_cfft_instance_q15,
                 "f64": arm_cfft_instance_f64}


class arm_rfft_fast_instance_f32(C.Structure):
    _fields_ = [("Sint", arm_cfft_instance_f32), ("fftLenRFFT", C.c_uint16), ("pTwiddleRFFT", C.POINTER(C.c_float))]


class arm_rfft_fast_instance_f64(C.Structure):
    _fields_ = [("Sint", arm_cfft_instance_f64), ("fftLenRFFT", C.c_uint16), ("pTwiddleRFFT", C.POINTER(C.c_double))]


def _mk_rfft_fix_instance(scalar, cfft):
    class Inst(C.Structure):
        _fields_ = [("fftLenReal", C.c_uint32), ("ifftFlagR", C.c_uint8), ("bitReverseFlagR", C.c_uint8),
                    ("twidCoefRModifier", C.c_uint32), ("pTwiddleAReal", C.POINTER(scalar)),
                    ("pTwiddleBReal", C.POINTER(scalar)), ("pCfft", C.POINTER(cfft))]
    return Inst


arm_rfft_instance_q31 = _mk_rfft_fix_instance(C.c_int32, arm_cfft_instance_q31)
arm_rfft_instance_q15 = _mk_rfft_fix_instance(C.c_int16, arm_cfft_instance_q15)
RFIX_INSTANCE = {"q31": arm_rfft_instance_q31, "q15": arm_rfft_instance_q15}


def _mk_radix_instance(scalar, with_oneby):
    class Inst(C.Structure):
        _fields_ = [("fftLen", C.c_uint16), ("ifftFlag", C.c_uint8), ("bitReverseFlag", C.c_uint8),
                    ("pTwiddle", C.POINTER(scalar)), ("pBitRevTable", C.POINTER(C.c_uint16)),
                    ("twidCoefModifier", C.c_uint16), ("bitRevFactor", C.c_uint16)] + ([("onebyfftLen", C.c_float)] if with_oneby else [])
    return Inst


# deprecated radix-4 / radix-2 instance API (arm_cfft_radix2_instance_f32 has the same fields as the radix-4 one)
arm_cfft_radix4_instance_f32 = _mk_radix_instance(C.c_float, True)
arm_cfft_radix4_instance_q31 = _mk_radix_instance(C.c_int32, False)
arm_cfft_radix4_instance_q15 = _mk_radix_instance(C.c_int16, False)
RADIX_INSTANCE = {"f32": arm_cfft_radix4_instance_f32, "q31": arm_cfft_radix4_instance_q31, "q15": arm_cfft_radix4_instance_q15}


class arm_mfcc_instance_f32(C.Structure):
    _fields_ = [("dctCoefs", C.POINTER(C.c_float)), ("filterCoefs", C.POINTER(C.c_float)),
                ("windowCoefs", C.POINTER(C.c_float)), ("filterPos", C.POINTER(C.c_uint32)),
                ("filterLengths", C.POINTER(C.c_uint32)), ("fftLen", C.c_uint32), ("nbMelFilters", C.c_uint32),
                ("nbDctOutputs", C.c_uint32), ("rfft", arm_rfft_fast_instance_f32)]


_libs = {}


def _load(name):
    path = os.path.join(LIBDIR, name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           f"or `make -C cmsis-dsp_b200/csrc` (there is no CPU fallback)")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def cuda():
    """libcmsisdsp_cuda.so with argtypes declared."""
    if "cuda" in _libs:
        return _libs["cuda"]
    L = _load("libcmsisdsp_cuda.so")
    vp, u32, u64, u8, i = C.c_void_p, C.c_uint32, C.c_uint64, C.c_uint8, C.c_int
    sig = {
        "cmsisdsp_cuda_device_count": ([], i), "cmsisdsp_cuda_set_device": ([i], i), "cmsisdsp_cuda_get_device": ([], i),
        "cmsisdsp_cuda_malloc": ([C.POINTER(vp), C.c_size_t], i), "cmsisdsp_cuda_free": ([vp], i),
        "cmsisdsp_cuda_host_alloc": ([C.POINTER(vp), C.c_size_t], i), "cmsisdsp_cuda_host_free": ([vp], i),
        "cmsisdsp_cuda_memcpy_h2d": ([vp, vp, C.c_size_t, vp], i), "cmsisdsp_cuda_memcpy_d2h": ([vp, vp, C.c_size_t, vp], i),
        "cmsisdsp_cuda_stream_create": ([C.POINTER(vp)], i), "cmsisdsp_cuda_stream_destroy": ([vp], i),
        "cmsisdsp_cuda_stream_synchronize": ([vp], i), "cmsisdsp_cuda_is_device_pointer": ([vp], i),
        "cmsisdsp_cuda_timer_begin": ([C.POINTER(vp), vp], i), "cmsisdsp_cuda_timer_end": ([vp, vp, C.POINTER(C.c_float)], i),
        "cmsisdsp_cuda_plan_upload": ([i, u32, vp, vp, C.c_uint16], i),
        "cmsisdsp_cuda_rfft_plan_upload": ([u32, vp], i),
        "cmsisdsp_cuda_plan_ready": ([i, u32], i), "cmsisdsp_cuda_rfft_plan_ready": ([u32], i),
        "cmsisdsp_cuda_cfft_f32": ([vp, u32, u64, u8, u8, vp], i),
        "cmsisdsp_cuda_cfft_q31": ([vp, u32, u64, u8, u8, vp], i),
        "cmsisdsp_cuda_cfft_q15": ([vp, u32, u64, u8, u8, vp], i),
        "cmsisdsp_cuda_cfft_f64": ([vp, u32, u64, u8, u8, vp], i),
        "cmsisdsp_cuda_rfft_fast_f32": ([vp, vp, u32, u64, u8, vp], i),
        "cmsisdsp_cuda_rfft_fast_f64": ([vp, vp, u32, u64, u8, vp], i),
        "cmsisdsp_cuda_rfft_f64_plan_upload": ([u32, vp], i), "cmsisdsp_cuda_rfft_f64_plan_ready": ([u32], i),
        "cmsisdsp_cuda_rfft_fix_plan_upload": ([i, u32, vp, vp, u32], i), "cmsisdsp_cuda_rfft_fix_plan_ready": ([i, u32], i),
        "cmsisdsp_cuda_rfft_q31": ([vp, vp, u32, u64, u8, u8, vp], i), "cmsisdsp_cuda_rfft_q15": ([vp, vp, u32, u64, u8, u8, vp], i),
        "cmsisdsp_cuda_cfft_f32_bitrev_order": ([vp, u32, u64, u8, vp], i), "cmsisdsp_cuda_pointer_device": ([vp], i),
        "cmsisdsp_cuda_radix2_plan_upload": ([i, u32, vp, u32], i),
        "cmsisdsp_cuda_cfft_radix2_q31": ([vp, u32, u64, u8, vp], i), "cmsisdsp_cuda_cfft_radix2_q15": ([vp, u32, u64, u8, vp], i),
        "cmsisdsp_cuda_cfft_mag_f32": ([vp, vp, u32, u64, u8, u8, vp], i), "cmsisdsp_cuda_cfft_peak_f32": ([vp, vp, vp, u32, u64, u8, vp], i),
        "cmsisdsp_cuda_last_error": ([], C.c_char_p), "cmsisdsp_cuda_launch_count": ([], u64),
        "cmsisdsp_cuda_set_kernel_flavour": ([i], i),
        "cmsisdsp_cuda_mfcc_plan_create": ([u32, u32, u32, vp, vp, vp, vp, vp, C.POINTER(vp)], i),
        "cmsisdsp_cuda_mfcc_plan_destroy": ([vp], i),
        "cmsisdsp_cuda_mfcc_f32": ([vp, vp, u64, vp, u64, vp], i),
        "cmsisdsp_cuda_kernel_info": ([i, u32] + [C.POINTER(i)] * 5, i),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)
        fn.argtypes, fn.restype = args, res
    _libs["cuda"] = L
    return L


def lib():
    """libcmsisdsp_b200.so (the CMSIS-DSP C API) with argtypes declared."""
    if "front" in _libs:
        return _libs["front"]
    cuda()
    L = _load("libcmsisdsp_b200.so")
    u8, u16, u32, i = C.c_uint8, C.c_uint16, C.c_uint32, C.c_int
    for k, inst in CFFT_INSTANCE.items():
        f = getattr(L, f"arm_cfft_init_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), u16], i
        for n in LENGTHS:
            f = getattr(L, f"arm_cfft_init_{n}_{k}")
            f.argtypes, f.restype = [C.POINTER(inst)], i
        f = getattr(L, f"arm_cfft_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p, u8, u8], None
        f = getattr(L, f"arm_cfft_batch_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p, u32, u8, u8], i
    L.arm_rfft_fast_init_f32.argtypes, L.arm_rfft_fast_init_f32.restype = [C.POINTER(arm_rfft_fast_instance_f32), u16], i
    for n in RLENGTHS:
        f = getattr(L, f"arm_rfft_fast_init_{n}_f32")
        f.argtypes, f.restype = [C.POINTER(arm_rfft_fast_instance_f32)], i
    L.arm_rfft_fast_f32.argtypes, L.arm_rfft_fast_f32.restype = [C.POINTER(arm_rfft_fast_instance_f32), C.c_void_p, C.c_void_p, u8], None
    L.arm_rfft_fast_batch_f32.argtypes = [C.POINTER(arm_rfft_fast_instance_f32), C.c_void_p, C.c_void_p, u32, u8]
    L.arm_rfft_fast_batch_f32.restype = i
    L.arm_rfft_fast_init_f64.argtypes, L.arm_rfft_fast_init_f64.restype = [C.POINTER(arm_rfft_fast_instance_f64), u16], i
    for n in RLENGTHS:
        f = getattr(L, f"arm_rfft_fast_init_{n}_f64")
        f.argtypes, f.restype = [C.POINTER(arm_rfft_fast_instance_f64)], i
    L.arm_rfft_fast_f64.argtypes, L.arm_rfft_fast_f64.restype = [C.POINTER(arm_rfft_fast_instance_f64), C.c_void_p, C.c_void_p, u8], None
    L.arm_rfft_fast_batch_f64.argtypes = [C.POINTER(arm_rfft_fast_instance_f64), C.c_void_p, C.c_void_p, u32, u8]
    L.arm_rfft_fast_batch_f64.restype = i
    L.arm_cuda_last_status.argtypes, L.arm_cuda_last_status.restype = [], i
    L.arm_cuda_set_devices.argtypes, L.arm_cuda_set_devices.restype = [C.POINTER(C.c_int32), u32], i
    L.arm_rfft_fast_window_batch_f32.argtypes = [C.POINTER(arm_rfft_fast_instance_f32), C.c_void_p, C.c_void_p, C.c_void_p, u32]
    L.arm_rfft_fast_window_batch_f32.restype = i
    L.arm_cfft_window_batch_f32.argtypes, L.arm_cfft_window_batch_f32.restype = [C.POINTER(arm_cfft_instance_f32), C.c_void_p, C.c_void_p, u32, u8], i
    L.arm_cuda_get_devices.argtypes, L.arm_cuda_get_devices.restype = [C.POINTER(C.c_int32), u32], u32
    L.arm_cuda_set_staging.argtypes, L.arm_cuda_set_staging.restype = [u32, u32], i
    L.arm_cuda_set_staging_ramp.argtypes, L.arm_cuda_set_staging_ramp.restype = [u32], i
    L.arm_cuda_release.argtypes, L.arm_cuda_release.restype = [], None
    L.arm_mfcc_release_plans.argtypes, L.arm_mfcc_release_plans.restype = [], None
    for name in ("arm_cfft_mag_batch_f32", "arm_cfft_mag_squared_batch_f32"):
        f = getattr(L, name)
        f.argtypes, f.restype = [C.POINTER(arm_cfft_instance_f32), C.c_void_p, C.c_void_p, u32, u8], i
    L.arm_cfft_peak_batch_f32.argtypes = [C.POINTER(arm_cfft_instance_f32), C.c_void_p, C.c_void_p, C.c_void_p, u32, u8]
    L.arm_cfft_peak_batch_f32.restype = i
    for name, k in (("radix4", "f32"), ("radix4", "q31"), ("radix4", "q15"), ("radix2", "f32"), ("radix2", "q31"), ("radix2", "q15")):
        inst = RADIX_INSTANCE[k]
        f = getattr(L, f"arm_cfft_{name}_init_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), u16, u8, u8], i
        f = getattr(L, f"arm_cfft_{name}_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p], None
        f = getattr(L, f"arm_cfft_{name}_batch_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p, u32], i
    for k, inst in RFIX_INSTANCE.items():
        f = getattr(L, f"arm_rfft_init_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), u32, u32, u32], i
        for n in RFIX_LENGTHS:
            f = getattr(L, f"arm_rfft_init_{n}_{k}")
            f.argtypes, f.restype = [C.POINTER(inst), u32, u32], i
        f = getattr(L, f"arm_rfft_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p, C.c_void_p], None
        f = getattr(L, f"arm_rfft_batch_{k}")
        f.argtypes, f.restype = [C.POINTER(inst), C.c_void_p, C.c_void_p, u32], i
    mp = C.POINTER(arm_mfcc_instance_f32)
    L.arm_mfcc_init_f32.argtypes, L.arm_mfcc_init_f32.restype = [mp, u32, u32, u32] + [C.c_void_p] * 5, i
    for n in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        f = getattr(L, f"arm_mfcc_init_{n}_f32")
        f.argtypes, f.restype = [mp, u32, u32] + [C.c_void_p] * 5, i
    L.arm_mfcc_f32.argtypes, L.arm_mfcc_f32.restype = [mp, C.c_void_p, C.c_void_p, C.c_void_p], None
    L.arm_mfcc_batch_f32.argtypes, L.arm_mfcc_batch_f32.restype = [mp, C.c_void_p, u32, C.c_void_p, u32], i
    _libs["front"] = L
    return L


def last_error():
    return cuda().cmsisdsp_cuda_last_error().decode()


def preset(kind, N):
    """The constant instance arm_cfft_sR_<kind>_len<N> exported by the front library."""
    return CFFT_INSTANCE[kind].in_dll(lib(), f"arm_cfft_sR_{kind}_len{N}")


def rfft_preset(N):
    return arm_rfft_fast_instance_f32.in_dll(lib(), f"arm_rfft_fast_sR_f32_len{N}")


def cfft_instance(kind, N):
    S = CFFT_INSTANCE[kind]()
    st = getattr(lib(), f"arm_cfft_init_{kind}")(C.byref(S), N)
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_cfft_init_{kind}({N}) -> {st}")
    return S


def rfft_instance(N):
    S = arm_rfft_fast_instance_f32()
    st = lib().arm_rfft_fast_init_f32(C.byref(S), N)
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_rfft_fast_init_f32({N}) -> {st}")
    return S


def rfft_f64_instance(N):
    S = arm_rfft_fast_instance_f64()
    st = lib().arm_rfft_fast_init_f64(C.byref(S), N)
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_rfft_fast_init_f64({N}) -> {st}")
    return S


def instance_tables(S, kind):
    """numpy copies of (twiddle table, bit-reversal swap list) an instance points at."""
    n = int(S.fftLen)
    ntw = 2 * n if kind in ("f32", "f64") else 3 * n // 2
    tw = np.ctypeslib.as_array(S.pTwiddle, shape=(ntw,)).copy()
    br = np.ctypeslib.as_array(S.pBitRevTable, shape=(int(S.bitRevLength),)).copy()
    return tw, br


# ---------------------------------------------------------------- host-buffer API (end to end)

def cfft_batch(kind, N, x, ifft=0, bitrev=1, inplace=False):
    """arm_cfft_batch_<kind> on a host numpy array [..., 2N]; returns the transformed array."""
    y = np.ascontiguousarray(x, dtype=NP_DTYPE[kind])
    if not inplace:
        y = y.copy()
    assert y.size % (2 * N) == 0
    S = cfft_instance(kind, N)
    st = getattr(lib(), f"arm_cfft_batch_{kind}")(C.byref(S), y.ctypes.data, y.size // (2 * N), int(ifft), int(bitrev))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_cfft_batch_{kind} -> {st}: {last_error()}")
    return y


def rfft_batch(N, x, ifft=0):
    p = np.ascontiguousarray(x, dtype=np.float32)
    assert p.size % N == 0
    out = np.empty_like(p)
    S = rfft_instance(N)
    st = lib().arm_rfft_fast_batch_f32(C.byref(S), p.ctypes.data, out.ctypes.data, p.size // N, int(ifft))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_rfft_fast_batch_f32 -> {st}: {last_error()}")
    return out


def rfft_f64_batch(N, x, ifft=0):
    """arm_rfft_fast_batch_f64 on a host array [..., N] of float64; returns the transformed array."""
    p = np.ascontiguousarray(x, dtype=np.float64)
    assert p.size % N == 0
    out = np.empty_like(p)
    S = rfft_f64_instance(N)
    st = lib().arm_rfft_fast_batch_f64(C.byref(S), p.ctypes.data, out.ctypes.data, p.size // N, int(ifft))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_rfft_fast_batch_f64 -> {st}: {last_error()}")
    return out


def cfft_mag_batch(N, x, ifft=0, squared=False):
    """arm_cfft_mag[_squared]_batch_f32 on a host array [..., 2N]; returns [frames, N] magnitudes."""
    src = np.ascontiguousarray(x, dtype=np.float32)
    assert src.size % (2 * N) == 0
    frames = src.size // (2 * N)
    out = np.empty((frames, N), dtype=np.float32)
    S = cfft_instance("f32", N)
    fn = lib().arm_cfft_mag_squared_batch_f32 if squared else lib().arm_cfft_mag_batch_f32
    st = fn(C.byref(S), src.ctypes.data, out.ctypes.data, frames, int(ifft))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_cfft_mag_batch_f32 -> {st}: {last_error()}")
    return out


def cfft_peak_batch(N, x, ifft=0):
    """arm_cfft_peak_batch_f32 on a host array [..., 2N]; returns (values [frames], indices [frames])."""
    src = np.ascontiguousarray(x, dtype=np.float32)
    assert src.size % (2 * N) == 0
    frames = src.size // (2 * N)
    val, idx = np.empty(frames, dtype=np.float32), np.empty(frames, dtype=np.uint32)
    S = cfft_instance("f32", N)
    st = lib().arm_cfft_peak_batch_f32(C.byref(S), src.ctypes.data, val.ctypes.data, idx.ctypes.data, frames, int(ifft))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_cfft_peak_batch_f32 -> {st}: {last_error()}")
    return val, idx


def cfft_radix_batch(kind, radix, N, x, ifft=0, bitrev=1):
    """deprecated arm_cfft_radix{4,2}_* instance API, batched: x [..., 2N] -> transformed copy"""
    y = np.ascontiguousarray(x, dtype=NP_DTYPE[kind]).copy()
    assert y.size % (2 * N) == 0
    S = RADIX_INSTANCE[kind]()
    st = getattr(lib(), f"arm_cfft_radix{radix}_init_{kind}")(C.byref(S), N, int(ifft), int(bitrev))
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_cfft_radix{radix}_init_{kind}({N}) -> {st}")
    st = getattr(lib(), f"arm_cfft_radix{radix}_batch_{kind}")(C.byref(S), y.ctypes.data, y.size // (2 * N))
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_cfft_radix{radix}_batch_{kind} -> {st}: {last_error()}")
    return y


def rfft_fix_instance(kind, N, ifft=0, bitrev=1):
    S = RFIX_INSTANCE[kind]()
    st = getattr(lib(), f"arm_rfft_init_{kind}")(C.byref(S), N, int(ifft), int(bitrev))
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_rfft_init_{kind}({N}) -> {st}")
    return S


def set_devices(devices=None):
    """arm_cuda_set_devices: the devices host-pointer batch calls fan out over (None / []: the default list)"""
    devices = list(devices or [])
    arr = (C.c_int32 * max(1, len(devices)))(*devices)
    st = lib().arm_cuda_set_devices(arr, len(devices))
    if st != ARM_MATH_SUCCESS:
        raise ValueError(f"arm_cuda_set_devices({devices}) -> {st}")


def get_devices():
    arr = (C.c_int32 * 16)()
    n = lib().arm_cuda_get_devices(arr, 16)
    return [int(arr[k]) for k in range(min(n, 16))]


def rfft_fix_batch(kind, N, x, ifft=0, bitrev=1):
    """arm_rfft_batch_<q31|q15> on a host array: forward [..., N] -> [frames, 2N]; inverse [..., 2N] -> [frames, N]."""
    src = np.ascontiguousarray(x, dtype=NP_DTYPE[kind])
    per = 2 * N if ifft else N
    assert src.size % per == 0
    frames = src.size // per
    out = np.empty((frames, N if ifft else 2 * N), dtype=NP_DTYPE[kind])
    S = rfft_fix_instance(kind, N, ifft, bitrev)
    st = getattr(lib(), f"arm_rfft_batch_{kind}")(C.byref(S), src.ctypes.data, out.ctypes.data, frames)
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"arm_rfft_batch_{kind} -> {st}: {last_error()}")
    return out


class Mfcc:
    """arm_mfcc_instance_f32 initialised from a config dict (fftLen, nbMel, nbDct, dct, pos, len, coefs,
    window as numpy arrays); keeps the arrays alive, as a C caller would keep its coefficient tables."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.arrs = [np.ascontiguousarray(cfg["dct"], np.float32), np.ascontiguousarray(cfg["pos"], np.uint32),
                     np.ascontiguousarray(cfg["len"], np.uint32), np.ascontiguousarray(cfg["coefs"], np.float32),
                     np.ascontiguousarray(cfg["window"], np.float32)]
        self.S = arm_mfcc_instance_f32()
        st = lib().arm_mfcc_init_f32(C.byref(self.S), int(cfg["fftLen"]), int(cfg["nbMel"]), int(cfg["nbDct"]),
                                     *[a.ctypes.data for a in self.arrs])
        if st != ARM_MATH_SUCCESS:
            raise ValueError(f"arm_mfcc_init_f32 -> {st}")

    def batch(self, x, hop=None, frames=None):
        """arm_mfcc_batch_f32 on a host signal x; returns [frames, nbDct]"""
        n = int(self.cfg["fftLen"])
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        hop = n if hop is None else int(hop)
        frames = (x.size - n) // hop + 1 if frames is None else int(frames)
        out = np.empty((frames, int(self.cfg["nbDct"])), dtype=np.float32)
        st = lib().arm_mfcc_batch_f32(C.byref(self.S), x.ctypes.data, hop, out.ctypes.data, frames)
        if st != ARM_MATH_SUCCESS:
            raise RuntimeError(f"arm_mfcc_batch_f32 -> {st}: {last_error()}")
        return out


# ---------------------------------------------------------------- device-pointer API (shim)

def ensure_plans(kind, N):
    S = cfft_instance(kind, N)
    rc = cuda().cmsisdsp_cuda_plan_upload(TYPE_ID[kind], N, C.cast(S.pTwiddle, C.c_void_p),
                                          C.cast(S.pBitRevTable, C.c_void_p), S.bitRevLength)
    if rc:
        raise RuntimeError(f"plan_upload({kind},{N}) -> {rc}: {last_error()}")


def ensure_rfft_plans(N):
    ensure_plans("f32", N // 2)
    S = rfft_instance(N)
    rc = cuda().cmsisdsp_cuda_rfft_plan_upload(N, C.cast(S.pTwiddleRFFT, C.c_void_p))
    if rc:
        raise RuntimeError(f"rfft_plan_upload({N}) -> {rc}: {last_error()}")


def ensure_rfft_fix_plans(kind, N):
    ensure_plans(kind, N // 2)
    S = rfft_fix_instance(kind, N)
    rc = cuda().cmsisdsp_cuda_rfft_fix_plan_upload(TYPE_ID[kind], N, C.cast(S.pTwiddleAReal, C.c_void_p),
                                                   C.cast(S.pTwiddleBReal, C.c_void_p), S.twidCoefRModifier)
    if rc:
        raise RuntimeError(f"rfft_fix_plan_upload({kind},{N}) -> {rc}: {last_error()}")


def rfft_fix_device(kind, N, d_in, d_out, n_frames, ifft=0, stream=0, bitrev=1):
    rc = getattr(cuda(), f"cmsisdsp_cuda_rfft_{kind}")(d_in, d_out, N, n_frames, int(ifft), int(bitrev), stream)
    if rc:
        raise RuntimeError(f"cmsisdsp_cuda_rfft_{kind} -> {rc}: {last_error()}")


def cfft_device(kind, N, dptr, n_frames, ifft=0, bitrev=1, stream=0):
    rc = getattr(cuda(), f"cmsisdsp_cuda_cfft_{kind}")(dptr, N, n_frames, int(ifft), int(bitrev), stream)
    if rc:
        raise RuntimeError(f"cmsisdsp_cuda_cfft_{kind} -> {rc}: {last_error()}")


def ensure_rfft_f64_plans(N):
    ensure_plans("f64", N // 2)
    S = rfft_f64_instance(N)
    rc = cuda().cmsisdsp_cuda_rfft_f64_plan_upload(N, C.cast(S.pTwiddleRFFT, C.c_void_p))
    if rc:
        raise RuntimeError(f"rfft_f64_plan_upload({N}) -> {rc}: {last_error()}")


def rfft_f64_device(N, d_in, d_out, n_frames, ifft=0, stream=0):
    rc = cuda().cmsisdsp_cuda_rfft_fast_f64(d_in, d_out, N, n_frames, int(ifft), stream)
    if rc:
        raise RuntimeError(f"cmsisdsp_cuda_rfft_fast_f64 -> {rc}: {last_error()}")


def rfft_device(N, d_in, d_out, n_frames, ifft=0, stream=0):
    rc = cuda().cmsisdsp_cuda_rfft_fast_f32(d_in, d_out, N, n_frames, int(ifft), stream)
    if rc:
        raise RuntimeError(f"cmsisdsp_cuda_rfft_fast_f32 -> {rc}: {last_error()}")


def kernel_info(op, N):
    v = [C.c_int(0) for _ in range(5)]
    rc = cuda().cmsisdsp_cuda_kernel_info(op, N, *[C.byref(x) for x in v])
    if rc:
        raise RuntimeError(f"kernel_info -> {rc}: {last_error()}")
    return dict(zip(("threads_per_cta", "frames_per_cta", "smem_bytes", "regs_per_thread", "ctas_per_sm"), [x.value for x in v]))


def shard_frames(n_frames, world_size, rank):
    """Contiguous block partition of the frame range over ranks:
    rank g owns frames [g*ceil(B/G), min(B, (g+1)*ceil(B/G)))."""
    per = -(-n_frames // world_size)
    lo = min(n_frames, rank * per)
    hi = min(n_frames, lo + per)
    return lo, hi
