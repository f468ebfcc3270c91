"""Call-compatible subset of the reference's Python wrapper (`import cmsisdsp as dsp`,
PythonWrapper/cmsisdsp_pkg/src/cmsisdsp_transform.c:2074-2545) over the B200 libraries:

    import cmsisdsp_b200.compat as dsp
    S = dsp.arm_cfft_instance_f32()
    status = dsp.arm_cfft_init_f32(S, 1024)
    y = dsp.arm_cfft_f32(S, x, 0, 1)              # returns the transformed array, like the reference's wrapper

Same names, argument order and return conventions for the FFT path (cfft f32/q31/q15/f64, rfft_fast_f32/f64, rfft q31/q15,
mfcc f32).  One extension: an input holding several frames back to back is transformed as a batch in ONE call
(the reference's wrapper takes exactly one frame).  A binding, not an implementation: everything goes through the
C ABI; errors raise RuntimeError with the shim's message.
"""
import ctypes as C

import numpy as np

from . import (ARM_MATH_SUCCESS, CFFT_INSTANCE, NP_DTYPE, RFIX_INSTANCE, arm_mfcc_instance_f32 as _mfcc_struct,
               arm_rfft_fast_instance_f32 as _rfft_struct, arm_rfft_fast_instance_f64 as _rfft64_struct, last_error, lib)


def _check(st, what):
    if st != ARM_MATH_SUCCESS:
        raise RuntimeError(f"{what} -> {st}: {last_error()}")


def _frames(x, dtype, per):
    a = np.ascontiguousarray(x, dtype=dtype).reshape(-1)
    if a.size == 0 or a.size % per:
        raise ValueError(f"input length {a.size} is not a multiple of the frame length {per}")
    return a, a.size // per


def _mk_cfft(kind):
    inst = CFFT_INSTANCE[kind]

    def init(S, fftLen):
        return int(getattr(lib(), f"arm_cfft_init_{kind}")(C.byref(S), int(fftLen)))

    def run(S, p1, ifftFlag, bitReverseFlag=1):
        y, n = _frames(p1, NP_DTYPE[kind], 2 * S.fftLen)
        y = y.copy()
        _check(getattr(lib(), f"arm_cfft_batch_{kind}")(C.byref(S), y.ctypes.data, n, int(ifftFlag), int(bitReverseFlag)), f"arm_cfft_{kind}")
        return y
    return inst, init, run


arm_cfft_instance_f32, arm_cfft_init_f32, arm_cfft_f32 = _mk_cfft("f32")
arm_cfft_instance_q31, arm_cfft_init_q31, arm_cfft_q31 = _mk_cfft("q31")
arm_cfft_instance_q15, arm_cfft_init_q15, arm_cfft_q15 = _mk_cfft("q15")

arm_cfft_instance_f64, arm_cfft_init_f64, arm_cfft_f64 = _mk_cfft("f64")

arm_rfft_fast_instance_f32 = _rfft_struct
arm_rfft_fast_instance_f64 = _rfft64_struct


def arm_rfft_fast_init_f64(S, fftLen):
    return int(lib().arm_rfft_fast_init_f64(C.byref(S), int(fftLen)))


def arm_rfft_fast_f64(S, p, ifftFlag):
    src, n = _frames(p, np.float64, S.fftLenRFFT)
    out = np.empty_like(src)
    _check(lib().arm_rfft_fast_batch_f64(C.byref(S), src.ctypes.data, out.ctypes.data, n, int(ifftFlag)), "arm_rfft_fast_f64")
    return out


def arm_rfft_fast_init_f32(S, fftLen):
    return int(lib().arm_rfft_fast_init_f32(C.byref(S), int(fftLen)))


def arm_rfft_fast_f32(S, p, ifftFlag):
    src, n = _frames(p, np.float32, S.fftLenRFFT)
    out = np.empty_like(src)
    _check(lib().arm_rfft_fast_batch_f32(C.byref(S), src.ctypes.data, out.ctypes.data, n, int(ifftFlag)), "arm_rfft_fast_f32")
    return out


def _mk_rfft_fix(kind):
    inst = RFIX_INSTANCE[kind]

    def init(S, fftLenReal, ifftFlagR, bitReverseFlag):
        return int(getattr(lib(), f"arm_rfft_init_{kind}")(C.byref(S), int(fftLenReal), int(ifftFlagR), int(bitReverseFlag)))

    def run(S, pSrc):
        N = int(S.fftLenReal)
        per_in, per_out = (2 * N, N) if S.ifftFlagR else (N, 2 * N)
        a = np.ascontiguousarray(pSrc, dtype=NP_DTYPE[kind]).reshape(-1)
        if S.ifftFlagR and a.size == N + 2:                      # the reference's inverse reads bins 0..N/2 only, and so does
            src, n = a, 1                                        # the C entry point: an N + 2 buffer is enough for one frame
        else:
            src, n = _frames(a, NP_DTYPE[kind], per_in)
        out = np.empty(n * per_out, dtype=NP_DTYPE[kind])
        _check(getattr(lib(), f"arm_rfft_batch_{kind}")(C.byref(S), src.ctypes.data, out.ctypes.data, n), f"arm_rfft_{kind}")
        return out
    return inst, init, run


arm_rfft_instance_q31, arm_rfft_init_q31, arm_rfft_q31 = _mk_rfft_fix("q31")
arm_rfft_instance_q15, arm_rfft_init_q15, arm_rfft_q15 = _mk_rfft_fix("q15")


class arm_mfcc_instance_f32(_mfcc_struct):
    """keeps the coefficient arrays alive, as the reference's wrapper object does"""
    _keep = None


def arm_mfcc_init_f32(S, fftLen, nbMelFilters, nbDctOutputs, dctCoefs, filterPos, filterLengths, filterCoefs, windowCoefs):
    S._keep = [np.ascontiguousarray(dctCoefs, np.float32), np.ascontiguousarray(filterPos, np.uint32),
               np.ascontiguousarray(filterLengths, np.uint32), np.ascontiguousarray(filterCoefs, np.float32),
               np.ascontiguousarray(windowCoefs, np.float32)]
    return int(lib().arm_mfcc_init_f32(C.byref(S), int(fftLen), int(nbMelFilters), int(nbDctOutputs), *[a.ctypes.data for a in S._keep]))


def arm_mfcc_f32(S, pSrc, pTmp=None):
    """one frame, or several back to back (non-overlapping); pTmp is accepted and ignored like the C API's"""
    src, n = _frames(pSrc, np.float32, S.fftLen)
    out = np.empty(n * S.nbDctOutputs, dtype=np.float32)
    _check(lib().arm_mfcc_batch_f32(C.byref(S), src.ctypes.data, int(S.fftLen), out.ctypes.data, n), "arm_mfcc_f32")
    return out
