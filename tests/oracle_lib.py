"""ctypes access to the CPU checkers under oracle/ (TEST INFRASTRUCTURE).

`oracle()`  -> oracle/liboracle.so            (the restatement, built on demand)
`ref()`     -> oracle/_ref/libcmsisdsp_ref.so (the reference's own sources; prebuilt
               in the build container, travels to the GPU box, None if absent)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
LENGTHS = [16, 32, 64, 128, 256, 512, 1024, 2048, 4096]
RLENGTHS = [32, 64, 128, 256, 512, 1024, 2048, 4096]
RFIX_LENGTHS = [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]    # arm_rfft_q31 / arm_rfft_q15 real lengths

_u16p = C.POINTER(C.c_uint16)


def _declare(lib, prefix):
    f = getattr
    for name, t in (("cfft_f32", C.c_float), ("cfft_q31", C.c_int32), ("cfft_q15", C.c_int16)):
        fn = f(lib, f"{prefix}_{name}_batch")
        fn.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        fn.restype = None
    fn = f(lib, f"{prefix}_rfft_fast_f32_batch")
    fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
    fn.restype = None
    for name, t in (("twiddle_f32", C.c_float), ("twiddle_q31", C.c_int32),
                    ("twiddle_q15", C.c_int16), ("twiddle_rfft_f32", C.c_float)):
        fn = f(lib, f"{prefix}_{name}")
        fn.argtypes = [C.c_uint32]
        fn.restype = C.POINTER(t)
    for name in ("bitrev_f32", "bitrev_fixed"):
        fn = f(lib, f"{prefix}_{name}")
        fn.argtypes = [C.c_uint32, _u16p]
        fn.restype = _u16p
    for name, t in (("q31", C.c_int32), ("q15", C.c_int16)):
        fn = f(lib, f"{prefix}_rfft_{name}_batch")
        fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int]
        fn.restype = None
        fn = f(lib, f"{prefix}_real_coef_{name}")
        fn.argtypes = [C.c_int]
        fn.restype = C.POINTER(t)


class _Lib:
    """Uniform numpy-level wrapper over liboracle.so / libcmsisdsp_ref.so."""

    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        _declare(lib, prefix)

    def _fn(self, name):
        return getattr(self.lib, f"{self.prefix}_{name}")

    def cfft(self, kind, N, x, ifft=0, bitrev=1, threads=1):
        """x: array [..., 2N] of float32 / int32 / int16; returns a transformed copy."""
        dt = {"f32": np.float32, "q31": np.int32, "q15": np.int16}[kind]
        y = np.ascontiguousarray(x, dtype=dt).copy()
        assert y.size % (2 * N) == 0
        self._fn(f"cfft_{kind}_batch")(N, y.ctypes.data, y.size // (2 * N), int(ifft), int(bitrev), int(threads))
        return y

    def cfft_f64(self, N, x, ifft=0, bitrev=1, twiddle=None, threads=1):
        """arm_cfft_f64 on x [..., 2N] float64; returns a transformed copy.  twiddle (oracle only): the table to
        use instead of the oracle's generated one (e.g. the compiled reference's twiddleCoefF64_N)."""
        y = np.ascontiguousarray(x, dtype=np.float64).copy()
        assert y.size % (2 * N) == 0
        fn = self._fn("cfft_f64_batch")
        if self.prefix == "orc":
            tw = None if twiddle is None else np.ascontiguousarray(twiddle, dtype=np.float64)
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_int]
            fn.restype = None
            fn(N, y.ctypes.data, y.size // (2 * N), int(ifft), int(bitrev), None if tw is None else tw.ctypes.data, int(threads))
        else:
            assert twiddle is None
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
            fn.restype = C.c_int
            assert fn(N, y.ctypes.data, y.size // (2 * N), int(ifft), int(bitrev)) == 0
        return y

    def rfft_f64(self, N, x, ifft=0, twiddle_c=None, twiddle_r=None):
        """arm_rfft_fast_f64 on x [..., N] float64 (copied: the reference's forward transform destroys its input)."""
        src = np.ascontiguousarray(x, dtype=np.float64)
        assert src.size % N == 0
        out = np.empty_like(src)
        fn = self._fn("rfft_fast_f64_batch")
        if self.prefix == "orc":
            tc = None if twiddle_c is None else np.ascontiguousarray(twiddle_c, dtype=np.float64)
            tr = None if twiddle_r is None else np.ascontiguousarray(twiddle_r, dtype=np.float64)
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p]
            fn.restype = None
            fn(N, src.ctypes.data, out.ctypes.data, src.size // N, int(ifft), None if tc is None else tc.ctypes.data,
               None if tr is None else tr.ctypes.data)
        else:
            assert twiddle_c is None and twiddle_r is None
            fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
            fn.restype = C.c_int
            assert fn(N, src.ctypes.data, out.ctypes.data, src.size // N, int(ifft)) == 0
        return out

    def twiddle_rfft_f64(self, N):
        fn = self._fn("twiddle_rfft_f64")
        fn.argtypes = [C.c_uint32]
        fn.restype = C.POINTER(C.c_double)
        return np.ctypeslib.as_array(fn(N), shape=(N,)).copy()

    def twiddle_f64(self, N):
        fn = self._fn("twiddle_f64")
        fn.argtypes = [C.c_uint32]
        fn.restype = C.POINTER(C.c_double)
        return np.ctypeslib.as_array(fn(N), shape=(2 * N,)).copy()

    def rfft(self, N, x, ifft=0, threads=1, return_clobbered=False):
        p = np.ascontiguousarray(x, dtype=np.float32).copy()
        assert p.size % N == 0
        out = np.empty_like(p)
        self._fn("rfft_fast_f32_batch")(N, p.ctypes.data, out.ctypes.data, p.size // N, int(ifft), int(threads))
        return (out, p) if return_clobbered else out

    def rfft_fix(self, kind, N, x, ifft=0, bitrev=1, threads=1):
        """arm_rfft_q31 / arm_rfft_q15.  forward: x [..., N] -> [..., 2N] (N complex bins, mirror included);
        inverse: x [..., 2N] (bins 0..N/2 are read) -> [..., N]."""
        dt = {"q31": np.int32, "q15": np.int16}[kind]
        src = np.ascontiguousarray(x, dtype=dt)
        per = 2 * N if ifft else N
        assert src.size % per == 0
        frames = src.size // per
        out = np.empty(frames * (N if ifft else 2 * N), dtype=dt)
        self._fn(f"rfft_{kind}_batch")(N, src.ctypes.data, out.ctypes.data, frames, int(ifft), int(bitrev), int(threads))
        return out.reshape(frames, -1)

    def cfft_mag(self, N, x, ifft=0, squared=False, peak=False):
        """arm_cfft_f32 + arm_cmplx_mag[_squared]_f32 (+ arm_max_f32) per frame: returns mags [frames, N], or
        (values [frames], indices [frames]) with peak=True."""
        src = np.ascontiguousarray(x, dtype=np.float32)
        assert src.size % (2 * N) == 0
        frames = src.size // (2 * N)
        mag = np.empty((frames, N), dtype=np.float32)
        val, idx = np.empty(frames, dtype=np.float32), np.empty(frames, dtype=np.uint32)
        fn = self._fn("cfft_mag_f32_batch")
        fn.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        fn.restype = None
        fn(N, src.ctypes.data, mag.ctypes.data, val.ctypes.data, idx.ctypes.data, frames, int(ifft), int(squared))
        return (val, idx) if peak else mag

    def cfft_radix(self, kind, radix, N, x, ifft=0, bitrev=1):
        """the reference's deprecated arm_cfft_radix{4,2}_* (compiled reference only)"""
        dt = {"f32": np.float32, "q31": np.int32, "q15": np.int16}[kind]
        y = np.ascontiguousarray(x, dtype=dt).copy()
        fn = self._fn("cfft_radix_batch")
        fn.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
        fn.restype = C.c_int
        assert fn({"f32": 0, "q31": 1, "q15": 2}[kind], radix, N, y.ctypes.data, y.size // (2 * N), int(ifft), int(bitrev)) == 0
        return y

    def cfft_radix2_fix(self, kind, N, x, ifft=0):
        """deprecated arm_cfft_radix2_q31 / _q15 (oracle restatement; the compiled reference: cfft_radix(kind, 2, ...))"""
        dt = {"q31": np.int32, "q15": np.int16}[kind]
        y = np.ascontiguousarray(x, dtype=dt).copy()
        fn = self._fn(f"cfft_radix2_{kind}_batch")
        fn.argtypes = [C.c_uint32, C.c_void_p, C.c_uint64, C.c_int]
        fn.restype = None
        fn(N, y.ctypes.data, y.size // (2 * N), int(ifft))
        return y

    def real_coef(self, kind, b):
        return np.ctypeslib.as_array(self._fn(f"real_coef_{kind}")(int(b)), shape=(8192,)).copy()

    def mfcc(self, cfg, x, stride=None, frames=None, threads=1):
        """arm_mfcc_f32 over frames taken from the 1-D signal x every `stride` samples.
        cfg: dict(fftLen, nbMel, nbDct, dct, pos, len, coefs, window) of numpy arrays."""
        n = int(cfg["fftLen"])
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
        stride = n if stride is None else int(stride)
        frames = (x.size - n) // stride + 1 if frames is None else int(frames)
        assert (frames - 1) * stride + n <= x.size
        out = np.empty((frames, int(cfg["nbDct"])), dtype=np.float32)
        arrs = [np.ascontiguousarray(cfg["dct"], np.float32), np.ascontiguousarray(cfg["pos"], np.uint32),
                np.ascontiguousarray(cfg["len"], np.uint32), np.ascontiguousarray(cfg["coefs"], np.float32),
                np.ascontiguousarray(cfg["window"], np.float32)]
        fn = self._fn("mfcc_f32_batch")
        fn.argtypes = [C.c_uint32] * 3 + [C.c_void_p] * 6 + [C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
        fn(n, int(cfg["nbMel"]), int(cfg["nbDct"]), *[a.ctypes.data for a in arrs], x.ctypes.data, stride,
           out.ctypes.data, frames, int(threads))
        return out

    def table(self, name, N):
        n = {"twiddle_f32": 2 * N, "twiddle_q31": 3 * N // 2, "twiddle_q15": 3 * N // 2,
             "twiddle_rfft_f32": N}[name]
        ptr = self._fn(name)(N)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    def bitrev(self, which, N):
        ln = C.c_uint16(0)
        ptr = self._fn(f"bitrev_{which}")(N, C.byref(ln))
        return np.ctypeslib.as_array(ptr, shape=(ln.value,)).copy()


_cache = {}


def oracle():
    if "orc" not in _cache:
        so = os.path.join(ODIR, "liboracle.so")
        srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.startswith("orc_")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", ODIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        _cache["orc"] = _Lib(C.CDLL(so), "orc")
    return _cache["orc"]


def ref(fast=False):
    key = "ref_fast" if fast else "ref"
    if key not in _cache:
        so = os.path.join(ODIR, "_ref", "libcmsisdsp_ref_fast.so" if fast else "libcmsisdsp_ref.so")
        if not os.path.exists(so) and os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", ODIR, "ref"], stdout=subprocess.DEVNULL)
        _cache[key] = _Lib(C.CDLL(so), "ref") if os.path.exists(so) else None
    return _cache[key]


def perm_from_swaps(N, tab):
    """Apply the ordered swap list to the identity: out[k] = in[perm[k]]."""
    a = np.arange(N)
    t = (np.asarray(tab, dtype=np.int64) // 8).reshape(-1, 2)
    for x, y in t:
        a[x], a[y] = a[y], a[x]
    return a


def mfcc_config(n):
    """The reference's own MFCC test configuration for fftLen n in {256, 512, 1024}
    (Testing/Source/Tests/mfccdata.c via tests/golden/mfcc_patterns.npz): 20 mel filters, 13 DCT outputs."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "mfcc_patterns.npz"))
    if n not in (256, 512, 1024):
        return _mfcc_config_synth(n, d["dct"])
    return dict(fftLen=n, nbMel=20, nbDct=13, dct=d["dct"], pos=d[f"pos/{n}"], len=d[f"len/{n}"],
                coefs=d[f"coefs/{n}"], window=d[f"window/{n}"])


def _mfcc_config_synth(n, dct, fs=16000.0, fmin=64.0, fmax=8000.0, nb_mel=20):
    """Same recipe as the reference's generator (cmsisdsp/mfcc.py:28-113: Hamming window, triangular
    filters equally spaced on the mel scale, packed as pos/len/coefs) for lengths the reference ships
    no test data for (2048, 4096)."""
    mel = lambda f: 2595.0 * np.log10(1.0 + f / 700.0)
    imel = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    edges = imel(np.linspace(mel(fmin), mel(fmax), nb_mel + 2))
    freqs = np.arange(n // 2) * fs / n
    pos, length, coefs = [], [], []
    for k in range(nb_mel):
        lo, ce, hi = edges[k], edges[k + 1], edges[k + 2]
        w = np.maximum(0.0, np.minimum((freqs - lo) / (ce - lo), (hi - freqs) / (hi - ce)))
        nz = np.nonzero(w)[0]
        pos.append(int(nz[0])); length.append(int(nz[-1] - nz[0] + 1))
        coefs.append(w[nz[0]:nz[-1] + 1])
    window = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(n) / n)
    return dict(fftLen=n, nbMel=nb_mel, nbDct=dct.shape[0] if dct.ndim == 2 else 13, dct=dct,
                pos=np.asarray(pos, dtype=np.uint32), len=np.asarray(length, dtype=np.uint32),
                coefs=np.concatenate(coefs).astype(np.float32), window=window.astype(np.float32))
