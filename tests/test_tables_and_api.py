"""CPU-side checks of the product's C front library (no GPU needed):
tables value-identical to the reference / oracle, struct layouts, init contract, and that
both shared objects export every symbol their public headers declare."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cmsisdsp_b200 as cd
from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, oracle, ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view({4: np.uint32, 2: np.uint16}[a.dtype.itemsize])


@pytest.mark.parametrize("N", LENGTHS)
def test_generated_tables_match_oracle_and_reference(N):
    checkers = [oracle()] + ([ref()] if ref() is not None else [])
    for kind, which in (("f32", "f32"), ("q31", "fixed"), ("q15", "fixed")):
        for S in (cd.cfft_instance(kind, N), cd.preset(kind, N)):
            tw, br = cd.instance_tables(S, kind)
            for chk in checkers:
                assert np.array_equal(bits(tw), bits(chk.table(f"twiddle_{kind}", N))), (kind, N)
                assert np.array_equal(br, chk.bitrev(which, N)), (kind, N)
    # f64: the generated twiddles are the oracle's (within 1 ulp of the reference's literals, which follow no rule:
    # test_oracle_vs_ref.py); the swap lists are the reference's own (its fixed-point lists)
    for S in (cd.cfft_instance("f64", N), cd.preset("f64", N)):
        tw, br = cd.instance_tables(S, "f64")
        assert np.array_equal(tw.view(np.uint64), oracle().twiddle_f64(N).view(np.uint64))
        for chk in checkers:
            assert np.array_equal(br, chk.bitrev("fixed", N))
        if ref() is not None:
            assert np.abs(tw.view(np.int64) - ref().twiddle_f64(N).view(np.int64)).max() <= 1
    if N >= 32:
        S = cd.rfft_f64_instance(N)
        assert S.fftLenRFFT == N and S.Sint.fftLen == N // 2
        assert bytes(S.Sint) == bytes(cd.cfft_instance("f64", N // 2))
        twr64 = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
        assert np.array_equal(twr64.view(np.uint64), oracle().twiddle_rfft_f64(N).view(np.uint64))
    if N >= 32:
        for S in (cd.rfft_instance(N), cd.rfft_preset(N)):
            assert S.fftLenRFFT == N and S.Sint.fftLen == N // 2
            twr = np.ctypeslib.as_array(S.pTwiddleRFFT, shape=(N,)).copy()
            for chk in checkers:
                assert np.array_equal(bits(twr), bits(chk.table("twiddle_rfft_f32", N)))


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_real_coef_tables_and_rfft_fix_instances(kind):
    """realCoefA/B tables generated at build time == oracle == compiled reference; the instance fields are the
    reference's (arm_rfft_init_q31.c:97-127: modifier = 8192 / fftLenReal, pCfft = preset of half the length)"""
    checkers = [oracle()] + ([ref()] if ref() is not None else [])
    for N in RFIX_LENGTHS:
        for ifft in (0, 1):
            S = cd.rfft_fix_instance(kind, N, ifft, 1)
            assert (S.fftLenReal, S.ifftFlagR, S.bitReverseFlagR, S.twidCoefRModifier) == (N, ifft, 1, 8192 // N)
            assert S.pCfft.contents.fftLen == N // 2
            assert bytes(S.pCfft.contents) == bytes(cd.preset(kind, N // 2))
    S = cd.rfft_fix_instance(kind, 8192)
    A = np.ctypeslib.as_array(S.pTwiddleAReal, shape=(8192,)).copy()
    B = np.ctypeslib.as_array(S.pTwiddleBReal, shape=(8192,)).copy()
    for chk in checkers:
        assert np.array_equal(A, chk.real_coef(kind, 0)) and np.array_equal(B, chk.real_coef(kind, 1))
    L = cd.lib()
    for bad in (0, 16, 48, 16384):
        assert getattr(L, f"arm_rfft_init_{kind}")(C.byref(S), bad, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR
    for N in RFIX_LENGTHS:
        S2 = cd.RFIX_INSTANCE[kind]()
        assert getattr(L, f"arm_rfft_init_{N}_{kind}")(C.byref(S2), 1, 1) == cd.ARM_MATH_SUCCESS
        assert bytes(S2) == bytes(cd.rfft_fix_instance(kind, N, 1, 1))


def test_deprecated_radix_api_instances():
    """arm_cfft_radix4_init_* / arm_cfft_radix2_init_f32 fill the instance like the reference (strided 4096-point twiddles,
    one shared bit reversal table); armBitRevTable generated == the compiled reference's"""
    L = cd.lib()
    tab = np.ctypeslib.as_array((C.c_uint16 * 1024).in_dll(L, "armBitRevTable")).copy()
    if ref() is not None:
        fn = ref().lib.ref_arm_bit_rev_table
        fn.restype = C.POINTER(C.c_uint16)
        assert np.array_equal(tab, np.ctypeslib.as_array(fn(), shape=(1024,)))
        for name, t in (("cfft_radix4_instance_f32", cd.arm_cfft_radix4_instance_f32), ("cfft_radix4_instance_q31", cd.arm_cfft_radix4_instance_q31),
                        ("cfft_radix4_instance_q15", cd.arm_cfft_radix4_instance_q15)):
            f = getattr(ref().lib, f"ref_sizeof_{name}")
            f.restype = C.c_uint32
            assert f() == C.sizeof(t)
    for name, kind, lens in (("radix4", "f32", (16, 64, 256, 1024, 4096)), ("radix4", "q31", (16, 64, 256, 1024, 4096)),
                             ("radix4", "q15", (16, 64, 256, 1024, 4096)), ("radix2", "f32", LENGTHS)):
        init = getattr(L, f"arm_cfft_{name}_init_{kind}")
        for N in lens:
            S = cd.RADIX_INSTANCE[kind]()
            assert init(C.byref(S), N, 1, 0) == cd.ARM_MATH_SUCCESS
            assert (S.fftLen, S.ifftFlag, S.bitReverseFlag, S.twidCoefModifier, S.bitRevFactor) == (N, 1, 0, 4096 // N, 4096 // N)
            assert S.pBitRevTable[0] == tab[4096 // N - 1]
            big = cd.preset(kind, 4096)
            assert C.addressof(S.pTwiddle.contents) == C.addressof(big.pTwiddle.contents)
            if kind == "f32":
                assert S.onebyfftLen == np.float32(1.0 / N)
        S = cd.RADIX_INSTANCE[kind]()
        for bad in (0, 8, 24, 8192) + ((32, 128, 2048) if name == "radix4" else ()):
            assert init(C.byref(S), bad, 0, 1) == cd.ARM_MATH_ARGUMENT_ERROR


def test_struct_layouts_match_reference():
    # 32 / 32 / 32 / 48 bytes on LP64
    assert C.sizeof(cd.arm_cfft_instance_f32) == 32
    assert C.sizeof(cd.arm_cfft_instance_q31) == 32
    assert C.sizeof(cd.arm_cfft_instance_q15) == 32
    assert C.sizeof(cd.arm_rfft_fast_instance_f32) == 48
    if ref() is not None:
        L = ref().lib
        for name, t in (("cfft_instance_f32", cd.arm_cfft_instance_f32), ("cfft_instance_q31", cd.arm_cfft_instance_q31),
                        ("cfft_instance_q15", cd.arm_cfft_instance_q15), ("cfft_instance_f64", cd.arm_cfft_instance_f64), ("rfft_fast_instance_f64", cd.arm_rfft_fast_instance_f64), ("rfft_fast_instance_f32", cd.arm_rfft_fast_instance_f32),
                        ("rfft_instance_q31", cd.arm_rfft_instance_q31), ("rfft_instance_q15", cd.arm_rfft_instance_q15)):
            fn = getattr(L, f"ref_sizeof_{name}")
            fn.restype = C.c_uint32
            assert fn() == C.sizeof(t)


def test_init_contract():
    L = cd.lib()
    for kind, inst in cd.CFFT_INSTANCE.items():
        S = inst()
        for bad in (0, 8, 24, 100, 8192, 4095):
            assert getattr(L, f"arm_cfft_init_{kind}")(C.byref(S), bad) == cd.ARM_MATH_ARGUMENT_ERROR
        for N in LENGTHS:
            assert getattr(L, f"arm_cfft_init_{kind}")(C.byref(S), N) == cd.ARM_MATH_SUCCESS
            assert S.fftLen == N
            S2 = inst()
            assert getattr(L, f"arm_cfft_init_{N}_{kind}")(C.byref(S2)) == cd.ARM_MATH_SUCCESS
            assert bytes(S) == bytes(S2)
    R = cd.arm_rfft_fast_instance_f32()
    for bad in (0, 16, 48, 8192):
        assert L.arm_rfft_fast_init_f32(C.byref(R), bad) == cd.ARM_MATH_ARGUMENT_ERROR
    # NULL instance -> argument error (arm_rfft_fast_init_f32.c:87)
    assert L.arm_rfft_fast_init_1024_f32(None) == cd.ARM_MATH_ARGUMENT_ERROR
    for N in RLENGTHS:
        assert L.arm_rfft_fast_init_f32(C.byref(R), N) == cd.ARM_MATH_SUCCESS


def _declared_functions(header):
    src = open(header).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b((?:arm|cmsisdsp_cuda)_[A-Za-z0-9_]+)\s*\(", src))
    return {n for n in names if not n.isupper()}


def test_shared_objects_export_every_declared_symbol():
    cu = cd.cuda()
    for name in _declared_functions(os.path.join(ROOT, "include", "cmsisdsp_cuda.h")):
        assert hasattr(cu, name), name
    fr = cd.lib()
    names = _declared_functions(os.path.join(ROOT, "include", "dsp", "transform_functions.h"))
    # 4x(9 per-length inits + init + exec) [f32, q31, q15, f64] + 2 x rfft (8+1+1) [f32, f64] + 6 batch + last_status + mfcc (8+1+1+1)
    # + rfft_q31/q15 (init, exec, batch each; their per-length inits are declared through a macro)
    # + 3 fused spectrum epilogues (mag, mag squared, peak) + deprecated radix-4/2 API (4 x (init, exec, batch))
    # + device list / staging / release (arm_cuda_set_devices, _get_devices, _set_staging, _release, arm_mfcc_release_plans)
    # + deprecated fixed-point radix-2 (2 x (init, exec, batch)) + 2 windowed transforms
    assert len(names) == 117
    for name in names:
        assert hasattr(fr, name), name
    for N in RFIX_LENGTHS:
        for kind in ("q31", "q15"):
            assert hasattr(fr, f"arm_rfft_init_{N}_{kind}")
    for N in LENGTHS:
        for kind in ("f32", "q31", "q15", "f64"):
            cd.preset(kind, N)


def test_shard_partition():
    for B in (0, 1, 7, 8, 65536, 1000003):
        for G in (1, 2, 4, 8):
            parts = [cd.shard_frames(B, G, r) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and a <= b and c <= d


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device every exec entry point must FAIL (ARM_MATH_CUDA_NO_DEVICE, returned or latched), never
    compute on the CPU: the buffers stay untouched.  Skipped where a GPU is visible."""
    cu, L = cd.cuda(), cd.lib()
    if cu.cmsisdsp_cuda_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    x = np.arange(2 * 64, dtype=np.float32)
    x0 = x.copy()
    S = cd.cfft_instance("f32", 64)
    assert L.arm_cfft_batch_f32(C.byref(S), x.ctypes.data, 1, 0, 1) == cd.ARM_MATH_CUDA_NO_DEVICE
    L.arm_cfft_f32(C.byref(S), x.ctypes.data, 0, 1)
    assert L.arm_cuda_last_status() == cd.ARM_MATH_CUDA_NO_DEVICE and np.array_equal(x, x0)
    out = np.zeros(64, dtype=np.float32)
    R = cd.rfft_instance(64)
    assert L.arm_rfft_fast_batch_f32(C.byref(R), x.ctypes.data, out.ctypes.data, 1, 0) == cd.ARM_MATH_CUDA_NO_DEVICE
    assert L.arm_cfft_mag_batch_f32(C.byref(S), x.ctypes.data, out.ctypes.data, 1, 0) == cd.ARM_MATH_CUDA_NO_DEVICE
    idx = np.zeros(1, dtype=np.uint32)
    assert L.arm_cfft_peak_batch_f32(C.byref(S), x.ctypes.data, out.ctypes.data, idx.ctypes.data, 1, 0) == cd.ARM_MATH_CUDA_NO_DEVICE
    for kind in ("q31", "q15"):
        xi = np.arange(2 * 64, dtype=cd.NP_DTYPE[kind])
        Si = cd.cfft_instance(kind, 64)
        assert getattr(L, f"arm_cfft_batch_{kind}")(C.byref(Si), xi.ctypes.data, 1, 0, 1) == cd.ARM_MATH_CUDA_NO_DEVICE
        Sr = cd.rfft_fix_instance(kind, 64)
        oi = np.zeros(128, dtype=cd.NP_DTYPE[kind])
        assert getattr(L, f"arm_rfft_batch_{kind}")(C.byref(Sr), xi.ctypes.data, oi.ctypes.data, 1) == cd.ARM_MATH_CUDA_NO_DEVICE
        assert not oi.any()
    xd = np.arange(2 * 64, dtype=np.float64)
    Sd = cd.cfft_instance("f64", 64)
    assert L.arm_cfft_batch_f64(C.byref(Sd), xd.ctypes.data, 1, 0, 1) == cd.ARM_MATH_CUDA_NO_DEVICE
    L.arm_cfft_f64(C.byref(Sd), xd.ctypes.data, 0, 1)
    assert L.arm_cuda_last_status() == cd.ARM_MATH_CUDA_NO_DEVICE and np.array_equal(xd, np.arange(2 * 64, dtype=np.float64))
    Rd = cd.rfft_f64_instance(64)
    od = np.zeros(64)
    assert L.arm_rfft_fast_batch_f64(C.byref(Rd), xd.ctypes.data, od.ctypes.data, 1, 0) == cd.ARM_MATH_CUDA_NO_DEVICE and not od.any()
    bad64 = cd.arm_rfft_fast_instance_f64()
    for badlen in (0, 16, 48, 8192):
        assert L.arm_rfft_fast_init_f64(C.byref(bad64), badlen) == cd.ARM_MATH_ARGUMENT_ERROR
    assert L.arm_rfft_fast_init_1024_f64(None) == cd.ARM_MATH_ARGUMENT_ERROR
    assert not out.any() and np.array_equal(x, x0)
    assert cd.last_error() != ""
    assert cu.cmsisdsp_cuda_launch_count() == 0
