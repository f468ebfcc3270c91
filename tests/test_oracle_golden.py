"""The oracle against the reference's OWN golden vectors and known-answer test
(SURVEY.md section 8(c)), plus the committed digests of the compiled reference's
outputs on seeded inputs (bit-level pin that also works where oracle/_ref is absent)."""
import hashlib

import numpy as np
import pytest

from golden_checks import (assert_like_reference, assert_rfft_fix_like_reference, golden_cases, golden_rfft_fix_cases,
                           ref_digests, rfft_fix_threshold_applies)
from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, oracle
from seeded_inputs import cfft_input, rfft_input
import os

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("kind", ["f32", "q31", "q15"])
def test_cfft_patterns(kind):
    n = 0
    for N, sig, ifft, x, ref in golden_cases(kind, "c"):
        out = oracle().cfft(kind, N, x, ifft, 1).reshape(-1)
        assert_like_reference(kind, "c", out, ref, N, ifft)
        n += 1
    assert n == 36          # 9 lengths x {noisy, step} x {fwd, inv}


def test_cfft_f64_patterns():
    """Testing/Patterns/DSP/Transform/TransformF64 with the thresholds of TransformCF64.cpp:6-8 (SNR 250 dB)."""
    n = 0
    for N, sig, ifft, x, ref in golden_cases("f64", "c"):
        out = oracle().cfft_f64(N, x, ifft, 1).reshape(-1)
        assert_like_reference("f64", "c", out, ref, N, ifft)
        n += 1
    assert n == 36


def test_rfft_fast_f64_patterns():
    """Real-FFT vectors of TransformF64 with the thresholds of TransformRF64.cpp:7-9."""
    n = 0
    for N, sig, ifft, x, ref in golden_cases("f64", "r"):
        out = oracle().rfft_f64(N, x, ifft).reshape(-1)
        assert_like_reference("f64", "r", out, ref, N, ifft)
        n += 1
    assert n == 32


def test_rfft_patterns():
    n = 0
    for N, sig, ifft, x, ref in golden_cases("f32", "r"):
        out = oracle().rfft(N, x, ifft).reshape(-1)
        assert_like_reference("f32", "r", out, ref, N, ifft)
        n += 1
    assert n == 32          # 8 lengths x {noisy, step} x {fwd, inv}


@pytest.mark.parametrize("kind", ["q31", "q15"])
def test_rfft_fixed_point_patterns(kind):
    n = 0
    for N, sig, ifft, x, ref in golden_rfft_fix_cases(kind):
        if not rfft_fix_threshold_applies(kind, N, sig, ifft):
            continue
        out = oracle().rfft_fix(kind, N, x, ifft, 1)
        snr, want = assert_rfft_fix_like_reference(kind, N, ifft, out, ref)
        assert want is None or snr >= want, (kind, N, sig, ifft, snr)
        n += 1
    assert n == {"q31": 24, "q15": 22}[kind]


def test_fft_bin_example_known_answer():
    """Examples/ARM/arm_fft_bin_example/arm_fft_bin_example_f32.c:127-155: CFFT-1024 of a
    10 kHz tone, magnitude, arg-max must be bin 213."""
    d = np.load(os.path.join(HERE, "golden", "fft_bin_example.npz"))
    y = oracle().cfft("f32", 1024, d["input"], 0, 1).reshape(-1, 2)
    mag = np.sqrt(y[:, 0].astype(np.float64) ** 2 + y[:, 1].astype(np.float64) ** 2)
    assert int(np.argmax(mag)) == int(d["ref_index"]) == 213      # bit-identical to the reference, so the first max wins


def test_digests_of_compiled_reference():
    dig = ref_digests()
    o = oracle()
    for kind in ("f32", "q31", "q15"):
        for N in LENGTHS:
            x = cfft_input(kind, N, frames=8, seed=N)
            for ifft in (0, 1):
                for bitrev in (0, 1):
                    y = o.cfft(kind, N, x, ifft, bitrev)
                    assert hashlib.sha256(y.tobytes()).hexdigest() == dig[f"cfft_{kind}/{N}/{ifft}/{bitrev}"], \
                        (kind, N, ifft, bitrev)
    for N in RLENGTHS:
        x = rfft_input(N, frames=8, seed=N)
        for ifft in (0, 1):
            y = o.rfft(N, x, ifft)
            assert hashlib.sha256(y.tobytes()).hexdigest() == dig[f"rfft_fast_f32/{N}/{ifft}"], (N, ifft)
    for kind in ("q31", "q15"):
        for N in RFIX_LENGTHS:
            x = cfft_input(kind, N // 2, frames=8, seed=3 * N)
            y = o.rfft_fix(kind, N, x, 0, 1)
            assert hashlib.sha256(y.tobytes()).hexdigest() == dig[f"rfft_{kind}/{N}/0"], (kind, N, 0)
            z = o.rfft_fix(kind, N, np.concatenate([y, cfft_input(kind, N, frames=4, seed=5 * N)]), 1, 1)
            assert hashlib.sha256(z.tobytes()).hexdigest() == dig[f"rfft_{kind}/{N}/1"], (kind, N, 1)


def test_unsupported_length_is_noop():
    x = np.arange(2 * 24, dtype=np.float32)
    assert np.array_equal(oracle().cfft("f32", 24, x), x)


# ------------------------------------------------------------------ MFCC (BASELINE config 4)

@pytest.mark.parametrize("n", [256, 512, 1024])
def test_mfcc_patterns(n):
    """the reference's own MFCC vectors and thresholds (Testing/Source/Tests/MFCCF32.cpp:7-16:
    SNR >= 115 dB, |err| <= 1e-5 + 1.2e-3 |ref|)"""
    from oracle_lib import mfcc_config
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "mfcc_patterns.npz"))
    cfg = mfcc_config(n)
    for sig in ("noise", "sine"):
        got = oracle().mfcc(cfg, d[f"{sig}/{n}/input"])[0].astype(np.float64)
        ref = d[f"{sig}/{n}/ref"].astype(np.float64)
        assert got.shape == ref.shape == (13,)
        snr = 10 * np.log10((ref ** 2).sum() / ((ref - got) ** 2).sum())
        assert snr >= 115.0, (n, sig, snr)
        assert np.all(np.abs(got - ref) <= 1e-5 + 1.2e-3 * np.abs(ref)), (n, sig)
