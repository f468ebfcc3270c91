#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory from the reference tree.

Run in the build container only (needs /root/reference and oracle/_ref):
    python tests/golden/make_golden.py

Outputs
  transform_patterns.npz  the reference's own test vectors for the FFT path
        (Testing/Patterns/DSP/Transform/Transform{F32,Q31,Q15}/*.txt, written by
        Testing/PatternGeneration/Transform.py from scipy.fftpack), re-packed as
        arrays keyed "<type>/<c|r>/<noisy|step>/<N>/<input|ref|ifft_input>".
  transform_patterns_f64.npz  same for Testing/Patterns/DSP/Transform/TransformF64/*.txt (arm_cfft_f64, arm_rfft_fast_f64).
  fft_bin_example.npz     Examples/ARM/arm_fft_bin_example/arm_fft_bin_data.c
        (testInput_f32_10khz; expected peak bin 213).
  mfcc_patterns.npz       Testing/Source/Tests/mfccdata.c coefficient arrays (DCT 13x20, Hamming
        windows, mel filter banks for 256/512/1024) and Testing/Patterns/DSP/Transform/MFCCF32/*
        (noise / sine inputs with their float64-computed references).
  ref_digests.json        sha256 of the outputs of the COMPILED REFERENCE
        (oracle/_ref/libcmsisdsp_ref.so) on seeded inputs, per
        (type, N, ifft, bitrev) -- pins bits where the reference's own tests only
        pin tolerances.  Inputs come from tests/seeded_inputs.py.
"""
import hashlib
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = "/root/reference"
PAT = os.path.join(REF, "Testing/Patterns/DSP/Transform")


def read_pattern(path):
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip() and not l.startswith("//")]
    tag, n = lines[0], int(lines[1])
    vals = [int(x, 16) for x in lines[2:2 + n]]
    if tag == "W":
        return np.array(vals, dtype=np.uint32)
    if tag == "H":
        return np.array(vals, dtype=np.uint16)
    if tag == "D":
        return np.array(vals, dtype=np.uint64)
    raise ValueError(f"unknown tag {tag} in {path}")


def main():
    if "--digests-only" not in sys.argv:
        patterns_main()
    digests_main()


def f64_patterns_main():
    """transform_patterns_f64.npz: the vectors of Testing/Patterns/DSP/Transform/TransformF64 (arm_cfft_f64, arm_rfft_fast_f64)."""
    out = {}
    d = os.path.join(PAT, "TransformF64")
    for fn in sorted(os.listdir(d)):
        m = re.match(r"(Complex|Real)(InputSamples|FFTSamples|InputIFFTSamples)_(Noisy|Step)_(\d+)_\d+_f64\.txt$", fn)
        if not m:
            continue
        cr, what, sig, n = m.groups()
        what = {"InputSamples": "input", "FFTSamples": "ref", "InputIFFTSamples": "ifft_input"}[what]
        out[f"f64/{'c' if cr == 'Complex' else 'r'}/{sig.lower()}/{n}/{what}"] = read_pattern(os.path.join(d, fn)).view(np.float64)
    np.savez_compressed(os.path.join(HERE, "transform_patterns_f64.npz"), **out)
    print("transform_patterns_f64.npz:", len(out), "arrays")


def patterns_main():
    out = {}
    for tname, ext, view in (("F32", "f32", np.float32), ("Q31", "q31", np.int32), ("Q15", "q15", np.int16)):
        d = os.path.join(PAT, f"Transform{tname}")
        for fn in sorted(os.listdir(d)):
            m = re.match(r"(Complex|Real)(InputSamples|FFTSamples|InputIFFTSamples)_(Noisy|Step)_(\d+)_\d+_" + ext + r"\.txt$", fn)
            if not m:
                continue
            cr, what, sig, n = m.groups()
            what = {"InputSamples": "input", "FFTSamples": "ref", "InputIFFTSamples": "ifft_input"}[what]
            key = f"{ext}/{'c' if cr == 'Complex' else 'r'}/{sig.lower()}/{n}/{what}"
            out[key] = read_pattern(os.path.join(d, fn)).view(view)
    np.savez_compressed(os.path.join(HERE, "transform_patterns.npz"), **out)
    print("transform_patterns.npz:", len(out), "arrays")
    f64_patterns_main()

    src = open(os.path.join(REF, "Examples/ARM/arm_fft_bin_example/arm_fft_bin_data.c")).read()
    body = src[src.index("testInput_f32_10khz"):]
    body = body[body.index("{") + 1:body.index("};")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    vals = np.array([float(t.rstrip("f")) for t in re.findall(r"[-+]?\d*\.\d+(?:[eE][-+]?\d+)?f?|[-+]?\d+\.?f?", body) if t.strip()],
                    dtype=np.float32)
    assert vals.size == 2048, vals.size
    np.savez_compressed(os.path.join(HERE, "fft_bin_example.npz"), input=vals, ref_index=np.int32(213))
    print("fft_bin_example.npz:", vals.size, "floats")

    # ---- MFCC: coefficient arrays of Testing/Source/Tests/mfccdata.c + the reference's patterns
    src = open(os.path.join(REF, "Testing/Source/Tests/mfccdata.c")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)

    def c_array(name, dtype):
        m = re.search(name + r"\[[A-Z0-9_]+\]\s*=\s*\{(.*?)\};", src, flags=re.S)
        toks = [t.strip().rstrip("f") for t in m.group(1).split(",") if t.strip()]
        return np.array([float(t) if dtype == np.float32 else int(t) for t in toks], dtype=dtype)

    mf = {"dct": c_array("mfcc_dct_coefs_config1_f32", np.float32)}
    for cfg, n in ((1, 1024), (2, 512), (3, 256)):
        mf[f"window/{n}"] = c_array(f"mfcc_window_coefs_config{cfg}_f32", np.float32)
        mf[f"pos/{n}"] = c_array(f"mfcc_filter_pos_config{cfg}_f32", np.uint32)
        mf[f"len/{n}"] = c_array(f"mfcc_filter_len_config{cfg}_f32", np.uint32)
        mf[f"coefs/{n}"] = c_array(f"mfcc_filter_coefs_config{cfg}_f32", np.float32)
        assert mf[f"window/{n}"].size == n and mf[f"coefs/{n}"].size == int(mf[f"len/{n}"].sum())
        for sig in ("Noise", "Sine"):
            for what in ("Input", "Ref"):
                fn = os.path.join(PAT, "MFCCF32", f"MFCC{sig}{what}_{n}_1_f32.txt")
                mf[f"{sig.lower()}/{n}/{what.lower()}"] = read_pattern(fn).view(np.float32)
    assert mf["dct"].size == 260
    np.savez_compressed(os.path.join(HERE, "mfcc_patterns.npz"), **mf)
    print("mfcc_patterns.npz:", len(mf), "arrays")



def digests_main():
    from oracle_lib import LENGTHS, RFIX_LENGTHS, RLENGTHS, ref
    from seeded_inputs import cfft_input, rfft_input
    r = ref()
    dig = {}
    for kind in ("f32", "q31", "q15"):
        for N in LENGTHS:
            x = cfft_input(kind, N, frames=8, seed=N)
            for ifft in (0, 1):
                for bitrev in (0, 1):
                    y = r.cfft(kind, N, x, ifft, bitrev)
                    dig[f"cfft_{kind}/{N}/{ifft}/{bitrev}"] = hashlib.sha256(y.tobytes()).hexdigest()
    for N in RLENGTHS:
        x = rfft_input(N, frames=8, seed=N)
        for ifft in (0, 1):
            y = r.rfft(N, x, ifft)
            dig[f"rfft_fast_f32/{N}/{ifft}"] = hashlib.sha256(y.tobytes()).hexdigest()
    for kind in ("q31", "q15"):
        for N in RFIX_LENGTHS:
            x = cfft_input(kind, N // 2, frames=8, seed=3 * N)           # [8, N] real samples incl. full-scale frames
            y = r.rfft_fix(kind, N, x, 0, 1)
            dig[f"rfft_{kind}/{N}/0"] = hashlib.sha256(y.tobytes()).hexdigest()
            z = r.rfft_fix(kind, N, np.concatenate([y, cfft_input(kind, N, frames=4, seed=5 * N)]), 1, 1)
            dig[f"rfft_{kind}/{N}/1"] = hashlib.sha256(z.tobytes()).hexdigest()
    with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
        json.dump(dig, f, indent=0, sort_keys=True)
    print("ref_digests.json:", len(dig), "digests")


if __name__ == "__main__":
    main()
