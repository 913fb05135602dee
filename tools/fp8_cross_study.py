#!/usr/bin/env python
"""CPU study for DESIGN.md section 9.7 (5): what would the fp32-parity mode lose if its two CROSS terms (hi*lo + lo*hi, 2^-11 of the result) were
computed from fp8 copies of the planes (tcgen05 kind::f8f6f4, twice the fp16 rate) instead of fp16 ones?

Operand model as in the product (pointwise.cuh / igemm.cuh): x_s = s*x with s a power of two putting max|x| in [2^13, 2^14), hi = fp16(x_s),
lo = fp16(x_s - hi). Products are accumulated in float64 here, so only OPERAND rounding is studied (accumulation effects are the same for all
variants). Data: the reference's gen_data mode-5 SGEMM operands (the oracle's generators) and a ReLU-like non-negative activation matrix.
Numbers reported: mrd (the reference's max relative difference, oracle.mrd) against the exact float64 product of the fp32 inputs.
  python tools/fp8_cross_study.py [M N K]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import boda_oracle as bo


def split16(x):
    s = 2.0 ** (13 - np.floor(np.log2(np.abs(x).max())))
    xs = (x.astype(np.float64) * s)
    hi = xs.astype(np.float16)
    lo = (xs - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), s


def round_mant(x, mant_bits, min_exp, max_val):
    """round-to-nearest-even to `mant_bits` explicit mantissa bits, exponents below min_exp become subnormal steps of 2^(min_exp - mant_bits), saturate at max_val"""
    x = np.asarray(x, np.float64)
    out = np.zeros_like(x)
    nz = x != 0
    e = np.floor(np.log2(np.abs(x[nz])))
    e = np.maximum(e, min_exp)
    q = 2.0 ** (e - mant_bits)
    out[nz] = np.clip(np.rint(x[nz] / q) * q, -max_val, max_val)
    return out


def e5m2(x):  # fp16's exponent range, 2 mantissa bits
    return round_mant(x, 2, -14, 57344.0)


def e4m3(x, pre):  # 3 mantissa bits, min normal 2^-6, max 448; `pre` = extra power-of-two scale applied before rounding (removed after)
    return round_mant(np.asarray(x) * pre, 3, -6, 448.0) / pre


def main():
    M, N, K = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (192, 160, 3456)
    cases = {"gen_data mode 5 (uniform, signed)": (bo.gen_sgemm_a(K, M, 5).T.copy(), bo.gen_sgemm_b(K, N, 5).T.copy())}
    rng = np.random.RandomState(7)
    act = np.maximum(rng.standard_normal((M, K)).astype(np.float32) * 3.0, 0.0) * (rng.rand(M, K) < 0.6)  # ReLU-like, 60 % dense, wide range
    w = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    cases["ReLU activations x small weights"] = (act.astype(np.float32), w)
    for name, (a, b) in cases.items():
        exact = a.astype(np.float64) @ b.astype(np.float64).T
        ah, al, sa = split16(a)
        bh, bl, sb = split16(b)
        inv = 1.0 / (sa * sb)
        main_t = ah @ bh.T
        res = {
            "one fp16 pass (hi*hi only)": main_t * inv,
            "three fp16 MMAs (product path)": (main_t + ah @ bl.T + al @ bh.T) * inv,
            "cross terms from e5m2 copies": (main_t + e5m2(ah) @ e5m2(bl).T + e5m2(al) @ e5m2(bh).T) * inv,
            "cross terms from e4m3 copies (hi * 2^-6, lo * 2^5)": (main_t + e4m3(ah, 2.0 ** -6) @ e4m3(bl, 2.0 ** 5).T + e4m3(al, 2.0 ** 5) @ e4m3(bh, 2.0 ** -6).T) * inv,
        }
        print("%s  (M=%d N=%d K=%d)" % (name, M, N, K))
        for k, v in res.items():
            print("   %-52s mrd vs exact %.2e" % (k, bo.mrd(exact.astype(np.float32), v.astype(np.float32))))


if __name__ == "__main__":
    main()
