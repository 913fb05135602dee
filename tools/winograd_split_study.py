#!/usr/bin/env python
"""CPU study for DESIGN.md section 8: is Winograd F(2x2, 3x3) compatible with the fp32-parity mode's fp16 hi/lo operand split?

The transforms run in fp32 (as they would in pack-like kernels), the TRANSFORMED operands are scaled by a power of two and split into fp16
hi + lo exactly as the product does for plain operands, the 16 per-position contractions over the channels use the three products
hi*hi + hi*lo + lo*hi (accumulated in float64 here: only operand effects are studied), the output transform runs in fp32.
Data: the reference's gen_data mode-5 generators (signed, products cancel) and ReLU-like activations. mrd against the exact float64 convolution.
  python tools/winograd_split_study.py [C OC H]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import boda_oracle as bo

BT = np.array([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], np.float64)
G = np.array([[1, 0, 0], [0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0, 0, 1]], np.float64)
AT = np.array([[1, 1, 1, 0], [0, 1, -1, -1]], np.float64)


def split3(a, b):
    """sum_k a[.., k] * b[.., k] over the last axis with both operands carried as scaled fp16 hi + lo and the three products of the product path"""
    def sp(x):
        s = 2.0 ** (13 - np.floor(np.log2(max(np.abs(x).max(), 1e-30))))
        xs = x.astype(np.float64) * s
        hi = xs.astype(np.float16).astype(np.float64)
        lo = (xs - hi).astype(np.float16).astype(np.float64)
        return hi, lo, s
    ah, al, sa = sp(a)
    bh, bl, sb = sp(b)
    return (ah @ bh.T + ah @ bl.T + al @ bh.T) / (sa * sb)


def direct_exact(x, w):  # x [C][H][W] zero-padded by 1, w [OC][C][3][3] -> [OC][H][W], float64
    C, H, W = x.shape
    xp = np.zeros((C, H + 2, W + 2)); xp[:, 1:-1, 1:-1] = x
    out = np.zeros((w.shape[0], H, W))
    for ky in range(3):
        for kx in range(3):
            out += np.tensordot(w[:, :, ky, kx].astype(np.float64), xp[:, ky:ky + H, kx:kx + W], axes=(1, 0))
    return out


def direct_split(x, w):
    C, H, W = x.shape
    xp = np.zeros((C, H + 2, W + 2), np.float32); xp[:, 1:-1, 1:-1] = x
    cols = np.stack([xp[:, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)], 0)  # [9][C][H][W]
    a = cols.transpose(2, 3, 0, 1).reshape(H * W, 9 * C)
    b = w.transpose(0, 2, 3, 1).reshape(w.shape[0], 9 * C)
    return split3(a, b).T.reshape(w.shape[0], H, W)


def winograd_split(x, w):
    C, H, W = x.shape
    th, tw = (H + 1) // 2, (W + 1) // 2
    xp = np.zeros((C, 2 * th + 2, 2 * tw + 2), np.float32); xp[:, 1:H + 1, 1:W + 1] = x
    U = np.einsum("ij,ocjk,lk->oilc", G, w.astype(np.float64), G).astype(np.float32)  # [OC][4][4][C], fp32 as a pack kernel would store it
    out = np.zeros((w.shape[0], 2 * th, 2 * tw))
    tiles = np.stack([xp[:, 2 * ty:2 * ty + 4, 2 * tx:2 * tx + 4] for ty in range(th) for tx in range(tw)], 0)  # [T][C][4][4]
    V = np.einsum("ij,tcjk,lk->tilc", BT, tiles.astype(np.float64), BT).astype(np.float32)  # [T][4][4][C]
    Mm = np.zeros((th * tw, 4, 4, w.shape[0]))
    for i in range(4):
        for j in range(4):
            Mm[:, i, j, :] = split3(V[:, i, j, :], U[:, i, j, :])  # one contraction per transform position (scales per position)
    Y = np.einsum("ij,tjko,lk->toil", AT, Mm.astype(np.float32).astype(np.float64), AT)  # [T][OC][2][2]
    for t in range(th * tw):
        ty, tx = divmod(t, tw)
        out[:, 2 * ty:2 * ty + 2, 2 * tx:2 * tx + 2] = Y[t]
    return out[:, :H, :W]


def main():
    C, OC, H = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (384, 64, 13)
    x5 = bo.gen_conv_in(1, C, H, H)[0]
    w5 = bo.gen_conv_filts(OC, C, 3, 3)
    rng = np.random.RandomState(3)
    xr = (np.maximum(rng.standard_normal((C, H, H)) * 3.0, 0.0) * (rng.rand(C, H, H) < 0.6)).astype(np.float32)
    wr = (rng.standard_normal((OC, C, 3, 3)) * 0.02).astype(np.float32)
    for name, (x, w) in {"gen_data mode 5 (signed)": (x5, w5), "ReLU activations x small weights": (xr, wr)}.items():
        ex = direct_exact(x, w)
        print("%s  C=%d OC=%d %dx%d" % (name, C, OC, H, H))
        print("   %-44s mrd vs exact %.2e" % ("direct, fp16 hi/lo split (product path)", bo.mrd(ex.astype(np.float32), direct_split(x, w).astype(np.float32))))
        print("   %-44s mrd vs exact %.2e" % ("Winograd F(2x2,3x3), split after transform", bo.mrd(ex.astype(np.float32), winograd_split(x, w).astype(np.float32))))


if __name__ == "__main__":
    main()
