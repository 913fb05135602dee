#!/usr/bin/env python
"""Timing experiments for the tap-reuse kernel (igemm3.cuh): which part of a (tap, block) step bounds it.
Each line: min kernel time (CUDA events) over a few launches; debug bit 0 = no TMA loads, bit 1 = no MMA issue (results are garbage)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import boda_b200 as bb
from oracle import boda_oracle as bo
from b200_harness import conv_op_text


def time_conv(rtc, tag, N, C, H, W, OC, KH, KW, py, px, iters=6):
    OH, OW = H + 2 * py - KH + 1, W + 2 * px - KW + 1
    fn = "c_" + tag
    rtc.compile(fn, conv_op_text(N, C, H, W, OC, KH, KW, 1, 1, py, px, 1))
    rng = np.random.RandomState(1)
    rtc.create_var_from_nda(fn + "_in", rng.rand(N, C, H, W).astype(np.float32), ["img", "chan", "y", "x"])
    rtc.create_var_from_nda(fn + "_f", rng.rand(OC, C, KH, KW).astype(np.float32), ["out_chan", "in_chan", "y", "x"])
    rtc.create_var_from_nda(fn + "_b", rng.rand(OC).astype(np.float32), ["out_chan"])
    rtc.create_var_with_dims(fn + "_o", [("img", N), ("chan", OC), ("y", OH), ("x", OW)])
    ids = [rtc.run(fn, {"in": fn + "_in", "filts": fn + "_f", "biases": fn + "_b", "out": fn + "_o"}) for _ in range(iters)]
    rtc.finish_and_sync()
    ms = min(rtc.get_kernel_dur(i) for i in ids[1:])
    for k in ("in", "f", "b", "o"):
        rtc.release_var(fn + "_" + k)
    return ms


SHAPES = [("conv3_3x3", 32, 256, 13, 13, 384, 3, 3, 1, 1), ("al_3x1_w16", 32, 256, 16, 16, 384, 3, 1, 1, 0), ("mis_1x3_w16", 32, 256, 16, 16, 384, 1, 3, 0, 1),
          ("conv2_5x5", 32, 96, 27, 27, 256, 5, 5, 2, 2)]
VARIANTS = [("im2col 1cta", dict(use_taps=0, use_2cta=0)), ("im2col 1cta noTMA", dict(use_taps=0, use_2cta=0, debug_flags=1)), ("im2col 1cta noMMA", dict(use_taps=0, use_2cta=0, debug_flags=2)),
            ("im2col 2cta", dict(use_taps=0, use_2cta=1)),
            ("taps 1cta", dict(use_taps=1, taps_2cta=0)), ("taps 1cta noTMA", dict(use_taps=1, taps_2cta=0, debug_flags=1)), ("taps 1cta noMMA", dict(use_taps=1, taps_2cta=0, debug_flags=2)),
            ("taps 1cta noTMA noMMA", dict(use_taps=1, taps_2cta=0, debug_flags=3)),
            ("taps 1cta b2", dict(use_taps=1, taps_2cta=0, taps_max_b_stages=3)), ("taps 1cta b4", dict(use_taps=1, taps_2cta=0, taps_max_b_stages=4)), ("taps 1cta a1", dict(use_taps=1, taps_2cta=0, taps_max_a_stages=1)),
            ("taps 1cta a2", dict(use_taps=1, taps_2cta=0, taps_max_a_stages=2)), ("taps 2cta", dict(use_taps=1, taps_2cta=1))]


def main():
    precs = sys.argv[1:] or ["bf16", "fp32"]
    for prec in precs:
        for name, kw in VARIANTS:
            rtc = bb.B200Compute(prec=prec, **kw)
            rtc.init()
            res = []
            for s in SHAPES:
                try:
                    res.append("%s=%.1fus" % (s[0], 1e3 * time_conv(rtc, s[0] + "_" + prec + "_" + name.replace(" ", "_"), *s[1:])))
                except Exception as e:
                    res.append("%s=ERR(%s)" % (s[0], str(e)[:40]))
            print("%-5s %-24s %s" % (prec, name, "  ".join(res)), flush=True)
            rtc.close()


if __name__ == "__main__":
    main()
