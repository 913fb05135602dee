#!/usr/bin/env python
"""Diagnosis of the contraction kernels on the AlexNet-ng B=32 layers (round-2 review item 3): per layer and precision, the kernel time
(min over launches, CUDA events) of the persistent CTA-pair im2col kernel -- normal, without TMA loads (debug bit 0: MMA on whatever shared
memory holds), without MMA issue (bit 1: loads + barrier traffic only) -- its per-role stall counters (bit 4: cycles the producer waits on
`empty`, the MMA warp on `full` / `tmem_empty`, the epilogue on `tmem_full`; printed by the library on stderr), and the tap-reuse kernel as
single CTAs and CTA pairs. Results with debug bits set are garbage by design; only their timing is used.
  python tools/diag_conv.py [bf16 fp32] > profiles/diag_r02_conv.txt 2>&1"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import boda_b200 as bb
from b200_harness import conv_op_text

SHAPES = [("conv2_5x5_96-256@27", 32, 96, 27, 27, 256, 5, 5, 1, 2), ("conv4_3x3_384-384@13", 32, 384, 13, 13, 384, 3, 3, 1, 1),
          ("conv5_3x3_384-256@13", 32, 384, 13, 13, 256, 3, 3, 1, 1), ("gn_3x3_128-192@28_b64", 64, 128, 28, 28, 192, 3, 3, 1, 1), ("k1_512-512@28", 32, 512, 28, 28, 512, 1, 1, 1, 0)]
VARIANTS = [("pair(r1)", dict(use_sk4=0)), ("sk4", dict()), ("sk4 im2col dp", dict(use_halo=0, use_streamk=0)), ("sk4 halo dp", dict(use_streamk=0)), ("sk4 halo dp noTMA", dict(use_streamk=0, debug_flags=1)),
            ("sk4 halo dp noMMA", dict(use_streamk=0, debug_flags=2)), ("sk4 stamps", dict(debug_flags=16)), ("sk4 halo dp stamps", dict(use_streamk=0, debug_flags=16))]


def time_conv(rtc, tag, N, C, H, W, OC, KH, KW, s, p, iters=6):
    OH, OW = (H + 2 * p - KH) // s + 1, (W + 2 * p - KW) // s + 1
    fn = "c_" + tag
    rtc.compile(fn, conv_op_text(N, C, H, W, OC, KH, KW, s, s, p, p, 1))
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


def main():
    precs = sys.argv[1:] or ["bf16", "fp32"]
    for prec in precs:
        for name, kw in VARIANTS:
            rtc = bb.B200Compute(prec=prec, **kw)
            rtc.init()
            res = []
            for s in SHAPES:
                iters = 3 if kw.get("debug_flags", 0) & 16 else 6
                gf = 2.0 * s[1] * s[5] * ((s[3] + 2 * s[9] - s[6]) // s[8] + 1) * ((s[4] + 2 * s[9] - s[7]) // s[8] + 1) * s[2] * s[6] * s[7] / 1e9
                try:
                    ms = time_conv(rtc, s[0].split("_")[0] + "_" + prec + "_" + name.replace(" ", "_"), *s[1:], iters=iters)
                    res.append("%s=%.1fus(%.0fTF)" % (s[0], 1e3 * ms, gf / ms))
                except Exception as e:
                    res.append("%s=ERR(%s)" % (s[0], str(e)[:60]))
            print("%-5s %-12s %s" % (prec, name, "  ".join(res)), flush=True)
            rtc.close()


if __name__ == "__main__":
    main()
