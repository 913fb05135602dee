#!/usr/bin/env python
"""Timing experiments for the contraction kernel: which side (TMA fill vs MMA issue) bounds a k-block."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import boda_b200 as bb
from oracle import boda_oracle as bo
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from b200_harness import conv_op_text

def time_conv(rtc, tag, N, C, H, W, OC, K, s, p, iters=6):
    OH, OW = bo.conv_out_sz(H, p, s, K), bo.conv_out_sz(W, p, s, K)
    txt = conv_op_text(N, C, H, W, OC, K, K, s, s, p, p, 1)
    fn = "c_" + tag
    rtc.compile(fn, txt)
    x = bo.gen_conv_in(N, C, H, W); w = bo.gen_conv_filts(OC, C, K, K); b = bo.gen_conv_biases(OC)
    rtc.create_var_from_nda(fn + "_in", x, ["img", "chan", "y", "x"]); rtc.create_var_from_nda(fn + "_f", w, ["out_chan", "in_chan", "y", "x"])
    rtc.create_var_from_nda(fn + "_b", b, ["out_chan"]); rtc.create_var_with_dims(fn + "_o", [("img", N), ("chan", OC), ("y", OH), ("x", OW)])
    ids = [rtc.run(fn, {"in": fn + "_in", "filts": fn + "_f", "biases": fn + "_b", "out": fn + "_o"}) for _ in range(iters)]
    rtc.finish_and_sync()
    ms = min(rtc.get_kernel_dur(i) for i in ids[1:])
    for k in ("in", "f", "b", "o"):
        rtc.release_var(fn + "_" + k)
    return ms

def main():
    shapes = [("conv4_b32", 32, 384, 13, 13, 384, 3, 1, 1), ("conv2_b32", 32, 96, 27, 27, 256, 5, 1, 2), ("k1_512_b32", 32, 512, 28, 28, 512, 1, 1, 0)]
    for prec in ("fp32",):
        for two_cta in (0,):
            for dbg in (0, 3, 3 + 4, 3 + 8, 3 + 4 + 8, 4, 8, 4 + 8):
                for chunk in (4,):
                    rtc = bb.B200Compute(prec=prec, use_2cta=two_cta, use_clusters=0, acc_chunk_kblks=chunk)
                    bb._chk(bb.lib().b200_rtc_set_option(rtc._h, b"debug_flags", str(dbg).encode()))
                    rtc.init()
                    res = ["%s=%.1fus" % (s[0], 1e3 * time_conv(rtc, "%s_%s_%d_%d_%d" % (s[0], prec, two_cta, dbg, chunk), *s[1:])) for s in shapes]
                    print("prec=%s 2cta=%d debug=%d(%s) chunk=%d : %s" % (prec, two_cta, dbg, "+".join(n for b, n in ((1, "noTMA"), (2, "noMMA"), (4, "noSTORE"), (8, "noDRAIN")) if dbg & b) or "normal", chunk, "  ".join(res)), flush=True)
                    rtc.close()

if __name__ == "__main__":
    main()
