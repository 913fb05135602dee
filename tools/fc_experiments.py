#!/usr/bin/env python
"""Timing experiments for inner-product shaped layers (swapped / split-K launches of the one-CTA contraction kernel): which phase bounds them."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import boda_b200 as bb
from kernel_experiments import time_conv

def main():
    shapes = [("gn_cls_fc1", 64, 128, 4, 4, 1024, 4, 1, 0), ("gn_cls_fc2", 64, 1024, 1, 1, 1000, 1, 1, 0), ("alex_fc8", 32, 4096, 1, 1, 1000, 1, 1, 0),
              ("alex_fc7", 32, 4096, 1, 1, 4096, 1, 1, 0), ("gn_icp_5x5", 64, 32, 14, 14, 64, 5, 1, 2)]
    for prec in sys.argv[1:] or ["bf16"]:
        for dbg in (0, 1, 2, 3, 4, 8, 15):
            rtc = bb.B200Compute(prec=prec)
            bb._chk(bb.lib().b200_rtc_set_option(rtc._h, b"debug_flags", str(dbg).encode()))
            rtc.init()
            res = ["%s=%.1fus" % (s[0], 1e3 * time_conv(rtc, "%s_%s_%d" % (s[0], prec, dbg), *s[1:])) for s in shapes]
            print("prec=%s debug=%d(%s) : %s" % (prec, dbg, "+".join(n for b, n in ((1, "noTMA"), (2, "noMMA"), (4, "noSTORE"), (8, "noDRAIN")) if dbg & b) or "normal", "  ".join(res)), flush=True)
            rtc.close()

if __name__ == "__main__":
    main()
