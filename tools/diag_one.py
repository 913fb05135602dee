#!/usr/bin/env python
"""One layer, a few variants, two launches each -- the workload for an `ncu --set full` A/B capture of the contraction kernels:
  ncu --set full --clock-control none --import-source on -k regex:igemm -o gpurun_out/prof python tools/diag_one.py k1 bf16"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import boda_b200 as bb
from diag_conv import SHAPES, time_conv

VARIANTS = {"r1": dict(use_sk4=0), "sk4_im2col_dp": dict(use_halo=0, use_streamk=0), "sk4_halo_dp": dict(use_streamk=0), "sk4": dict()}


def main():
    shape = next(s for s in SHAPES if s[0].startswith(sys.argv[1]))
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    for name in (sys.argv[3].split(",") if len(sys.argv) > 3 else list(VARIANTS)):
        rtc = bb.B200Compute(prec=prec, **VARIANTS[name])
        rtc.init()
        ms = time_conv(rtc, shape[0].split("_")[0] + "_" + name, *shape[1:], iters=2)
        print(name, shape[0], prec, "%.1f us" % (1e3 * ms), flush=True)
        rtc.close()


if __name__ == "__main__":
    main()
