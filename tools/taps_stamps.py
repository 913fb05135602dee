import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import boda_b200 as bb
from taps_experiments import time_conv, SHAPES
for prec in ("bf16", "fp32"):
    for dbg in [int(x) for x in (sys.argv[1:] or ['4','5','6','7'])]:
        rtc = bb.B200Compute(prec=prec, use_taps=1, taps_2cta=0, debug_flags=dbg); rtc.init()
        s = SHAPES[0]
        print(prec, "debug", dbg, "%.1f us" % (1e3 * time_conv(rtc, "x%d%s" % (dbg, prec), *s[1:], iters=3)), flush=True)
        rtc.close()
