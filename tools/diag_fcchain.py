#!/usr/bin/env python
"""AlexNet-ng forward with the fc_chain kernel's event stamps (debug_flags bit 4): where the chain's time goes, per layer, per barrier.
  python tools/diag_fcchain.py [prec] [batch] [extra opts, e.g. fc_l2_ahead=0,fc_l2_next=0]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import boda_b200 as bb
from boda_b200 import nets


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    txt, i, o = nets.alexnet_ng_conv(batch)
    params = nets.synth_params(txt)
    x = nets.synth_input((batch, 3, 227, 227))
    extra = ("," + sys.argv[3]) if len(sys.argv) > 3 else ""
    fwd = bb.B200ConvFwd(txt, "(prec=%s,use_graph=0,debug_flags=16%s)" % (prec, extra))
    for k, v in params.items():
        fwd.set_param(k, v)
    for it in range(2):
        print("--- forward", it, flush=True)
        fwd.run_fwd({i: x}, [o])


if __name__ == "__main__":
    main()
