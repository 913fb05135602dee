#!/usr/bin/env python
"""Unique Convolution signatures of a net as op lines for tools/ops_prof.py --ops-fn (the reference's write_op_sigs, src/rtc_fwd.cc:246-264).
usage: net_op_sigs.py NET BATCH [IN_SZ] > ops.txt        NET in boda_b200.nets.NETS; host-only, no GPU"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import boda_b200 as bb
from boda_b200 import nets

if __name__ == "__main__":
    net, batch = sys.argv[1], int(sys.argv[2])
    txt = nets.NETS[net](batch, int(sys.argv[3]))[0] if len(sys.argv) > 3 else nets.NETS[net](batch)[0]
    print("\n".join(bb.pipe_op_sigs(txt)))
