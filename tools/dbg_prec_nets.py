import sys, re; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np
import boda_b200 as bb
from boda_b200 import nets
from oracle import net_oracle, boda_oracle as bo
for prec in ("bf16","fp16"):
    txt, i, o = nets.googlenet_conv(2)
    params = nets.synth_params(txt); x = nets.synth_input((2,3,224,224))
    names=[]
    for line in txt.splitlines():
        m = re.search(r"tops=([^,)]+)", line)
        if m:
            for t in m.group(1).split(":"):
                if t not in names: names.append(t)
    f = bb.B200ConvFwd(txt, "(prec=%s)"%prec)
    for k,v in params.items(): f.set_param(k,v)
    got = f.run_fwd({i:x}, names)
    ref = net_oracle.run_pipe(txt, {i:x}, params, round_to=("bf16" if prec=="bf16" else np.float16), acc64=True)
    bad=0
    for n in names:
        m = bo.mrd(ref[n], got[n])
        if m > 2e-2 or n in ("conv1","pool1","cls3_fc"):
            print(prec, n, got[n].shape, "mrd %.3e"%m, "max|ref| %.2f"%np.abs(ref[n]).max()); bad+=1
            if bad>8: break
