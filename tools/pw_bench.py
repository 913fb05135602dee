#!/usr/bin/env python
"""Micro-benchmark of the bandwidth-bound functions (pool / lrn / pack via a 1x1 conv) at the AlexNet-ng B=32 shapes, with and without the
abs-max side channel. Meant to run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` (launch list)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import boda_b200 as bb
from b200_harness import nchw_dims_text

rtc = bb.B200Compute()
rtc.init()
rtc.create_var_with_dims("cells", [("cell", 4)], tn="uint32_t")
rng = np.random.RandomState(0)


def pool(tag, shape, k, s, reps=3, absmax=False):
    N, C, H, W = shape
    OH, OW = -(-(H - k) // s) + 1, -(-(W - k) // s) + 1
    out = (N, C, OH, OW)
    txt = "(str_vals=(type=Pooling),nda_vals=(avg_pool=(tn=uint32_t,v=0),kern_sz=(tn=none,dims=(y=%d,x=%d)),stride=(tn=none,dims=(y=%d,x=%d)),in_pad=(tn=none,dims=(y=0,x=0)),in=(%s),out=(%s)))" % (
        k, k, s, s, nchw_dims_text(shape), nchw_dims_text(out))
    rtc.compile(tag, txt)
    rtc.create_var_from_nda(tag + "_i", rng.randn(*shape).astype(np.float32), ["img", "chan", "y", "x"])
    rtc.create_var_with_dims(tag + "_o", list(zip(["img", "chan", "y", "x"], out)))
    args = {"in": tag + "_i", "out": tag + "_o"}
    if absmax:
        args.update({"out_absmax_cells": "cells", "out_absmax_ix": 0})
    ids = [rtc.run(tag, args) for _ in range(reps)]
    rtc.finish_and_sync()
    print("%-28s %s -> %s  %.1f us (events, min of %d)" % (tag, shape, out, 1e3 * min(rtc.get_dur(i, i) for i in ids), reps), flush=True)
    rtc.release_var(tag + "_i"); rtc.release_var(tag + "_o")


def lrn(tag, shape, reps=3, absmax=False):
    txt = "(str_vals=(type=LRN),nda_vals=(local_size=(tn=uint32_t,v=5),alpha=(tn=float,v=0.0001),beta=(tn=float,v=0.75),k=(tn=float,v=1.0),in=(%s),out=(%s)))" % (nchw_dims_text(shape), nchw_dims_text(shape))
    rtc.compile(tag, txt)
    rtc.create_var_from_nda(tag + "_i", rng.randn(*shape).astype(np.float32), ["img", "chan", "y", "x"])
    rtc.create_var_with_dims(tag + "_o", list(zip(["img", "chan", "y", "x"], shape)))
    args = {"in": tag + "_i", "out": tag + "_o"}
    if absmax:
        args.update({"out_absmax_cells": "cells", "out_absmax_ix": 1})
    ids = [rtc.run(tag, args) for _ in range(reps)]
    rtc.finish_and_sync()
    print("%-28s %s  %.1f us (events, min of %d)" % (tag, shape, 1e3 * min(rtc.get_dur(i, i) for i in ids), reps), flush=True)
    rtc.release_var(tag + "_i"); rtc.release_var(tag + "_o")


for am in (False, True):
    sfx = "_absmax" if am else ""
    pool("pool1" + sfx, (32, 96, 55, 55), 3, 2, absmax=am)
    pool("pool2" + sfx, (32, 256, 27, 27), 3, 2, absmax=am)
    pool("pool5" + sfx, (32, 256, 13, 13), 3, 2, absmax=am)
    lrn("lrn1" + sfx, (32, 96, 55, 55), absmax=am)
    lrn("lrn2" + sfx, (32, 256, 27, 27), absmax=am)
rtc.close()
