#!/usr/bin/env python
"""First-light diagnostics on a B200: run a ladder of ops through the C ABI and print mrd vs the CPU oracle."""
import os, sys, time, traceback
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import boda_b200 as bb
from oracle import boda_oracle as bo

def conv_op_text(N, C, H, W, OC, KH, KW, s, p, relu=1):
    OH, OW = bo.conv_out_sz(H, p, s, KH), bo.conv_out_sz(W, p, s, KW)
    return ("(str_vals=(type=Convolution),nda_vals=(biases=(dims=(out_chan=%d)),filts=(dims=(out_chan=%d,in_chan=%d,y=%d,x=%d)),"
            "in=(dims=(img=%d,chan=%d,y=%d,x=%d)),in_pad=(tn=none,dims=(y=%d,x=%d)),kern_sz=(tn=none,dims=(y=%d,x=%d)),"
            "out=(dims=(img=%d,chan=%d,y=%d,x=%d)),out_chans=(tn=uint32_t,v=%d),stride=(tn=none,dims=(y=%d,x=%d)),conv_has_relu=(tn=uint32_t,v=%d)))"
            % (OC, OC, C, KH, KW, N, C, H, W, p, p, KH, KW, N, OC, OH, OW, OC, s, s, relu))

def run_conv(rtc, tag, N, C, H, W, OC, KH, KW, s, p, relu=1, iters=3):
    txt = conv_op_text(N, C, H, W, OC, KH, KW, s, p, relu)
    op = bo.parse_op(txt)
    ins = bo.gen_op_inputs(op, 5)
    ref = bo.conv_fwd(ins["in"], ins["filts"], ins["biases"], (s, s), (p, p), relu=bool(relu))
    fn = "conv_" + tag
    rtc.compile(fn, txt)
    for k, names in (("in", ["img", "chan", "y", "x"]), ("filts", ["out_chan", "in_chan", "y", "x"]), ("biases", ["out_chan"])):
        rtc.create_var_from_nda(fn + "_" + k, ins[k], names)
    rtc.create_var_with_dims(fn + "_out", list(zip(["img", "chan", "y", "x"], ref.shape)))
    args = {"in": fn + "_in", "filts": fn + "_filts", "biases": fn + "_biases", "out": fn + "_out"}
    ids = [rtc.run(fn, args) for _ in range(iters)]
    rtc.finish_and_sync()
    got = rtc.copy_var_to_nda(fn + "_out")
    ms = min(rtc.get_dur(i, i) for i in ids[1:]) if iters > 1 else rtc.get_dur(ids[0], ids[0])
    fl = bo.op_flops(op)
    m = bo.mrd(ref, got)
    ex = torch.nn.functional.conv2d(torch.from_numpy(ins["in"]).double(), torch.from_numpy(ins["filts"]).double(), torch.from_numpy(ins["biases"]).double(), stride=s, padding=p)
    ex = (ex.clamp_min(0) if relu else ex).numpy()
    print("%-28s mrd=%.3e  gpu-vs-exact=%.3e oracle-vs-exact=%.3e ms=%.4f  TF/s=%.1f  max|ref|=%.1f" % (tag, m, bo.mrd_np(ex, got), bo.mrd_np(ex, ref), ms, fl / ms / 1e9, np.abs(ref).max()), flush=True)
    if not (m < 1e-3):
        d = np.abs(ref.astype(np.float64) - got) / np.maximum(1, np.maximum(np.abs(ref), np.abs(got)))
        ix = np.unravel_index(np.argmax(d), d.shape)
        print("   worst at", ix, "ref", ref[ix], "got", got[ix], " frac_bad=%.4f" % float((d > 1e-3).mean()))
        print("   got[0,0,0,:8]", got[0, 0, 0, :8], "\n   ref[0,0,0,:8]", ref[0, 0, 0, :8])
    for k in ("in", "filts", "biases", "out"):
        rtc.release_var(fn + "_" + k)
    return m

def run_sgemm(rtc, tag, M, N, K, mode=5):
    txt = "(str_vals=(type=sgemm),nda_vals=(a=(dims=(K=%d,M=%d)),b=(dims=(K=%d,N=%d)),c=(dims=(M=%d,N=%d))))" % (K, M, K, N, M, N)
    a, b = bo.gen_sgemm_a(K, M, mode), bo.gen_sgemm_b(K, N, mode)
    ref = bo.sgemm(a, b)
    fn = "sgemm_" + tag
    rtc.compile(fn, txt)
    rtc.create_var_from_nda(fn + "_a", a, ["K", "M"]); rtc.create_var_from_nda(fn + "_b", b, ["K", "N"])
    rtc.create_var_with_dims(fn + "_c", [("M", M), ("N", N)])
    ids = [rtc.run(fn, {"a": fn + "_a", "b": fn + "_b", "c": fn + "_c"}) for _ in range(3)]
    rtc.finish_and_sync()
    got = rtc.copy_var_to_nda(fn + "_c")
    ms = min(rtc.get_dur(i, i) for i in ids[1:])
    m = bo.mrd(ref, got)
    ex = a.astype(np.float64).T @ b.astype(np.float64)
    print("%-28s mrd=%.3e  gpu-vs-exact=%.3e oracle-vs-exact=%.3e ms=%.4f  TF/s=%.1f exact=%s" % (tag, m, bo.mrd_np(ex, got), bo.mrd_np(ex, ref), ms, 2.0 * M * N * K / ms / 1e9, np.array_equal(ref, got)), flush=True)
    if not (m < 1e-3):
        print("   got[:2,:6]", got[:2, :6], "\n   ref[:2,:6]", ref[:2, :6])
    for k in "abc":
        rtc.release_var(fn + "_" + k)
    return m

def main():
    print(bb.lib().b200_version().decode(), "devices:", bb.device_count(), flush=True)
    torch.set_num_threads(os.cpu_count())
    for chunk in (1, 2, 4, 1000000):
        print("==== acc_chunk_kblks =", chunk, flush=True)
        rtc = bb.B200Compute(prec="fp32", acc_chunk_kblks=chunk)
        rtc.init()
        print(rtc.get_plat_tag())
        steps = [
            lambda: run_sgemm(rtc, "s128_%d" % chunk, 128, 128, 128),
            lambda: run_sgemm(rtc, "s128_id_%d" % chunk, 128, 128, 128, 600),
            lambda: run_sgemm(rtc, "s2048_id_%d" % chunk, 2048, 2048, 2048, 600),
            lambda: run_sgemm(rtc, "s2048_%d" % chunk, 2048, 2048, 2048),
            lambda: run_sgemm(rtc, "s_ragged_%d" % chunk, 200, 72, 136),
            lambda: run_conv(rtc, "k1_64to128_%d" % chunk, 2, 64, 12, 12, 128, 1, 1, 1, 0),
            lambda: run_conv(rtc, "k3p1_64to128_%d" % chunk, 2, 64, 13, 13, 128, 3, 3, 1, 1),
            lambda: run_conv(rtc, "k3p1_c32_oc48_%d" % chunk, 3, 32, 9, 11, 48, 3, 3, 1, 1),
            lambda: run_conv(rtc, "k5p2_96to256_%d" % chunk, 5, 96, 27, 27, 256, 5, 5, 1, 2),
            lambda: run_conv(rtc, "k11s4_3to96_%d" % chunk, 2, 3, 227, 227, 96, 11, 11, 4, 0),
            lambda: run_conv(rtc, "fc6_%d" % chunk, 8, 256, 6, 6, 4096, 6, 6, 1, 0),
            lambda: run_conv(rtc, "fc7_%d" % chunk, 8, 4096, 1, 1, 4096, 1, 1, 1, 0),
            lambda: run_conv(rtc, "alex_conv3_b20_%d" % chunk, 20, 256, 13, 13, 384, 3, 3, 1, 1),
            lambda: run_conv(rtc, "alex_conv4_b32_%d" % chunk, 32, 384, 13, 13, 384, 3, 3, 1, 1),
            lambda: run_conv(rtc, "alex_conv2_b32_%d" % chunk, 32, 96, 27, 27, 256, 5, 5, 1, 2),
        ]
        for s in steps:
            try:
                s()
            except Exception as e:
                print("FAILED:", type(e).__name__, e, flush=True)
                traceback.print_exc()
        rtc.close()

if __name__ == "__main__":
    main()
