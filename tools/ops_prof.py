#!/usr/bin/env python
"""ops-prof for be=b200 (the reference's `boda ops-prof`, src/rtc_prof.cc:139-371, for one back-end): for every op line of --ops-fn
(current or stale syntax) generate the inputs ON DEVICE with gen_data (mode 5), run the op, compare the full output with the CPU oracle
(mrd), and report time, effective TFLOP/s (2*M*N*K, src/latex-util.H:116-133), arithmetic intensity and fraction of the measured
B200 roofline min(peak_TC, AI * HBM_BW).  Timing: CUDA events, >= 3 warm-ups, median of --iters launches (the reference times ONE cold
launch, src/rtc_prof.cc:107,123 -- deliberately not copied).  `kernel_ms` is the contraction kernel alone, `call_ms` adds the operand
packing the call performs (activation transpose/split; filters are packed once per weight version).

  python tools/ops_prof.py --ops-fn ops/c3-conv-ops-small.txt --prec fp16 --out profiles/ops_prof_c3_fp16.md
"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import boda_b200 as bb
from oracle import boda_oracle as bo
from b200_harness import with_relu


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), ("measured" if d else "fallback")


def round_to(a, prec):
    import torch
    if prec == "fp16":
        return torch.from_numpy(a).to(torch.float16).float().numpy()
    if prec == "bf16":
        return torch.from_numpy(a).to(torch.bfloat16).float().numpy()
    return a


def vendor_time_ms(op, prec, iters, warmup):
    """The vendor-library comparator of the reference's `use_culibs=1` leg (cudnn_conv: src/culibs-wrap.cc:94-212, cublas_sgemm: :214-243, dispatch
    src/nvrtc_util.cc:369-371) on the same op in the same run, through torch (cuDNN 9 convolution with autotuning / cuBLASLt GEMM). BENCH
    INFRASTRUCTURE ONLY: nothing in boda_b200/ calls it. Returns {label: median ms} -- for fp32 both the true-fp32 path and the TF32 tensor-core
    path the libraries would pick by default; for fp16 / bf16 the NHWC (channels_last) tensor-core path. Bias + ReLU are applied as the
    library would (separate elementwise kernels inside the timed region for conv; none for sgemm), synthetic inputs."""
    import torch
    import torch.nn.functional as F
    dev = torch.device("cuda")
    out = {}
    torch.backends.cudnn.benchmark = True

    def timed(fn):
        for _ in range(max(warmup, 3)):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    modes = [("fp32", torch.float32, False), ("tf32", torch.float32, True)] if prec == "fp32" else [(prec, torch.float16 if prec == "fp16" else torch.bfloat16, False)]
    for label, dt, tf32 in modes:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        if op.type == "Convolution":
            i, f = op.get_dims("in").dims, op.get_dims("filts").dims
            x = torch.randn(i["img"], i["chan"], i["y"], i["x"], device=dev, dtype=dt)
            w = torch.randn(f["out_chan"], f["in_chan"], f["y"], f["x"], device=dev, dtype=dt)
            b = torch.randn(f["out_chan"], device=dev, dtype=dt)
            if dt != torch.float32:
                x, w = x.contiguous(memory_format=torch.channels_last), w.contiguous(memory_format=torch.channels_last)
            st, pd = op.pt("stride", (1, 1)), op.pt("in_pad", (0, 0))
            out["cudnn_" + label] = timed(lambda: F.relu_(F.conv2d(x, w, b, stride=st, padding=pd)))
            out["cudnn_" + label + "_conv_only"] = timed(lambda: F.conv2d(x, w, None, stride=st, padding=pd))
        else:
            a, bb_ = op.get_dims("a").dims, op.get_dims("b").dims
            A = torch.randn(a["K"], a["M"], device=dev, dtype=dt)
            Bm = torch.randn(bb_["K"], bb_["N"], device=dev, dtype=dt)
            out["cublas_" + label] = timed(lambda: torch.matmul(A.t(), Bm))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--compare", action="store_true", help="also time the vendor libraries (cuDNN conv / cuBLASLt GEMM through torch) on the same ops in the same run: "
                    "the reference's use_culibs=1 comparator (src/culibs-wrap.cc). Bench infrastructure only.")
    ap.add_argument("--ops-fn", required=True)
    ap.add_argument("--prec", default="fp32", choices=["fp32", "fp16", "bf16"])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--wisdom-out", default="", help="also write the results as a Boda wisdom file (op_wisdom_t records with nda_digest known-good vectors, "
                    "src/op-tuner.cc:98-130) that Boda's wis-ana / ops-prof flows can read")
    ap.add_argument("--opts", default="", help="extra back-end options, e.g. use_taps=0,taps_2cta=1,acc_chunk_kblks_16=4")
    ap.add_argument("--max-check-gflop", type=float, default=40.0, help="skip the CPU oracle compare for ops larger than this")
    args = ap.parse_args()
    peak_tf, hbm_gbs, src = peaks()
    rtc = bb.B200Compute(prec=args.prec, **dict(kv.split("=") for kv in args.opts.split(",") if kv))
    rtc.init()
    rows = []
    for li, line in enumerate(open(args.ops_fn)):
        line = line.strip()
        if not line:
            continue
        op = bo.parse_op(line)
        fn = "op%d" % li
        txt = with_relu(line) if "dims_vals" not in line else line  # stale lines carry no nda_vals list to extend: relu flag via arg below
        if op.type == "Convolution":
            if "dims_vals" in line:  # translate to the current syntax with conv_has_relu=1 (per-op flows force it, src/cnn_op.cc:337)
                d = {k: op.get_dims(k).dims for k in ("in", "filts", "out")}
                from b200_harness import conv_op_text
                i, f = d["in"], d["filts"]
                txt = conv_op_text(i["img"], i["chan"], i["y"], i["x"], f["out_chan"], f["y"], f["x"], *op.pt("stride", (1, 1)), *op.pt("in_pad", (0, 0)), 1)
            names = {"in": ["img", "chan", "y", "x"], "filts": ["out_chan", "in_chan", "y", "x"], "biases": ["out_chan"], "out": ["img", "chan", "y", "x"]}
            outs = ["out"]
        else:
            names = {"a": ["K", "M"], "b": ["K", "N"], "c": ["M", "N"]}
            outs = ["c"]
            txt = "(str_vals=(type=sgemm),nda_vals=(a=(dims=(K=%d,M=%d)),b=(dims=(K=%d,N=%d)),c=(dims=(M=%d,N=%d))))" % (
                op.get_dims("a").dims["K"], op.get_dims("a").dims["M"], op.get_dims("b").dims["K"], op.get_dims("b").dims["N"], op.get_dims("c").dims["M"], op.get_dims("c").dims["N"])
        rtc.compile(fn, txt)
        for k, dn in names.items():
            d = op.get_dims(k).dims
            rtc.create_var_with_dims(fn + "_" + k, [(n, d[n]) for n in dn])
        ins = [k for k in names if k not in outs]
        flops = bo.op_flops(op)
        do_check = (not args.no_check) and flops / 1e9 <= args.max_check_gflop
        host = {}
        if args.prec == "fp32":
            for k in ins:  # device-side gen_data, exactly the reference flow (src/rtc_prof.cc:73-90)
                d = op.get_dims(k).dims
                g = "%s_gen_%s" % (fn, k)
                rtc.compile(g, "(str_vals=(type=gen_data,func_name=gen_data_%s_%s),nda_vals=(%s=(dims=(%s)),vi=(tn=float,v=0.0),mode=(tn=uint32_t,v=5)))"
                            % (op.type, k, k, ",".join("%s=%d" % (n, d[n]) for n in names[k])))
                rtc.run(g, {k: fn + "_" + k})
            if do_check:
                host = bo.gen_op_inputs(op, 5)
        else:  # fp16 / bf16 storage: same hash values rounded to the storage type on the host, oracle fed the rounded values
            host = {k: (round_to(v, args.prec) if k != "biases" else v) for k, v in bo.gen_op_inputs(op, 5).items()}
            for k in ins:
                rtc.copy_nda_to_var(fn + "_" + k, host[k])
        amap = {k: fn + "_" + k for k in names}
        for _ in range(args.warmup):
            rtc.run(fn, amap)
        rtc.finish_and_sync()
        rtc.release_per_call_id_data()
        ids = []
        for _ in range(args.iters):
            if ins:  # mark the activation operand as rewritten so every timed call repacks it, as a forward pass would
                rtc.get_var_raw_native_pointer(fn + "_" + ins[0])
            ids.append(rtc.run(fn, amap))
        rtc.finish_and_sync()
        call_ms = float(np.median([rtc.get_dur(i, i) for i in ids]))
        kern_ms = float(np.median([rtc.get_kernel_dur(i) for i in ids]))
        m = float("nan")
        if args.wisdom_out:  # known-good digest of the output actually computed + the timing, in the reference's own record format
            outv = rtc.copy_var_to_nda(fn + "_" + outs[0])
            rec = bb.wisdom_record(line, [(outs[0], bb.nda_digest_hex(outs[0], outv, names[outs[0]]))], "(use_be=b200,prec=%s)" % args.prec,
                                   rtc.get_plat_tag(), kern_ms * 1e-3, run_op_text=txt)
            with open(args.wisdom_out, "a" if li else "w") as wf:
                wf.write(rec)
        if do_check:
            got = rtc.copy_var_to_nda(fn + "_" + outs[0])
            ref = bo.run_op(op, host, acc64=True)[outs[0]]
            m = bo.mrd(ref, got)
        esz = 4  # algorithmic bytes use the op's declared element type: fp32 tensors at the boundary (src/latex-util.H:119,133)
        nbytes = esz * sum(int(np.prod(op.get_dims(k).shape())) for k in names)
        ai = flops / nbytes
        roof_tf = min(peak_tf, ai * hbm_gbs / 1e3)
        tf_k, tf_c = flops / kern_ms / 1e9, flops / call_ms / 1e9
        desc = line[:0]
        if op.type == "Convolution":
            i, f, o = op.get_dims("in").dims, op.get_dims("filts").dims, op.get_dims("out").dims
            desc = "conv %dx%d/%d/%d %d->%d @%dx%d B=%d" % (f["y"], f["x"], op.pt("stride", (1, 1))[0], op.pt("in_pad", (0, 0))[0], i["chan"], f["out_chan"], i["y"], i["x"], i["img"])
        else:
            a = op.get_dims("a").dims
            desc = "sgemm M=%d N=%d K=%d" % (a["M"], op.get_dims("b").dims["N"], a["K"])
        vend = vendor_time_ms(op, args.prec, args.iters, args.warmup) if args.compare else {}
        rows.append(dict(op=desc, gflop=flops / 1e9, mbytes=nbytes / 1e6, ai=ai, kernel_ms=kern_ms, call_ms=call_ms, tflops_kernel=tf_k, tflops_call=tf_c,
                         roof_tflops=roof_tf, frac_roof=tf_k / roof_tf, frac_tc_peak=tf_k / peak_tf, mrd=m, vendor_ms=vend,
                         vendor_tflops={k: flops / v / 1e9 for k, v in vend.items()}))
        if vend:
            print("    vendor: " + "  ".join("%s %.4f ms %.1f TF/s (b200 call / vendor = %.2fx)" % (k, v, flops / v / 1e9, call_ms / v) for k, v in vend.items()), flush=True)
        print("%-44s %7.2f GF  AI %6.0f  kernel %8.4f ms %7.1f TF/s  call %8.4f ms %7.1f TF/s  roof %6.0f  frac %.3f  mrd %.2e" %
              (desc, flops / 1e9, ai, kern_ms, tf_k, call_ms, tf_c, roof_tf, tf_k / roof_tf, m), flush=True)
        for k in names:
            rtc.release_var(fn + "_" + k)
        rtc.release_all_funcs()
        rtc.release_per_call_id_data()
    if args.out:
        with open(args.out, "w") as f:
            f.write("# ops-prof, be=b200, prec=%s, %s\n\n" % (args.prec, os.path.basename(args.ops_fn)))
            f.write("Peaks (%s): tensor %.1f TFLOP/s (bf16 cuBLAS burst), HBM %.0f GB/s. Roofline = min(peak, AI x HBM). Inputs: reference gen_data mode 5%s. "
                    "Timing: CUDA events, median of %d after %d warm-ups; `kernel` = contraction kernel alone, `call` = with per-call operand packing. "
                    "mrd = max rel diff (floor 1) vs the CPU oracle (acc64).\n\n" % (src, peak_tf, hbm_gbs, "" if args.prec == "fp32" else " rounded to " + args.prec, args.iters, args.warmup))
            f.write("| op | GFLOP | MB | AI F/B | kernel ms | kernel TF/s | call ms | call TF/s | roofline TF/s | kernel / roofline | kernel / TC peak | mrd |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %.2f | %.1f | %.0f | %.4f | %.1f | %.4f | %.1f | %.0f | %.3f | %.3f | %.1e |\n" %
                        (r["op"], r["gflop"], r["mbytes"], r["ai"], r["kernel_ms"], r["tflops_kernel"], r["call_ms"], r["tflops_call"], r["roof_tflops"], r["frac_roof"], r["frac_tc_peak"], r["mrd"]))
            if args.compare and rows:
                keys = list(rows[0]["vendor_ms"].keys())
                f.write("\nVendor comparator (the reference's `use_culibs=1` leg, src/culibs-wrap.cc:94-243): the same ops in the same run through torch -- cuDNN 9 convolution "
                        "(autotuned; `_conv_only` = without the bias + ReLU kernels) / cuBLASLt GEMM; for fp32 both true fp32 and the TF32 tensor-core path, for fp16 / bf16 the "
                        "NHWC tensor-core path on 16-bit tensors (no fp32 NCHW boundary, no operand packing: the library's best case). Bench infrastructure only. "
                        "ratio = this back-end's CALL time (packing + kernel, fp32 NCHW in and out) / vendor time; < 1 = faster than the library.\n\n")
                f.write("| op | b200 call ms | " + " | ".join("%s ms | ratio" % k for k in keys) + " |\n|---|---|" + "---|---|" * len(keys) + "\n")
                for r in rows:
                    f.write("| %s | %.4f | " % (r["op"], r["call_ms"]) + " | ".join("%.4f | %.2f" % (r["vendor_ms"][k], r["call_ms"] / r["vendor_ms"][k]) for k in keys) + " |\n")
        with open(os.path.splitext(args.out)[0] + ".json", "w") as f:
            json.dump(rows, f, indent=1)
    rtc.close()


if __name__ == "__main__":
    main()
