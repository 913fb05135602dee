"""Helpers shared by the GPU parity tests: drive single ops through the C ABI the way Boda's ops-prof flow does
(src/rtc_prof.cc:44-126: gen_func -> create a var per arg -> gen_data per IN arg -> run -> copy OUTs to host)."""
import itertools

import numpy as np

import boda_b200 as bb

_uid = itertools.count()


def conv_op_text(N, C, H, W, OC, KH, KW, sy=1, sx=1, py=0, px=0, relu=1):
    OH = (H + 2 * py - KH) // sy + 1
    OW = (W + 2 * px - KW) // sx + 1
    return ("(str_vals=(type=Convolution),nda_vals=(biases=(dims=(out_chan=%d)),filts=(dims=(out_chan=%d,in_chan=%d,y=%d,x=%d)),"
            "in=(dims=(img=%d,chan=%d,y=%d,x=%d)),in_pad=(tn=none,dims=(y=%d,x=%d)),kern_sz=(tn=none,dims=(y=%d,x=%d)),"
            "out=(dims=(img=%d,chan=%d,y=%d,x=%d)),out_chans=(tn=uint32_t,v=%d),stride=(tn=none,dims=(y=%d,x=%d)),conv_has_relu=(tn=uint32_t,v=%d)))"
            % (OC, OC, C, KH, KW, N, C, H, W, py, px, KH, KW, N, OC, OH, OW, OC, sy, sx, relu))


def with_relu(op_text, relu=1):
    """Per-op flows force conv_has_relu=1 (src/cnn_op.cc:337); op lists on disk do not carry the flag."""
    if "conv_has_relu" in op_text or "type=Convolution" not in op_text:
        return op_text
    assert op_text.endswith("))")
    return op_text[:-2] + ",conv_has_relu=(tn=uint32_t,v=%d)))" % relu


class OpRunner:
    """One rtc_compute_t instance; vars are named <func>_<arg> and released after each op."""

    def __init__(self, prec="fp32", acc_chunk_kblks=None, **opts):
        self.rtc = bb.B200Compute(prec=prec, acc_chunk_kblks=acc_chunk_kblks, **opts)
        self.rtc.init()

    def close(self):
        self.rtc.close()

    def run_conv(self, op_text, inp, filts, biases, out_shape, iters=1, res=None):
        rtc = self.rtc
        fn = "conv_%d" % next(_uid)
        rtc.compile(fn, op_text)
        try:
            rtc.create_var_from_nda(fn + "_in", inp, ["img", "chan", "y", "x"])
            rtc.create_var_from_nda(fn + "_filts", filts, ["out_chan", "in_chan", "y", "x"])
            rtc.create_var_from_nda(fn + "_biases", biases, ["out_chan"])
            rtc.create_var_with_dims(fn + "_out", list(zip(["img", "chan", "y", "x"], out_shape)))
            args = {"in": fn + "_in", "filts": fn + "_filts", "biases": fn + "_biases", "out": fn + "_out"}
            if res is not None:  # residual join in the epilogue: out = relu?((conv + bias) + res)
                rtc.create_var_from_nda(fn + "_res", res, ["img", "chan", "y", "x"])
                args["res"] = fn + "_res"
            ids = [rtc.run(fn, args) for _ in range(iters)]
            rtc.finish_and_sync()
            out = rtc.copy_var_to_nda(fn + "_out")
            self.last_ms = min(rtc.get_dur(i, i) for i in ids)
            return out
        finally:
            for k in ("in", "filts", "biases", "out", "res"):
                try:
                    rtc.release_var(fn + "_" + k)
                except bb.RtException:
                    pass
            rtc.release_func(fn)
            rtc.release_per_call_id_data()

    def run_conv_gen(self, op_text, mode=5):
        """ops-prof flow with on-device gen_data for every IN arg (src/rtc_prof.cc:73-90)."""
        from oracle import boda_oracle as bo
        rtc = self.rtc
        op = bo.parse_op(op_text)
        fn = "convg_%d" % next(_uid)
        rtc.compile(fn, op_text)
        dims = {"in": ["img", "chan", "y", "x"], "filts": ["out_chan", "in_chan", "y", "x"], "biases": ["out_chan"], "out": ["img", "chan", "y", "x"]}
        try:
            for k, names in dims.items():
                d = op.get_dims(k).dims
                rtc.create_var_with_dims(fn + "_" + k, [(n, d[n]) for n in names])
            for k in ("in", "filts", "biases"):
                d = op.get_dims(k).dims
                gfn = "%s_gen_%s" % (fn, k)
                rtc.compile(gfn, "(str_vals=(type=gen_data,func_name=gen_data_Convolution_%s),nda_vals=(%s=(dims=(%s)),vi=(tn=float,v=0.0),mode=(tn=uint32_t,v=%d)))"
                            % (k, k, ",".join("%s=%d" % (n, d[n]) for n in dims[k]), mode))
                rtc.run(gfn, {k: fn + "_" + k})
            rtc.run(fn, {k: fn + "_" + k for k in dims})
            rtc.finish_and_sync()
            return rtc.copy_var_to_nda(fn + "_out")
        finally:
            for k in dims:
                try:
                    rtc.release_var(fn + "_" + k)
                except bb.RtException:
                    pass
            rtc.release_all_funcs()
            rtc.release_per_call_id_data()

    def run_sgemm(self, a, b, gen_mode=None):
        rtc = self.rtc
        K, M = a.shape
        N = b.shape[1]
        fn = "sgemm_%d" % next(_uid)
        rtc.compile(fn, "(str_vals=(type=sgemm),nda_vals=(a=(dims=(K=%d,M=%d)),b=(dims=(K=%d,N=%d)),c=(dims=(M=%d,N=%d))))" % (K, M, K, N, M, N))
        try:
            if gen_mode is None:
                rtc.create_var_from_nda(fn + "_a", a, ["K", "M"])
                rtc.create_var_from_nda(fn + "_b", b, ["K", "N"])
            else:
                rtc.create_var_with_dims(fn + "_a", [("K", K), ("M", M)])
                rtc.create_var_with_dims(fn + "_b", [("K", K), ("N", N)])
                for k, d in (("a", "K=%d,M=%d" % (K, M)), ("b", "K=%d,N=%d" % (K, N))):
                    rtc.compile(fn + "_gen_" + k, "(str_vals=(type=gen_data,func_name=gen_data_sgemm_%s),nda_vals=(%s=(dims=(%s)),vi=(tn=float,v=0.0),mode=(tn=uint32_t,v=%d)))" % (k, k, d, gen_mode))
                    rtc.run(fn + "_gen_" + k, {k: fn + "_" + k})
            rtc.create_var_with_dims(fn + "_c", [("M", M), ("N", N)])
            cid = rtc.run(fn, {"a": fn + "_a", "b": fn + "_b", "c": fn + "_c"})
            rtc.finish_and_sync()
            self.last_ms = rtc.get_dur(cid, cid)
            return rtc.copy_var_to_nda(fn + "_c")
        finally:
            for k in "abc":
                try:
                    rtc.release_var(fn + "_" + k)
                except bb.RtException:
                    pass
            rtc.release_all_funcs()
            rtc.release_per_call_id_data()

    def run_unary(self, func, op_text, x, out_shape, in_name="in", out_name="out", extra=None):
        rtc = self.rtc
        fn = "%s_%d" % (func, next(_uid))
        rtc.compile(fn, op_text)
        try:
            rtc.create_var_from_nda(fn + "_i", x, ["img", "chan", "y", "x"])
            args = {in_name: fn + "_i"}
            if out_name != in_name:
                rtc.create_var_with_dims(fn + "_o", list(zip(["img", "chan", "y", "x"], out_shape)))
                args[out_name] = fn + "_o"
            args.update(extra or {})
            rtc.run(fn, args)
            rtc.finish_and_sync()
            return rtc.copy_var_to_nda(fn + ("_o" if out_name != in_name else "_i"))
        finally:
            for k in "io":
                try:
                    rtc.release_var(fn + "_" + k)
                except bb.RtException:
                    pass
            rtc.release_func(fn)
            rtc.release_per_call_id_data()


def nchw_dims_text(shape):
    return "dims=(img=%d,chan=%d,y=%d,x=%d)" % tuple(shape)
