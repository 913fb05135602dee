"""Net descriptions in Boda's op-line text form (one conv_op_t per line) for has_conv_fwd_t::init, plus synthetic
parameter generation.

The reference builds its conv_pipe from Caffe prototxts (src/caffepb.cc:166-326); no prototxt or caffemodel travels
with this repo, so the architectures of the BASELINE configs are restated here from nets/<name>/train_val.prototxt
(TEST phase: Data/Accuracy/SoftmaxWithLoss layers dropped, src/caffepb.cc:250-261; Dropout kept as an in-place
identity). Weights are synthetic: the reference's deterministic hash (test/rtc/gen-util.h) scaled He-style so
activations stay O(input) through the net (no caffemodel ships with the reference either, INSTALL.md:199-209).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


class PipeBuilder:
    def __init__(self):
        self.lines: List[str] = []
        self.convs: List[Tuple[str, int, int, int, int]] = []  # tag, OC, C, KH, KW (filled lazily by param_shapes)

    def data(self, name: str, img: int, chan: int, y: int, x: int):
        self.lines.append("(node=%s,dims=(img=%d,chan=%d,y=%d,x=%d))" % (name, img, chan, y, x))
        return name

    def conv(self, tag: str, bot: str, top: str, out_chans: int, k, stride=1, pad=0, relu: Optional[str] = None, bias: bool = True):
        ky, kx = (k, k) if isinstance(k, int) else k
        self.lines.append(
            "(tag=%s,str_vals=(type=Convolution),nda_vals=(kern_sz=(tn=none,dims=(y=%d,x=%d)),stride=(tn=none,dims=(y=%d,x=%d)),"
            "in_pad=(tn=none,dims=(y=%d,x=%d)),out_chans=(tn=uint32_t,v=%d)%s),bots=%s,tops=%s)"
            % (tag, ky, kx, stride, stride, pad, pad, out_chans, "" if bias else ",bias_term=(tn=uint32_t,v=0)", bot, top))
        if relu:
            self.relu(relu, top)
        return top

    def batch_norm(self, tag: str, node: str, eps: float = 1e-5):
        """Caffe BatchNorm with use_global_stats (in place): params <tag>_mean, <tag>_var (chan) and <tag>_sf (1)."""
        self.lines.append("(tag=%s,str_vals=(type=BatchNorm),nda_vals=(eps=(tn=float,v=%r)),bots=%s,tops=%s)" % (tag, float(eps), node, node))
        return node

    def scale(self, tag: str, node: str):
        """Caffe Scale with bias_term (in place): params <tag>_gamma, <tag>_beta (chan)."""
        self.lines.append("(tag=%s,str_vals=(type=Scale),bots=%s,tops=%s)" % (tag, node, node))
        return node

    def inner_product(self, tag: str, bot: str, top: str, out_chans: int):
        self.lines.append("(tag=%s,str_vals=(type=InnerProduct),nda_vals=(out_chans=(tn=uint32_t,v=%d)),bots=%s,tops=%s)" % (tag, out_chans, bot, top))
        return top

    def relu(self, tag: str, node: str):
        self.lines.append("(tag=%s,str_vals=(type=ReLU),bots=%s,tops=%s)" % (tag, node, node))
        return node

    def dropout(self, tag: str, node: str, ratio: float = 0.5):
        self.lines.append("(tag=%s,str_vals=(type=Dropout),nda_vals=(dropout_ratio=(tn=float,v=%g)),bots=%s,tops=%s)" % (tag, ratio, node, node))
        return node

    def lrn(self, tag: str, bot: str, top: str, local_size=5, alpha=1e-4, beta=0.75, k=1.0):
        self.lines.append("(tag=%s,str_vals=(type=LRN),nda_vals=(local_size=(tn=uint32_t,v=%d),alpha=(tn=float,v=%r),beta=(tn=float,v=%r),k=(tn=float,v=%r)),bots=%s,tops=%s)"
                          % (tag, local_size, float(alpha), float(beta), float(k), bot, top))
        return top

    def pool(self, tag: str, bot: str, top: str, k: Optional[int], stride=1, pad=0, avg=False):
        if k is None:  # global pooling
            nda = "avg_pool=(tn=uint32_t,v=%d)" % (1 if avg else 0)
        else:
            nda = ("avg_pool=(tn=uint32_t,v=%d),kern_sz=(tn=none,dims=(y=%d,x=%d)),stride=(tn=none,dims=(y=%d,x=%d)),in_pad=(tn=none,dims=(y=%d,x=%d))"
                   % (1 if avg else 0, k, k, stride, stride, pad, pad))
        self.lines.append("(tag=%s,str_vals=(type=Pooling),nda_vals=(%s),bots=%s,tops=%s)" % (tag, nda, bot, top))
        return top

    def concat(self, tag: str, bots: List[str], top: str):
        self.lines.append("(tag=%s,str_vals=(type=Concat),bots=%s,tops=%s)" % (tag, ":".join(bots), top))
        return top

    def eltwise(self, tag: str, bots: List[str], top: str):
        self.lines.append("(tag=%s,str_vals=(type=Eltwise),bots=%s,tops=%s)" % (tag, ":".join(bots), top))
        return top

    def softmax(self, tag: str, bot: str, top: str):
        self.lines.append("(tag=%s,str_vals=(type=Softmax),bots=%s,tops=%s)" % (tag, bot, top))
        return top

    def text(self) -> str:
        return "\n".join(self.lines) + "\n"


def alexnet_ng_conv(batch: int = 32, in_sz: int = 227) -> Tuple[str, str, str]:
    """nets/alexnet_ng_conv/train_val.prototxt (BASELINE config C2). Returns (pipe_text, input node, output node)."""
    p = PipeBuilder()
    p.data("data", batch, 3, in_sz, in_sz)
    p.conv("conv1", "data", "conv1", 96, 11, 4, 0, relu="relu1")
    p.lrn("norm1", "conv1", "norm1")
    p.pool("pool1", "norm1", "pool1", 3, 2)
    p.conv("conv2", "pool1", "conv2", 256, 5, 1, 2, relu="relu2")
    p.lrn("norm2", "conv2", "norm2")
    p.pool("pool2", "norm2", "pool2", 3, 2)
    p.conv("conv3", "pool2", "conv3", 384, 3, 1, 1, relu="relu3")
    p.conv("conv4", "conv3", "conv4", 384, 3, 1, 1, relu="relu4")
    p.conv("conv5", "conv4", "conv5", 256, 3, 1, 1, relu="relu5")
    p.pool("pool5", "conv5", "pool5", 3, 2)
    p.conv("fc6-conv", "pool5", "fc6", 4096, 6, 1, 0, relu="relu6")
    p.dropout("drop6", "fc6")
    p.conv("fc7-conv", "fc6", "fc7", 4096, 1, 1, 0, relu="relu7")
    p.dropout("drop7", "fc7")
    p.conv("fc8-conv", "fc7", "fc8", 1000, 1, 1, 0)
    return p.text(), "data", "fc8"


def nin_imagenet(batch: int = 32, in_sz: int = 227) -> Tuple[str, str, str]:
    """nets/nin_imagenet/train_val.prototxt: 4 x (kxk conv + two 1x1 cccp convs), max pools, 6x6 average pool."""
    p = PipeBuilder()
    p.data("data", batch, 3, in_sz, in_sz)
    p.conv("conv1", "data", "conv1", 96, 11, 4, 0, relu="relu0")
    p.conv("cccp1", "conv1", "cccp1", 96, 1, relu="relu1")
    p.conv("cccp2", "cccp1", "cccp2", 96, 1, relu="relu2")
    p.pool("pool0", "cccp2", "pool0", 3, 2)
    p.conv("conv2", "pool0", "conv2", 256, 5, 1, 2, relu="relu3")
    p.conv("cccp3", "conv2", "cccp3", 256, 1, relu="relu5")
    p.conv("cccp4", "cccp3", "cccp4", 256, 1, relu="relu6")
    p.pool("pool2", "cccp4", "pool2", 3, 2)
    p.conv("conv3", "pool2", "conv3", 384, 3, 1, 1, relu="relu7")
    p.conv("cccp5", "conv3", "cccp5", 384, 1, relu="relu8")
    p.conv("cccp6", "cccp5", "cccp6", 384, 1, relu="relu9")
    p.pool("pool3", "cccp6", "pool3", 3, 2)
    p.dropout("drop", "pool3")
    p.conv("conv4-1024", "pool3", "conv4", 1024, 3, 1, 1, relu="relu10")
    p.conv("cccp7-1024", "conv4", "cccp7", 1024, 1, relu="relu11")
    p.conv("cccp8-1024", "cccp7", "cccp8", 1000, 1, relu="relu12")
    p.pool("pool4", "cccp8", "pool4", 6, 1, 0, avg=True)
    return p.text(), "data", "pool4"


def googlenet_conv(batch: int = 64, in_sz: int = 224) -> Tuple[str, str, str]:
    """nets/googlenet_conv/train_val.prototxt (BASELINE config C4): 7x7/2 stem, 9 inception modules (4-way Concat), LRN after pool1 and
    conv2, the two auxiliary classifiers (all present in the TEST graph) and the 7x7 average pool + 1x1 classifier. 64 Convolution layers.
    Returns (pipe_text, input node, main output node `cls3_fc`); the auxiliary outputs are `cls1_fc2` and `cls2_fc2`."""
    p = PipeBuilder()
    p.data("data", batch, 3, in_sz, in_sz)
    p.conv("conv1", "data", "conv1", 64, 7, 2, 3, relu="relu1")
    p.pool("pool1", "conv1", "pool1", 3, 2)
    p.lrn("norm1", "pool1", "norm1")
    p.conv("reduction2", "norm1", "reduction2", 64, 1, relu="relu_reduction2")
    p.conv("conv2", "reduction2", "conv2", 192, 3, 1, 1, relu="relu2")
    p.lrn("norm2", "conv2", "norm2")
    p.pool("pool2", "norm2", "pool2", 3, 2)

    def inception(n, bot, top, r1, r2, o0, o1, o2, o3):
        pre = "icp%d_" % n
        p.conv(pre + "reduction1", bot, pre + "reduction1", r1, 1, relu="relu_" + pre + "reduction1")
        p.conv(pre + "reduction2", bot, pre + "reduction2", r2, 1, relu="relu_" + pre + "reduction2")
        p.pool(pre + "pool", bot, pre + "pool", 3, 1, 1)
        p.conv(pre + "out0", bot, pre + "out0", o0, 1, relu="relu_" + pre + "out0")
        p.conv(pre + "out1", pre + "reduction1", pre + "out1", o1, 3, 1, 1, relu="relu_" + pre + "out1")
        p.conv(pre + "out2", pre + "reduction2", pre + "out2", o2, 5, 1, 2, relu="relu_" + pre + "out2")
        p.conv(pre + "out3", pre + "pool", pre + "out3", o3, 1, relu="relu_" + pre + "out3")
        p.concat(top, [pre + "out0", pre + "out1", pre + "out2", pre + "out3"], top)
        return top

    def aux(n, bot):
        pre = "cls%d_" % n
        p.pool(pre + "pool", bot, pre + "pool", 5, 3, 0, avg=True)
        p.conv(pre + "reduction", pre + "pool", pre + "reduction", 128, 1, relu="relu_" + pre + "reduction")
        p.conv(pre + "fc1-conv", pre + "reduction", pre + "fc1", 1024, 4, 1, 0, relu="relu_" + pre + "fc1")
        p.dropout(pre + "drop", pre + "fc1", 0.7)
        p.conv(pre + "fc2-conv", pre + "fc1", pre + "fc2", 1000, 1)

    inception(1, "pool2", "icp2_in", 96, 16, 64, 128, 32, 32)
    inception(2, "icp2_in", "icp2_out", 128, 32, 128, 192, 96, 64)
    p.pool("icp3_in", "icp2_out", "icp3_in", 3, 2)
    inception(3, "icp3_in", "icp3_out", 96, 16, 192, 208, 48, 64)
    aux(1, "icp3_out")
    inception(4, "icp3_out", "icp4_out", 112, 24, 160, 224, 64, 64)
    inception(5, "icp4_out", "icp5_out", 128, 24, 128, 256, 64, 64)
    inception(6, "icp5_out", "icp6_out", 144, 32, 112, 288, 64, 64)
    aux(2, "icp6_out")
    inception(7, "icp6_out", "icp7_out", 160, 32, 256, 320, 128, 128)
    p.pool("icp8_in", "icp7_out", "icp8_in", 3, 2)
    inception(8, "icp8_in", "icp8_out", 160, 32, 256, 320, 128, 128)
    inception(9, "icp8_out", "icp9_out", 192, 48, 384, 384, 128, 128)
    p.pool("cls3_pool", "icp9_out", "cls3_pool", 7, 1, 0, avg=True)
    p.dropout("cls3_drop", "cls3_pool", 0.4)
    p.conv("cls3_fc-conv", "cls3_pool", "cls3_fc", 1000, 1)
    return p.text(), "data", "cls3_fc"


def resnet50(batch: int = 32, in_sz: int = 224, stages=(3, 4, 6, 3)) -> Tuple[str, str, str]:
    """nets/resnet-50/train_val.prototxt (BASELINE config C5): 7x7/2 stem, max pool, 16 bottleneck blocks (1x1 -> 3x3 -> 1x1, stride 2 on the
    first 1x1 of stages 3-5, projection shortcut on the first block of every stage), every Convolution (bias_term false except conv1)
    followed in place by BatchNorm(use_global_stats) + Scale(bias_term), Eltwise SUM + ReLU joins, 7x7 average pool, InnerProduct(1000), Softmax.
    The reference cannot run this net (its BatchNorm / Scale / Eltwise readers are stubs, src/caffepb.cc:231-232,308); SURVEY section 8 f4."""
    p = PipeBuilder()
    p.data("data", batch, 3, in_sz, in_sz)

    def cbs(name, bn_suffix, bot, oc, k, stride=1, pad=0, relu=True, bias=False):
        p.conv(name, bot, name, oc, k, stride, pad, bias=bias)
        p.batch_norm("bn" + bn_suffix, name)
        p.scale("scale" + bn_suffix, name)
        if relu:
            p.relu(name + "_relu", name)
        return name

    cbs("conv1", "_conv1", "data", 64, 7, 2, 3, bias=True)
    p.pool("pool1", "conv1", "pool1", 3, 2)
    prev, width = "pool1", 64
    for si, n_blocks in enumerate(stages):
        stage = si + 2
        for b in range(n_blocks):
            blk = "%d%s" % (stage, "abcdefgh"[b])
            stride = 2 if (b == 0 and stage > 2) else 1
            if b == 0:
                short = cbs("res%s_branch1" % blk, blk + "_branch1", prev, width * 4, 1, stride, 0, relu=False)
            else:
                short = prev
            x = cbs("res%s_branch2a" % blk, blk + "_branch2a", prev, width, 1, stride, 0)
            x = cbs("res%s_branch2b" % blk, blk + "_branch2b", x, width, 3, 1, 1)
            x = cbs("res%s_branch2c" % blk, blk + "_branch2c", x, width * 4, 1, 1, 0, relu=False)
            p.eltwise("res%s" % blk, [short, x], "res%s" % blk)
            p.relu("res%s_relu" % blk, "res%s" % blk)
            prev = "res%s" % blk
        width *= 2
    p.pool("pool5", prev, "pool5", 7, 1, 0, avg=True)
    p.inner_product("fc1000", "pool5", "fc1000", 1000)
    p.softmax("prob", "fc1000", "prob")
    return p.text(), "data", "prob"


def tiny_resnet(batch: int = 2) -> Tuple[str, str, str]:
    """Two-stage miniature of resnet50 (same op kinds and wiring, 33x33 input) small enough for the CPU oracle."""
    p = PipeBuilder()
    p.data("data", batch, 3, 33, 33)

    def cbs(name, bot, oc, k, stride=1, pad=0, relu=True, bias=False, bn=True, sc=True):
        p.conv(name, bot, name, oc, k, stride, pad, bias=bias)
        if bn:
            p.batch_norm("bn_" + name, name)
        if sc:
            p.scale("scale_" + name, name)
        if relu:
            p.relu(name + "_relu", name)
        return name

    cbs("conv1", "data", 16, 7, 2, 3, bias=True)
    p.pool("pool1", "conv1", "pool1", 3, 2)
    prev = "pool1"
    for blk, (width, stride, proj) in {"2a": (8, 1, True), "2b": (8, 1, False), "3a": (16, 2, True)}.items():
        short = cbs("res%s_branch1" % blk, prev, width * 4, 1, stride, 0, relu=False) if proj else prev
        x = cbs("res%s_branch2a" % blk, prev, width, 1, stride, 0)
        x = cbs("res%s_branch2b" % blk, x, width, 3, 1, 1, bn=(blk != "2b"))       # a Scale-only fold
        x = cbs("res%s_branch2c" % blk, x, width * 4, 1, 1, 0, relu=False, sc=(blk != "3a"))  # a BatchNorm-only fold
        p.eltwise("res%s" % blk, [short, x], "res%s" % blk)
        p.relu("res%s_relu" % blk, "res%s" % blk)
        prev = "res%s" % blk
    p.pool("pool5", prev, "pool5", None, avg=True)
    p.inner_product("fc", "pool5", "fc", 10)
    p.softmax("prob", "fc", "prob")
    return p.text(), "data", "prob"


def tiny_net(batch: int = 3) -> Tuple[str, str, str]:
    """A small net touching every forward op kind (conv variants, LRN, max/avg/global pool, concat, eltwise, softmax)."""
    p = PipeBuilder()
    p.data("data", batch, 3, 31, 29)
    p.conv("c1", "data", "c1", 24, 5, 2, 1, relu="r1")
    p.lrn("n1", "c1", "n1", 5, 1e-2, 0.75, 1.0)
    p.pool("p1", "n1", "p1", 3, 2)
    p.conv("b1", "p1", "b1", 16, 1, relu="rb1")
    p.conv("b2", "p1", "b2", 40, 3, 1, 1, relu="rb2")
    p.pool("b3p", "p1", "b3p", 3, 1, 1)
    p.conv("b3", "b3p", "b3", 8, 1, relu="rb3")
    p.concat("cat", ["b1", "b2", "b3"], "cat")
    p.conv("c2", "cat", "c2", 64, 3, 1, 1)
    p.conv("c2b", "cat", "c2b", 64, 1, 1, 0)
    p.eltwise("sum", ["c2", "c2b"], "sum")
    p.relu("rsum", "sum")
    p.pool("pa", "sum", "pa", 2, 2, 0, avg=True)
    p.conv("fc", "pa", "fc", 10, (3, 3), 1, 0)
    p.pool("gp", "fc", "gp", None, avg=True)
    p.softmax("prob", "gp", "prob")
    return p.text(), "data", "prob"


# ---- synthetic parameters ----------------------------------------------------------------------------------------

def _det_hash_rand(ix: np.ndarray) -> np.ndarray:
    """The reference's input hash (test/rtc/gen-util.h:1-9), vectorised: uniform in [-5, 5]."""
    h = ix.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    return (h.astype(np.float32).astype(np.float64) * (10.0 / 4294967296.0) - 5.0).astype(np.float32)


def hash_fill(shape, salt: int, scale: float) -> np.ndarray:
    n = int(np.prod(shape))
    return (_det_hash_rand(np.arange(n, dtype=np.uint32) + np.uint32(salt & 0xFFFFFFFF)) * np.float32(scale)).reshape(shape).astype(np.float32)


def conv_param_shapes(pipe_text: str) -> Dict[str, Tuple[int, ...]]:
    """Walk the pipe text (dims inferred with the reference's size rules, src/conv_util.cc:167-226) and return the shape of every parameter
    node: <tag>_filts (OC,C,KH,KW), <tag>_biases (OC,) [absent when bias_term=0], BatchNorm <tag>_mean / _var (C,) and _sf (1,), Scale
    <tag>_gamma / _beta (C,). An InnerProduct is the convolution whose window is its whole input."""
    import re
    dims: Dict[str, Tuple[int, int, int]] = {}  # node -> (chan, y, x)
    shapes: Dict[str, Tuple[int, ...]] = {}

    def yx(line, name, dflt):
        m = re.search(name + r"=\(tn=none,dims=\(y=(\d+),x=(\d+)\)\)", line)
        return (int(m.group(1)), int(m.group(2))) if m else dflt

    for line in pipe_text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        m = re.match(r"\(node=([^,]+),dims=\(img=(\d+),chan=(\d+),y=(\d+),x=(\d+)\)\)", line)
        if m:
            dims[m.group(1)] = (int(m.group(3)), int(m.group(4)), int(m.group(5)))
            continue
        tag = re.search(r"tag=([^,]+),", line).group(1)
        typ = re.search(r"type=([A-Za-z]+)", line).group(1)
        bots = re.search(r"bots=([^,)]+)", line).group(1).split(":")
        tops = re.search(r"tops=([^,)]+)", line).group(1).split(":")
        c, h, w = dims[bots[0]]
        if typ in ("Convolution", "InnerProduct"):
            oc = int(re.search(r"out_chans=\(tn=uint32_t,v=(\d+)\)", line).group(1))
            ky, kx = (h, w) if typ == "InnerProduct" else yx(line, "kern_sz", None)
            sy, sx = yx(line, "stride", (1, 1))
            py, px = yx(line, "in_pad", (0, 0))
            shapes[tag + "_filts"] = (oc, c, ky, kx)
            if "bias_term=(tn=uint32_t,v=0)" not in line:
                shapes[tag + "_biases"] = (oc,)
            dims[tops[0]] = (oc, (h + 2 * py - ky) // sy + 1, (w + 2 * px - kx) // sx + 1)
        elif typ == "Pooling":
            k = yx(line, "kern_sz", None)
            if k is None:
                dims[tops[0]] = (c, 1, 1)
            else:
                sy, sx = yx(line, "stride", (1, 1))
                py, px = yx(line, "in_pad", (0, 0))
                oh = 1 if h + 2 * py < k[0] else -(-(h + 2 * py - k[0]) // sy) + 1
                ow = 1 if w + 2 * px < k[1] else -(-(w + 2 * px - k[1]) // sx) + 1
                dims[tops[0]] = (c, oh, ow)
        elif typ == "Concat":
            dims[tops[0]] = (sum(dims[b][0] for b in bots), h, w)
        elif typ == "BatchNorm":
            shapes[tag + "_mean"] = (c,)
            shapes[tag + "_var"] = (c,)
            shapes[tag + "_sf"] = (1,)
        elif typ == "Scale":
            shapes[tag + "_gamma"] = (c,)
            shapes[tag + "_beta"] = (c,)
        else:
            dims[tops[0]] = dims[bots[0]]
    return shapes


def conv_algorithmic_elems(pipe_text: str, per_op: bool = False):
    """Sum over the Convolution ops of numel(in) + numel(out) + numel(filts) + numel(biases): the reference's algorithmic-bytes figure / 4
    (src/latex-util.H:119, pysrc/flops.py:101), from the C++ pipe IR's dims inference. per_op=True: {tag: (elems, out_pixels_per_image)}."""
    import re
    import boda_b200 as bb
    d = bb.pipe_describe(pipe_text)
    tot = 0
    ops = {}
    for line in pipe_text.splitlines():
        if "type=Convolution" not in line and "type=InnerProduct" not in line:
            continue
        tag = re.search(r"tag=([^,]+),", line).group(1)
        bot = re.search(r"bots=([^,)]+)", line).group(1).split(":")[0]
        top = re.search(r"tops=([^,)]+)", line).group(1).split(":")[0]
        e = 0
        for n in (bot, top, tag + "_filts", tag + "_biases"):
            if n in d["nodes"]:
                e += int(np.prod([sz for _, sz in d["nodes"][n]]))
        tot += e
        td = dict(d["nodes"][top])
        ops[tag] = (e, td["y"] * td["x"])
    return ops if per_op else tot


def synth_params(pipe_text: str, seed: int = 0) -> Dict[str, np.ndarray]:
    """filts ~ U(-a,a) with a = sqrt(6/K) (variance 2/K), biases ~ U(-0.5,0.5)/5: per-layer salts, no RNG state. BatchNorm blobs follow
    Caffe's storage convention (mean / var blobs hold sf x the statistic): sf = 2, mean ~ U(-.1,.1), var ~ U(.5,1.5); Scale gamma ~ U(.5,1.5),
    beta ~ U(-.1,.1); two ResNet-specific choices keep activations O(10) through 16 residual blocks (see the comments below)."""
    out = {}
    for i, (name, shape) in enumerate(sorted(conv_param_shapes(pipe_text).items())):
        salt = 8753985 + 7919 * i + 104729 * seed
        if name.endswith("_filts"):
            k = shape[1] * shape[2] * shape[3]
            out[name] = hash_fill(shape, salt, np.sqrt(6.0 / k) / 5.0)
        elif name.endswith("_sf"):
            out[name] = np.full(shape, 2.0, np.float32)
        elif name.endswith("_mean"):
            out[name] = (2.0 * hash_fill(shape, salt + 1, 0.1 / 5.0)).astype(np.float32)
        elif name.endswith("_var"):
            out[name] = (2.0 * (1.0 + hash_fill(shape, salt + 2, 0.5 / 5.0))).astype(np.float32)
            if name == "bn_conv1_var":  # the stem sees +-125 pixel-scale inputs: a variance that brings its output to O(1), as trained statistics do
                out[name] *= np.float32(9e4)
        elif name.endswith("_gamma"):
            out[name] = (1.0 + hash_fill(shape, salt + 3, 0.5 / 5.0)).astype(np.float32)
            if "branch2c" in name:  # damp the residual branches so 16 blocks of x + f(x) stay O(10) (cf. zero-gamma initialisation)
                out[name] *= np.float32(0.35)
        elif name.endswith("_beta"):
            out[name] = hash_fill(shape, salt + 4, 0.1 / 5.0)
        else:
            out[name] = hash_fill(shape, salt + 39475612, 0.1 / 5.0)
    return out


def synth_input(shape, seed: int = 0, scale: float = 25.0) -> np.ndarray:
    return hash_fill(shape, 234234567 + 15485863 * seed, scale)


NETS = {"alexnet_ng_conv": alexnet_ng_conv, "nin_imagenet": nin_imagenet, "googlenet_conv": googlenet_conv, "resnet50": resnet50, "tiny_net": tiny_net,
        "tiny_resnet": tiny_resnet}
