"""CPU suite for the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/boda_b200.h
declares, refuses to compute without a device (no CPU fallback), and its host-side logic (op-line grammar, conv_pipe
dims inference, error codes) behaves like the reference's."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bb():
    import boda_b200
    if not os.path.exists(boda_b200.LIB_PATH):
        from boda_b200 import build
        build.build()
    return boda_b200


def test_header_symbols_all_exported(bb):
    hdr = open(os.path.join(ROOT, "include", "boda_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    out = subprocess.check_output(["nm", "-D", "--defined-only", bb.LIB_PATH], text=True)
    exported = set(re.findall(r" T (b200_[a-z0-9_]+)", out))
    assert declared <= exported, sorted(declared - exported)
    assert set(bb.ABI_SYMBOLS) == declared, sorted(set(bb.ABI_SYMBOLS) ^ declared)
    bb.lib()  # ctypes load + every signature bound


def test_no_gpu_fails_loudly(bb):
    if bb.device_count() > 0:
        pytest.skip("a GPU is present")
    r = bb.B200Compute()
    with pytest.raises(bb.RtException, match="no CPU fallback"):
        r.init()
    with pytest.raises(bb.RtException):
        from boda_b200 import nets
        bb.B200ConvFwd(nets.tiny_net(1)[0], "")


def test_cuda_sass_is_blackwell_native(bb):
    """The shipped library carries sm_100a SASS with tcgen05 (UTCHMMA), TMA (UTMALDG incl. im2col) and TMEM loads (LDTM)."""
    sass = subprocess.run(["cuobjdump", "-sass", bb.LIB_PATH], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG.2D", "UTMALDG.4D.IM2COL", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_pipe_text_matches_oracle_dims(bb, oracle):
    """nets.py's parameter-shape walk and the oracle's executor agree on the AlexNet-ng layer shapes."""
    from boda_b200 import nets
    txt, i, o = nets.alexnet_ng_conv(32)
    shapes = nets.conv_param_shapes(txt)
    assert shapes["conv1_filts"] == (96, 3, 11, 11) and shapes["conv2_filts"] == (256, 96, 5, 5)
    assert shapes["fc6-conv_filts"] == (4096, 256, 6, 6) and shapes["fc8-conv_biases"] == (1000,)
    assert sum(int(np.prod(s)) for s in shapes.values()) == 62378344  # 62.4 M params
    p = nets.synth_params(txt)
    assert p["conv3_filts"].dtype == np.float32 and abs(float(p["conv3_filts"].mean())) < 1e-3


@pytest.mark.parametrize("net,kw", [("alexnet_ng_conv", dict(batch=32)), ("nin_imagenet", dict(batch=4)), ("googlenet_conv", dict(batch=2)),
                                    ("resnet50", dict(batch=2)), ("tiny_resnet", dict(batch=2)), ("tiny_net", dict(batch=3))])
def test_cpp_pipe_dims_match_python_walk(bb, net, kw):
    """The C++ conv_pipe graph IR (text parser + calc_dims, host-only) against nets.py's independent shape walk: same parameter nodes with the
    same shapes for every net family, incl. ResNet-50's BatchNorm / Scale / InnerProduct / bias-less convolutions."""
    from boda_b200 import nets
    txt, i, o = nets.NETS[net](**kw)
    d = bb.pipe_describe(txt)
    shapes = nets.conv_param_shapes(txt)
    assert sorted(d["params"]) == sorted(shapes)
    for n, shp in shapes.items():
        assert tuple(sz for _, sz in d["nodes"][n]) == tuple(shp), n
    assert o in d["nodes"] and i in d["nodes"]
    if net == "alexnet_ng_conv":
        assert d["conv_flops"] == 72656390144  # 72.66 GFLOP per 32-image forward
    if net == "resnet50":
        assert tuple(sz for _, sz in d["nodes"]["pool5"]) == (2, 2048, 1, 1) and tuple(sz for _, sz in d["nodes"]["res3a"]) == (2, 512, 28, 28)
        assert abs(d["conv_flops"] / 2 / 2 - 3.86e9) < 0.1e9  # ~3.86 GMAC per image


def test_cpp_pipe_errors(bb):
    from boda_b200 import nets
    with pytest.raises(bb.RtException):  # BatchNorm that is not in place on a Convolution output
        bb.pipe_describe("(node=data,dims=(img=1,chan=3,y=8,x=8))\n(tag=bn,str_vals=(type=BatchNorm),bots=data,tops=other)\n")
    with pytest.raises(bb.RtException):
        bb.pipe_describe("(node=data,dims=(img=1,chan=3,y=8,x=8))\n(tag=c,str_vals=(type=Convolution),nda_vals=(out_chans=(tn=uint32_t,v=4)),bots=data,tops=c)\n")  # no kern_sz


def test_pipe_op_signatures_equal_the_reference_op_list(bb):
    """write_op_sigs restated (src/rtc_fwd.cc:246-264): the unique Convolution signatures of AlexNet-ng, NiN (native 224 crop) and GoogLeNet at
    batch 1 / 5 / 20, printed by the C++ pipe IR's canonical op printer, are TEXT-IDENTICAL to the 204 ops of the reference's
    conv-ops-1-5-20-nin-alex-gn list (taken from its committed test/good_tr/conv-full-gen5/wisdom.wis) -- which pins the three net
    descriptions, the dims inference and the op-line printer against the reference's own output."""
    import json
    from boda_b200 import nets
    gold = {}
    for e in json.load(open(os.path.join(ROOT, "tests", "golden", "wisdom_digests.json")))["tests"]["conv-full-gen5"]:
        gold.setdefault(int(e["op"].split("in=(dims=(img=")[1].split(",")[0]), set()).add(e["op"])
    assert sorted(gold) == [1, 5, 20] and all(len(v) == 68 for v in gold.values())
    for B in (1, 5, 20):
        mine = set()
        for txt in (nets.alexnet_ng_conv(B)[0], nets.nin_imagenet(B, in_sz=224)[0], nets.googlenet_conv(B)[0]):
            sigs = bb.pipe_op_sigs(txt)
            assert sigs == sorted(set(sigs))  # unique, sorted
            mine |= set(sigs)
        assert mine == gold[B], (B, sorted(mine ^ gold[B])[:2])
    # the per-op sweep file of BASELINE config C2 is exactly AlexNet-ng's signature set at batch 32
    c2 = [l.strip() for l in open(os.path.join(ROOT, "ops", "c2-alexnet-ng-b32-convs.txt")) if l.strip()]
    assert set(c2) == set(bb.pipe_op_sigs(nets.alexnet_ng_conv(32)[0])) and len(c2) == 8
    # every signature is a valid op line for the per-op tier (round trip through the op parser of b200_wis_ana's reader)
    one = sorted(gold[20])[0]
    rec = bb.wisdom_record(one, [], "(use_be=b200)", "b200:x", 1e-3)
    assert bb.wis_ana(rec, s_plat="b200:")["rows"][0]["op"] == one


def test_stale_op_syntax_canonicalises_to_the_current_one(bb):
    """The reference's older op lists (test/conv-ops-small.txt, test/sgemm-ops-tiny.txt) use the stale `type/dims_vals` syntax (SURVEY Appendix
    A). The C++ op parser + canonical printer turn them into the current-syntax lines its newer lists and wisdom files hold."""
    import json
    stale = ["(type=Convolution,dims_vals=(biases=(out_chan=96),filts=(out_chan=96,in_chan=3,y=11,x=11),in=(img=20,chan=3,y=227,x=227),in_pad=(y=0,x=0),"
             "kern_sz=(y=11,x=11),out=(img=20,chan=96,y=55,x=55),stride=(y=4,x=4)),str_vals=(out_chans=96))",
             "(type=Convolution,dims_vals=(biases=(out_chan=1024),filts=(out_chan=1024,in_chan=384,y=3,x=3),in=(img=20,chan=384,y=6,x=6),in_pad=(y=1,x=1),"
             "kern_sz=(y=3,x=3),out=(img=20,chan=1024,y=6,x=6),stride=(y=1,x=1)),str_vals=(out_chans=1024))",
             "(type=sgemm,dims_vals=(a=(K=2048,M=2048),b=(K=2048,N=2048),c=(M=2048,N=2048)))"]
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "wisdom_digests.json")))["tests"]
    known = set(e["op"] for e in g["conv-full-gen5"]) | set(e["op"] for e in g["sgemm-gen5"])
    for l in stale:
        c = bb.op_canonical(l)
        assert c in known and bb.op_canonical(c) == c  # the reference's own text for this op; canonical lines are fixed points
    for fn in ("c1-sgemm-ops-tiny.txt", "c2-alexnet-ng-b32-convs.txt", "c3-conv-ops-small.txt"):
        for l in open(os.path.join(ROOT, "ops", fn)):
            if l.strip():
                assert bb.op_canonical(l.strip()) == l.strip()
    # scalar / float / typed-dims ndas print in the reference's minimal form
    assert bb.op_canonical("(str_vals=(type=LRN),nda_vals=(alpha=(tn=float,v=0.0001),local_size=(tn=uint32_t,v=5),in=(tn=float,dims=(img=2,chan=3))))") == \
        "(str_vals=(type=LRN),nda_vals=(alpha=(tn=float,v=0.0001),in=(dims=(img=2,chan=3)),local_size=(tn=uint32_t,v=5)))"
    with pytest.raises(bb.RtException):
        bb.op_canonical("(type=Convolution,bogus_vals=())")
