"""Host-only tests of the forward PLAN (b200_fwd_plan: graph passes + launch plans, no device): concat by offset, residual joins inside the
producing convolution, producer-written planes, abs-max cells. The same decisions are exercised numerically by tests/test_gpu_nets.py; here
their structure is checked on the CPU, on the four BASELINE nets at their bench batch sizes."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def bb():
    import boda_b200
    boda_b200.lib()
    return boda_b200


def _kinds(plan):
    k = {}
    for f, _ in plan["calls"]:
        k[f.split("__")[0]] = k.get(f.split("__")[0], 0) + 1
    return k


def test_resnet50_joins_are_planned_into_their_convolutions(bb):
    from boda_b200 import nets
    txt, i, o = nets.resnet50(32)
    plan = bb.fwd_plan(txt, "")
    assert _kinds(plan) == {"conv": 54, "pool": 2, "softmax": 1}  # 53 convolutions + fc1000; no reduce kernels left
    assert len(plan["join"]) == 16 and len(plan["prep"]) == 53     # one BatchNorm/Scale fold per convolution
    order = {f.split("__")[1]: n for n, (f, _) in enumerate(plan["calls"])}
    writer_of = {a["out"]: f.split("__")[1] for f, a in plan["calls"] if "out" in a}
    for conv_tag, (join_node, res_node) in plan["join"].items():
        assert conv_tag.endswith("_branch2c") and join_node == conv_tag[:-len("_branch2c")]
        args = dict(plan["calls"][order[conv_tag]][1])
        assert args["out"] == join_node and args["res"] == res_node
        assert res_node in plan["absmax"] and join_node in plan["absmax"]    # the residual's max bounds the fp16 planes of the join
        feeds_conv = any(f.startswith("conv__") and a.get("in") == join_node for f, a in plan["calls"])
        assert "res_absmax_cells" in args and ("out_pack" in args) == feeds_conv  # the epilogue also writes the consumers' planes (res5c feeds a pool)
        assert order[writer_of[res_node]] < order[conv_tag]                  # the shortcut is complete before the convolution runs
        # projection blocks join with their branch1 convolution, identity blocks with the previous join
        assert res_node == (join_node + "_branch1" if join_node.endswith("a") else writer_of[res_node]) or res_node.startswith("res")
    unfused = bb.fwd_plan(txt, "(fuse_eltwise=0)")
    assert _kinds(unfused) == {"conv": 54, "pool": 2, "softmax": 1, "reduce": 16} and not unfused["join"]
    assert len(unfused["calls"]) == len(plan["calls"]) + 16


def test_googlenet_concat_inputs_are_written_in_place(bb):
    from boda_b200 import nets
    txt, i, o = nets.googlenet_conv(64)
    desc = bb.pipe_describe(txt)["nodes"]
    plan = bb.fwd_plan(txt, "(prec=bf16)")
    assert _kinds(plan) == {"conv": 64, "pool": 16, "lrn": 2} and len(plan["alias"]) == 36   # 9 inception Concats x 4 inputs, no copy kernels
    chans = lambda n: dict(desc[n])["chan"]
    per_cat = {}
    for node, (cat, ocix) in plan["alias"].items():
        per_cat.setdefault(cat, []).append((ocix, chans(node)))
        call = next(a for f, a in plan["calls"] if a.get("out") == node)
        assert call["out_concat"] == cat and int(call["out_ocix"]) == ocix and ocix % 8 == 0
    assert len(per_cat) == 9
    for cat, parts in per_cat.items():  # the four slices tile the Concat output exactly
        parts.sort()
        assert parts[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(parts, parts[1:])) and parts[-1][0] + parts[-1][1] == chans(cat)
    copies = bb.fwd_plan(txt, "(prec=bf16,concat_by_offset=0)")
    assert _kinds(copies).get("copy") == 36 and not copies["alias"]
    # bf16 planes pass through Concats; the fp16 planes of the fp32-parity mode do not (each producer would derive its own scale)
    n_pack = lambda p: sum(1 for _, a in p["calls"] if "out_pack" in a)
    assert n_pack(plan) > n_pack(bb.fwd_plan(txt, "")) > 0 and n_pack(bb.fwd_plan(txt, "(pack_by_producers=0)")) == 0


@pytest.mark.parametrize("net,batch,n_calls", [("alexnet_ng_conv", 32, 9), ("nin_imagenet", 32, 16)])
def test_plain_chains_plan_one_call_per_layer(bb, net, batch, n_calls):
    from boda_b200 import nets
    txt, i, o = nets.NETS[net](batch)
    plan = bb.fwd_plan(txt, "")
    assert len(plan["calls"]) == n_calls and not plan["alias"] and not plan["join"] and not plan["prep"]
    convs = [a for f, a in plan["calls"] if f.startswith("conv__")]
    assert all("in_absmax_cells" in a for a in convs[1:])  # every convolution but the first gets max|in| from its producer


def test_a_plan_cannot_compute(bb):
    """plan_only instances exist for inspection; anything that would need the device refuses (there is no CPU fallback)."""
    from boda_b200 import nets
    txt, i, o = nets.tiny_resnet(2)
    plan = bb.fwd_plan(txt, "")
    assert len(plan["join"]) == 3 and "reduce" not in _kinds(plan)
    f = bb.B200ConvFwd(txt, "(plan_only=1)")
    with pytest.raises(bb.RtException):
        f.set_param("conv1_filts", np.zeros((16, 3, 7, 7), np.float32))
    with pytest.raises(bb.RtException):
        f.run_fwd({i: np.zeros((2, 3, 33, 33), np.float32)}, [o])
    with pytest.raises(bb.RtException):
        bb.fwd_plan(txt, "(bogus=1)")
    with pytest.raises(bb.RtException):
        bb.fwd_plan(txt, "prec=bf16")  # not a (key=value,...) list


RES_BLOCK = '''
input: "data" input_dim: 2 input_dim: 16 input_dim: 12 input_dim: 12
%s
layer { name: "sum" type: "Eltwise" bottom: "short" bottom: "main" top: "sum" }
layer { name: "sum_relu" type: "ReLU" bottom: "sum" top: "sum" }
layer { name: "next" type: "Convolution" bottom: "sum" top: "next" convolution_param { num_output: 24 kernel_size: 3 pad: 1 } }
'''
SHORT = 'layer { name: "short" type: "Convolution" bottom: "data" top: "short" convolution_param { num_output: 32 kernel_size: 1 } }\n'
MAIN = ('layer { name: "main" type: "Convolution" bottom: "data" top: "main" convolution_param { num_output: 32 kernel_size: 3 pad: 1 bias_term: false } }\n'
        'layer { name: "bn" type: "BatchNorm" bottom: "main" top: "main" batch_norm_param { use_global_stats: true } }\n'
        'layer { name: "sc" type: "Scale" bottom: "main" top: "main" scale_param { bias_term: true } }\n')


def test_join_rules_on_a_prototxt_block(bb):
    """From a Caffe prototxt (f1) to the plan (f4): the join goes into the convolution that runs LAST; a branch with other readers, or with
    a ReLU of its own in front of the join, keeps the separate reduce kernel."""
    def plan_of(layers, opts=""):
        return bb.fwd_plan(bb.pipe_from_prototxt(RES_BLOCK % layers), opts)
    p = plan_of(SHORT + MAIN)
    assert p["join"] == {"main": ("sum", "short")} and "reduce" not in _kinds(p)
    call = dict(next(a for f, a in p["calls"] if f.startswith("conv__main__")))
    assert call["out"] == "sum" and call["res"] == "short" and call["filts"] == "main_filts__folded" and "out_pack" in call
    assert "short" in p["absmax"] and [f for f, _ in p["prep"]] == ["bn_fold__main"]
    # the shortcut produced after the main branch: the join moves into the shortcut's convolution instead
    assert plan_of(MAIN + SHORT)["join"] == {"short": ("sum", "main")}
    # a second reader of the would-be bypassed node, or a ReLU on it, forbids the fusion on that side ...
    two_readers = plan_of(SHORT + MAIN.replace('name: "sc"', 'name: "sc"') + 'layer { name: "p" type: "Pooling" bottom: "main" top: "p" pooling_param { pool: MAX kernel_size: 2 stride: 2 } }\n')
    assert two_readers["join"] == {} and _kinds(two_readers)["reduce"] == 1
    relu_first = plan_of(SHORT + MAIN + 'layer { name: "mr" type: "ReLU" bottom: "main" top: "main" }\n')
    assert relu_first["join"] == {} and _kinds(relu_first)["reduce"] == 1
    # ... and the option turns it off altogether
    assert plan_of(SHORT + MAIN, "(fuse_eltwise=0)")["join"] == {}


def test_alexnet_launch_plans_match_the_measured_launch_list(bb):
    """The per-layer launch plans. Round-2 kernel (igemm4.cuh, use_sk4=1): conv1 in its im2col mode (row-merged operand), conv2-5 in halo mode,
    fc6-8 with the weights as the 128-row operand; stream-K where whole tiles would leave more than 8 % of a round idle (conv3/4: 75 tiles on 74
    pairs, conv5: 50, the inner-product layers: 16 / 16 / 4 tiles). By default the round-1 plans, which equal what ncu saw on the B200
    (profiles/launches_r01_final_ncu_graph_nodes.md): igemm_umma_2cta_kernel<96|128, 2> with grids 148 / 148 / 132 / 132 / 132 for conv1..conv5,
    igemm_umma_kernel<32, 2> with grids (32,1,4), (32,1,4), (8,1,8) for fc6..fc8 (fuse_fc_chain=0; by default those three are one fc_chain call)."""
    from boda_b200 import nets
    txt, i, o = nets.alexnet_ng_conv(32)

    def plans(opts):
        plan = bb.fwd_plan(txt, opts)
        return plan, [(f.split("__")[1], {k[5:]: v for k, v in a.items() if k.startswith("plan:")}) for f, a in plan["calls"] if f.startswith("conv__")]
    plan, got = plans("(use_sk4=1)")
    want = [("conv1", "im2col", 96, 379, 1, 11), ("conv2", "halo", 128, 210, 0, 50), ("conv3", "halo", 128, 75, 1, 36), ("conv4", "halo", 128, 75, 1, 54),
            ("conv5", "halo", 128, 50, 1, 54), ("fc6-conv", "2d", 32, 16, 1, 144), ("fc7-conv", "2d", 32, 16, 1, 64), ("fc8-conv", "2d", 32, 4, 1, 64)]
    assert all(p["kernel"] == "sk4" and p["grid"] == "148x1x1" for _, p in got), got
    assert [(t, p["mode"], int(p["bn"]), int(p["tiles"]), int(p["streamk"]), int(p["kblks"])) for t, p in got] == want
    assert [int(p["swapped"]) for _, p in got] == [0, 0, 0, 0, 0, 1, 1, 1] and [int(p["rowmerge"]) for _, p in got] == [1, 0, 0, 0, 0, 0, 0, 0]
    # producers write their consumers' planes in the consumer's layout: pool1 -> conv2 (5x5, pad 2), pool2 -> conv3, conv3 -> conv4, conv4 -> conv5 (3x3, pad 1)
    pads = {f.split("__")[1]: (a.get("out_pack_py"), a.get("out_pack_px")) for f, a in plan["calls"] if "out_pack" in a}
    assert pads == {"pool1": ("2", "2"), "pool2": ("1", "1"), "conv3": ("1", "1"), "conv4": ("1", "1"), "pool5": (None, None)}, pads
    _, got1 = plans("(fuse_fc_chain=0)")
    want1 = [("conv1", "pair", 96, "148x1x1", 11), ("conv2", "pair", 128, "148x1x1", 50), ("conv3", "pair", 128, "132x1x1", 36), ("conv4", "pair", 128, "132x1x1", 54),
             ("conv5", "pair", 96, "132x1x1", 54), ("fc6-conv", "single", 32, "32x1x4", 144), ("fc7-conv", "single", 32, "32x1x4", 64), ("fc8-conv", "single", 32, "8x1x8", 64)]
    assert [(t, p["kernel"], int(p["bn"]), p["grid"], int(p["kblks"])) for t, p in got1] == want1
    # a smaller chip (plan_num_sms) re-plans: fewer SM pairs -> fewer persistent clusters
    small = bb.fwd_plan(txt, "(plan_num_sms=64)")
    g = dict(next(a for f, a in small["calls"] if f.startswith("conv__conv2__")))
    assert g["plan:grid"] == "64x1x1"
    # pooling / LRN calls carry no launch plan
    assert not any(k.startswith("plan:") for f, a in plan["calls"] if not f.startswith("conv__") for k in a)


def test_reference_op_tune_knob_names_are_accepted_and_ignored(bb):
    """Existing `--op-tune=(...)` command lines keep working (SURVEY section 5 "Config/flags"): the reference's op_tune_t knob names
    (src/cnn_op.H:10-32) select among CUCL variants that do not exist here; they are accepted, flat or nested, and change nothing.
    use_be naming another back-end and truly unknown keys are still errors (NESI rejects unused keys, src/nesi.cc:25-35)."""
    from boda_b200 import nets
    txt, i, o = nets.tiny_net(4)
    base = bb.fwd_plan(txt, "")["calls"]
    for opts in ("(k1conv=1,tconv=1)", "(op_tune=(use_be=b200,use_culibs=0,MNt=8:8,MNb=8:16,Kb=8,use_local_mem=1,prof_variant=0,vw=8,k1conv=1,tconv=2,tconv_max_ksz=11:11,ipconv=1))",
                 "(ipconv=1,prec=fp32)"):
        assert bb.fwd_plan(txt, opts)["calls"] == base, opts
    for bad in ("(op_tune=(use_be=ocl))", "(op_tune=(bogus_knob=1))", "(bogus_option=1)"):
        with pytest.raises(bb.RtException):
            bb.fwd_plan(txt, bad)


def test_node_listed_twice_in_a_concat_is_copied_both_times(bb):
    """A node that appears twice among one Concat's bottoms has two destinations: it is not written in place (concat by offset aliases a node
    to ONE channel offset) but copied at both offsets, as the reference does (src/rtc_fwd.cc:267-280)."""
    from boda_b200.nets import PipeBuilder
    p = PipeBuilder()
    p.data("data", 2, 8, 12, 12)
    p.conv("a", "data", "a", 16, 3, 1, 1, relu="ra")
    p.conv("b", "data", "b", 8, 1, relu="rb")
    p.concat("cat", ["a", "b", "a"], "cat")
    p.conv("c", "cat", "c", 8, 1)
    plan = bb.fwd_plan(p.text(), "")
    assert "a" not in plan["alias"] and plan["alias"]["b"][1] == 16   # b alone is written in place, at channel 16
    copies = [a for f, a in plan["calls"] if f.startswith("copy__")]
    assert len(copies) == 2 and all(c["in"] == "a" for c in copies)


def test_lrn_in_front_of_a_max_pool_is_planned_into_the_pool(bb):
    """AlexNet's norm1 -> pool1 and norm2 -> pool2 (and GoogLeNet's norm2 -> pool2) run as ONE kernel: the pool call reads the LRN's input and
    carries the LRN parameters by value, the lrn call is gone. GoogLeNet's norm1 sits BEHIND pool1 and feeds a convolution: not fused.
    fuse_lrn_pool=0 restores one call per layer."""
    from boda_b200 import nets
    txt, i, o = nets.alexnet_ng_conv(32)
    plan = bb.fwd_plan(txt, "")
    assert plan["lrnpool"] == {"pool1": ("norm1", "conv1"), "pool2": ("norm2", "conv2")}
    assert not [f for f, _ in plan["calls"] if f.startswith("lrn__")]
    pools = {a["out"]: a for f, a in plan["calls"] if f.startswith("pool__")}
    assert pools["pool1"]["in"] == "conv1" and pools["pool1"]["lrn_local_size"] == "5" and "lrn_local_size" not in pools["pool5"]
    assert "in_absmax_cells" in pools["pool1"]  # the planes pool1 writes for conv2 are scaled by max|conv1|, published by conv1
    off = bb.fwd_plan(txt, "(fuse_lrn_pool=0)")
    assert not off["lrnpool"] and len(off["calls"]) == len(plan["calls"]) + 2 and len([f for f, _ in off["calls"] if f.startswith("lrn__")]) == 2
    txt, i, o = nets.googlenet_conv(8)
    plan = bb.fwd_plan(txt, "")
    assert plan["lrnpool"] == {"pool2": ("norm2", "conv2")}
    assert [a["out"] for f, a in plan["calls"] if f.startswith("lrn__")] == ["norm1"]
    assert bb.fwd_plan(nets.googlenet_conv(64)[0], "(prec=bf16)")["lrnpool"] == {}


def test_inner_product_chains_are_planned_as_one_call(bb):
    """AlexNet fc6 -> fc7 -> fc8 at batch <= 32 is ONE fc_chain call (fcchain.cuh) carrying every layer's filts / biases / out and abs-max
    cells; GoogLeNet's auxiliary classifier pairs too; at batch 64 (two image tiles) and for ResNet-50's single fc1000 nothing is chained.
    fuse_fc_chain=0 and the round-2 contraction kernel (use_sk4=1, which runs these layers itself) restore one call per layer."""
    from boda_b200 import nets
    txt, i, o = nets.alexnet_ng_conv(32)
    plan = bb.fwd_plan(txt, "")
    assert plan["fcchain"] == [["fc6-conv", "fc7-conv", "fc8-conv"]]
    (f, a), = [(f, a) for f, a in plan["calls"] if f.startswith("fc_chain__")]
    assert plan["calls"][-1][0] == f and a["in"] == "pool5" and [a["out%d" % k] for k in range(3)] == ["fc6", "fc7", "fc8"]
    assert [a["filts%d" % k] for k in range(3)] == ["fc6-conv_filts", "fc7-conv_filts", "fc8-conv_filts"] and "biases2" in a
    assert "in_absmax_ix" in a and "out0_absmax_ix" in a and "out1_absmax_ix" in a and "out2_absmax_ix" not in a  # nobody reads max|fc8|
    for opts in ("(fuse_fc_chain=0)", "(use_sk4=1)"):
        off = bb.fwd_plan(txt, opts)
        assert not off["fcchain"] and len(off["calls"]) == len(plan["calls"]) + 2
    assert bb.fwd_plan(nets.googlenet_conv(8)[0], "")["fcchain"] == [["cls1_fc1-conv", "cls1_fc2-conv"], ["cls2_fc1-conv", "cls2_fc2-conv"]]
    assert bb.fwd_plan(nets.googlenet_conv(64)[0], "(prec=bf16)")["fcchain"] == []
    assert bb.fwd_plan(nets.resnet50(32)[0], "")["fcchain"] == []
    assert bb.fwd_plan(nets.alexnet_ng_conv(64)[0], "")["fcchain"] == []
