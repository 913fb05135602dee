"""CPU tests of the Caffe prototxt reader (boda_b200/csrc/caffe_prototxt.cu): host-only, no protobuf, no GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SMALL = r'''
name: "small"   # comment after a field
layer { name: "data" type: "Data" top: "data" top: "label" include { phase: TRAIN }
  data_param { batch_size: 256 source: "x" } transform_param { crop_size: 99 } }
layer { name: "data" type: "Data" top: "data" top: "label" include { phase: TEST }
  data_param { batch_size: 50 source: "y" backend: LMDB } transform_param { crop_size: 31 mirror: false } }
layer { name: "conv1" type: "Convolution" bottom: "data" top: "conv1"
  param { lr_mult: 1 } convolution_param { num_output: 8 kernel_size: 5 stride: 2 pad: 1 weight_filler { type: "gaussian" std: 0.01 } } }
layer { name: "relu1" type: "ReLU" bottom: "conv1" top: "conv1" }
layer { name: "norm1" type: "LRN" bottom: "conv1" top: "norm1" lrn_param { local_size: 3 alpha: 0.0001 beta: 0.75 } }
layer { name: "pool1" type: "Pooling" bottom: "norm1" top: "pool1" pooling_param { pool: MAX kernel_size: 3 stride: 2 } }
layer { name: "c2" type: "Convolution" bottom: "pool1" top: "c2" convolution_param { num_output: 6 kernel_h: 3 kernel_w: 1 pad_h: 1 pad_w: 0 bias_term: false } }
layer { name: "bn2" type: "BatchNorm" bottom: "c2" top: "c2" batch_norm_param { use_global_stats: true } }
layer { name: "sc2" type: "Scale" bottom: "c2" top: "c2" scale_param { bias_term: true } }
layer { name: "c3" type: "Convolution" bottom: "pool1" top: "c3" convolution_param { num_output: 6 kernel_size: 1 } }
layer { name: "sum" type: "Eltwise" bottom: "c2" bottom: "c3" top: "sum" }
layer { name: "cat" type: "Concat" bottom: "sum" bottom: "c3" top: "cat" }
layer { name: "drop" type: "Dropout" bottom: "cat" top: "cat" dropout_param { dropout_ratio: 0.4 } }
layer { name: "gp" type: "Pooling" bottom: "cat" top: "gp" pooling_param { pool: AVE global_pooling: true } }
layer { name: "fc" type: "InnerProduct" bottom: "gp" top: "fc" inner_product_param { num_output: 10 } }
layer { name: "prob" type: "Softmax" bottom: "fc" top: "prob" }
layer { name: "acc" type: "Accuracy" bottom: "fc" bottom: "label" top: "acc" include { phase: TEST } }
layer { name: "loss" type: "SoftmaxWithLoss" bottom: "fc" bottom: "label" top: "loss" }
'''


@pytest.fixture(scope="module")
def bb():
    import boda_b200 as m
    m.lib()
    return m


def test_small_net_translation(bb):
    pipe = bb.pipe_from_prototxt(SMALL, in_dims={"img": 4})
    d = bb.pipe_describe(pipe)
    n = {k: tuple(sz for _, sz in v) for k, v in d["nodes"].items()}
    assert n["data"] == (4, 3, 31, 31)                      # the TEST-phase Data layer, batch overridden, label top dropped
    assert n["conv1"] == (4, 8, 15, 15) and n["pool1"] == (4, 8, 7, 7) and n["c2"] == (4, 6, 7, 7)
    assert n["c2_filts"] == (6, 8, 3, 1) and "c2_biases" not in n and n["bn2_sf"] == (1,) and n["sc2_gamma"] == (6,)
    assert n["cat"] == (4, 12, 7, 7) and n["gp"] == (4, 12, 1, 1) and n["fc_filts"] == (10, 12, 1, 1)
    assert "prob" not in n and "acc" not in n and "loss" not in n   # Softmax / Accuracy / SoftmaxWithLoss are dropped (src/caffepb.cc:250-261,308)
    assert "type=Dropout" in pipe and "dropout_ratio=(tn=float,v=0.4)" in pipe
    keep = bb.pipe_describe(bb.pipe_from_prototxt(SMALL, in_dims={"img": 4}, keep_softmax=True))
    assert "prob" in keep["nodes"]
    cut = bb.pipe_describe(bb.pipe_from_prototxt(SMALL, out_node_name="pool1"))
    assert "pool1" in cut["nodes"] and "c2" not in cut["nodes"] and cut["nodes"]["data"][0] == ("img", 50)


def test_input_blobs_and_v1_layers(bb):
    v1 = '''input: "data" input_dim: 2 input_dim: 3 input_dim: 16 input_dim: 16
layers { name: "c" type: CONVOLUTION bottom: "data" top: "c" convolution_param { num_output: 4 kernel_size: 3 } }
layers { name: "r" type: RELU bottom: "c" top: "c" }
layers { name: "p" type: POOLING bottom: "c" top: "p" pooling_param { pool: AVE kernel_size: 2 stride: 2 } }'''
    d = bb.pipe_describe(bb.pipe_from_prototxt(v1))
    assert tuple(sz for _, sz in d["nodes"]["p"]) == (2, 4, 7, 7)
    shp = 'input: "x" input_shape { dim: 1 dim: 3 dim: 8 dim: 8 }\nlayer { name: "c" type: "Convolution" bottom: "x" top: "c" convolution_param { num_output: 2 kernel_size: 3 pad: 1 } }'
    assert tuple(sz for _, sz in bb.pipe_describe(bb.pipe_from_prototxt(shp, in_dims={"img": 5}))["nodes"]["c"]) == (5, 2, 8, 8)


def test_errors(bb):
    for bad in ('layer { name: "c" type: "Convolution" bottom: "d" top: "c" convolution_param { num_output: 2 } }',          # no kernel size
                'input: "d" input_dim: 1\nlayer { name: "r" type: "ReLU" bottom: "d" top: "d" }',                            # 1 input_dim, not 4
                'layer { name: "x" type: "Deconvolution" bottom: "d" top: "x" }',                                           # unsupported kind
                'layer { name: "c" type: "Convolution" bottom: "d" top: "c" convolution_param { num_output: 2 kernel_size: 3 group: 2 } }',
                'layer { name: "c" type: "ReLU" bottom: "d" top: "d" ',                                                      # missing }
                'layer { name: "d" type: "Dropout" bottom: "a" top: "b" }',                                                  # not in place -> clone
                # parameters that change the numerics and that this reader does not implement are rejected, never silently dropped:
                'layer { name: "c" type: "Convolution" bottom: "d" top: "c" convolution_param { num_output: 2 kernel_size: 3 dilation: 2 } }',
                'layer { name: "r" type: "ReLU" bottom: "d" top: "d" relu_param { negative_slope: 0.1 } }',
                'layer { name: "k" type: "Concat" bottom: "a" bottom: "b" top: "k" concat_param { axis: 2 } }',
                'layer { name: "s" type: "Scale" bottom: "d" top: "d" scale_param { bias_term: true axis: 0 } }',
                'layer { name: "p" type: "Pooling" bottom: "d" top: "p" pooling_param { pool: MAX kernel_size: 2 stride: 2 round_mode: FLOOR } }'):
        with pytest.raises(bb.RtException):
            bb.pipe_from_prototxt(bad)
    with pytest.raises(bb.RtException):
        bb.pipe_from_prototxt(SMALL, in_dims={"bogus": 3})
    with pytest.raises(bb.RtException):
        bb.pipe_from_prototxt(SMALL, out_node_name="nope")


@pytest.mark.parametrize("net", ["alexnet_ng_conv", "nin_imagenet", "googlenet_conv", "resnet50"])
def test_nets_py_matches_the_reference_prototxts(bb, net):
    """`boda_b200/nets.py` restates the BASELINE nets; the golden file holds what the reference's own prototxts translate to (generated here
    by tests/golden/make_pipes_from_prototxt.py): same nodes, dims, parameter nodes, op count and conv FLOPs."""
    from boda_b200 import nets
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "pipes_from_prototxt.json")))[net]
    txt, i, o = nets.NETS[net](g["in_dims"]["img"])
    d = bb.pipe_describe(txt)
    assert d["ops"] == g["ops"] and d["conv_flops"] == g["conv_flops"] and sorted(d["params"]) == g["params"]
    assert {n: [list(x) for x in dims] for n, dims in d["nodes"].items()} == g["nodes"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/nets"), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("ref_dir", ["alexnet_ng_conv", "googlenet_conv", "googlenet_conv_no_lrn", "nin_imagenet", "nin_imagenet_nopad", "resnet-50", "resnet-101",
                                     "resnet-152", "VGG16-v2-conv", "vgg_19", "squeezenet-1.0", "firenet-v0", "alexnet_ng_conv_nd_nl", "stratosnet-conv"])
def test_every_reference_net_parses_and_infers_dims(bb, ref_dir):
    txt = open("/root/reference/nets/%s/train_val.prototxt" % ref_dir).read()
    d = bb.pipe_describe(bb.pipe_from_prototxt(txt, in_dims={"img": 2}, keep_softmax=True))
    assert d["ops"] > 5 and d["conv_flops"] > 0
