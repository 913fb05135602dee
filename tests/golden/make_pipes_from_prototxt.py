#!/usr/bin/env python
"""Generate tests/golden/pipes_from_prototxt.json: the reference's own net descriptions (nets/*/train_val.prototxt, Caffe text format) run
through this repo's prototxt reader (b200_pipe_from_prototxt) and the C++ pipe IR (b200_pipe_describe) IN THIS CONTAINER, where
/root/reference exists. Stored per net: the in_dims used, every node with its dims, the parameter nodes, op count and conv FLOPs -- facts
about the reference's nets, so the CPU tests can check `boda_b200/nets.py`'s restatements against them on machines without the reference."""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
import boda_b200 as bb

REF = os.environ.get("BODA_REF", "/root/reference")
NETS = {  # golden key -> (reference dir, in_dims, keep_softmax)
    "alexnet_ng_conv": ("alexnet_ng_conv", {"img": 32}, False),
    "nin_imagenet": ("nin_imagenet", {"img": 4, "y": 227, "x": 227}, False),  # the reference's tests override the 224 crop to 227 (src/test_compute.cc:224)
    "googlenet_conv": ("googlenet_conv", {"img": 2}, False),
    "resnet50": ("resnet-50", {"img": 2}, True),
}
out = {}
for key, (d, in_dims, keep_softmax) in NETS.items():
    txt = open(os.path.join(REF, "nets", d, "train_val.prototxt")).read()
    pipe = bb.pipe_from_prototxt(txt, in_dims=in_dims, keep_softmax=keep_softmax)
    desc = bb.pipe_describe(pipe)
    out[key] = {"in_dims": in_dims, "keep_softmax": keep_softmax, "ops": desc["ops"], "conv_flops": desc["conv_flops"], "params": sorted(desc["params"]),
                "nodes": {n: [list(x) for x in dims] for n, dims in sorted(desc["nodes"].items())}}
with open(os.path.join(ROOT, "tests", "golden", "pipes_from_prototxt.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=True)
print({k: (v["ops"], len(v["nodes"])) for k, v in out.items()})
