"""Whole-net CPU oracle -- TEST INFRASTRUCTURE ONLY.

Interprets the same conv_pipe text that has_conv_fwd_t::init receives (one conv_op_t per line) with the per-op CPU
oracle, following the reference's forward semantics: in-place ReLU/Dropout on their node, conv+ReLU fused only when
ReLU is the first in-place op of the conv's output (src/rtc_fwd.cc:486-494 -- numerically identical to running it
separately), Dropout = identity at test time, Concat along chan, Eltwise = SUM, Softmax with max initialised to 0.
Returns every node, so tests can compare all nodes the way test_compute does (src/test_compute.cc:165-169).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import boda_oracle as bo


def _split(s):
    return [x for x in s.split(":") if x]


def run_pipe(pipe_text: str, inputs: Dict[str, np.ndarray], params: Dict[str, np.ndarray], round_to=None, acc64: bool = False) -> Dict[str, np.ndarray]:
    """acc64: convolutions accumulate in double (same order), see boda_oracle.conv_fwd. round_to: None | np.float16 | 'bf16' -- pre-round conv operands to the storage type (BASELINE configs C3/C4)."""
    nodes: Dict[str, np.ndarray] = {k: np.ascontiguousarray(v, np.float32) for k, v in inputs.items()}

    def rnd(a):
        if round_to is None:
            return a
        if round_to == "bf16":
            u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
            u = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16  # round to nearest even on the top 16 bits
            return u.astype(np.uint32).view(np.float32).reshape(a.shape)
        return a.astype(round_to).astype(np.float32)

    for line in pipe_text.splitlines():
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        d = bo.parse_lexp(line)
        if "node" in d:
            assert d["node"] in nodes, "missing input for source node " + d["node"]
            want = tuple(int(v) for v in d["dims"].values())
            assert nodes[d["node"]].shape == want, (nodes[d["node"]].shape, want)
            continue
        typ = d["str_vals"]["type"]
        tag, bots, tops = d["tag"], _split(d["bots"]), _split(d["tops"])
        nv = {k: bo._nda_from_lexp(v) for k, v in (d.get("nda_vals") or {}).items()}

        def yx(name, dflt):
            return (nv[name].dims["y"], nv[name].dims["x"]) if name in nv else dflt

        if typ in ("Convolution", "InnerProduct"):  # InnerProduct = the convolution whose window is its whole input (src/caffepb.cc:276-279)
            w = params[tag + "_filts"]
            b = params[tag + "_biases"] if (tag + "_biases") in params else np.zeros(w.shape[0], np.float32)  # bias_term: false
            nodes[tops[0]] = bo.conv_fwd(rnd(nodes[bots[0]]), rnd(w), b, yx("stride", (1, 1)), yx("in_pad", (0, 0)), relu=False, acc64=acc64)
        elif typ == "BatchNorm":  # Caffe BatchNormLayer::Forward_cpu with use_global_stats: blobs are divided by the scale-factor blob (0 -> 0)
            assert bots == tops
            sf = float(params[tag + "_sf"].ravel()[0])
            s = np.float32(0.0 if sf == 0 else 1.0 / sf)
            mean, var = params[tag + "_mean"] * s, params[tag + "_var"] * s
            eps = np.float32(float(nv["eps"].v) if "eps" in nv else 1e-5)
            x = nodes[bots[0]]
            nodes[tops[0]] = ((x - mean[None, :, None, None]) / np.sqrt(var + eps)[None, :, None, None]).astype(np.float32)
        elif typ == "Scale":  # Caffe ScaleLayer with bias_term: y = gamma * x + beta per channel
            assert bots == tops
            x = nodes[bots[0]]
            nodes[tops[0]] = (x * params[tag + "_gamma"][None, :, None, None] + params[tag + "_beta"][None, :, None, None]).astype(np.float32)
        elif typ == "ReLU":
            assert bots == tops
            nodes[tops[0]] = bo.relu(nodes[bots[0]])
        elif typ == "Dropout":
            assert bots == tops
        elif typ == "LRN":
            nodes[tops[0]] = bo.lrn_fwd(nodes[bots[0]], int(nv["local_size"].v), float(nv["alpha"].v), float(nv["beta"].v), float(nv["k"].v))
        elif typ == "Pooling":
            nodes[tops[0]] = bo.pool_fwd(nodes[bots[0]], yx("kern_sz", None), yx("stride", (1, 1)), yx("in_pad", (0, 0)),
                                         avg_pool=bool(int(nv["avg_pool"].v)) if "avg_pool" in nv else False)
        elif typ == "Concat":
            nodes[tops[0]] = bo.concat([nodes[b] for b in bots])
        elif typ in ("Eltwise", "Reduce"):
            nodes[tops[0]] = bo.reduce_sum([nodes[b] for b in bots])
        elif typ == "Softmax":
            nodes[tops[0]] = bo.softmax(nodes[bots[0]])
        else:
            raise ValueError("net_oracle: unhandled op type " + typ)
    return nodes
