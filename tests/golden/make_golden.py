#!/usr/bin/env python
"""Decode the reference's golden per-op digests into a small fixture that travels with the repo.

Run in the dev container (where /root/reference exists):  python tests/golden/make_golden.py
Reads  <ref>/test/good_tr/<test>/wisdom.wis  for the six ops-prof tests that pin the rtc_fwd conv/sgemm path
 and writes tests/golden/wisdom_digests.json: for every op its op-line text, and for every
known-good output var the raw digest hex plus the decoded header (dims, seed, min/max, #samples) for readability.
The wisdom text format is src/op-tuner.cc:70-130; the digest binary layout is src/boda_base.cc:329-363.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import boda_oracle as bo  # noqa: E402

REF = os.environ.get("BODA_REF", "/root/reference")
TESTS = ["sgemm-gen600", "sgemm-gen5", "conv-gen5", "conv-debug", "conv-full-gen5", "ops-prof-conv-3x3-cudnn-boda"]


def read_wisdom(fn):
    ops = []
    with open(fn) as f:
        lines = [l.rstrip("\n") for l in f]
    i = 0
    while i < len(lines):
        if lines[i] == "op_wisdom_t":
            op_line = lines[i + 1]
            i += 2
            kgs = []
            while lines[i] != "/op_wisdom_t":
                if lines[i] == "kg":
                    kgs.append({"var": lines[i + 1], "digest_hex": lines[i + 2]})
                    i += 3
                else:
                    i += 1  # op_tune_wisdom_t blocks etc: timings only, not golden data
            ops.append({"op": op_line, "kgs": kgs})
        i += 1
    return ops


def main():
    out = {"source": "moskewcz/boda test/good_tr/*/wisdom.wis", "tests": {}}
    for t in TESTS:
        ops = read_wisdom(os.path.join(REF, "test", "good_tr", t, "wisdom.wis"))
        for o in ops:
            for kg in o["kgs"]:
                d = bo.decode_digest(kg["digest_hex"])
                kg["decoded"] = {"dims": dict(zip(d.dim_names, d.sizes)), "seed": d.seed, "min_v": d.min_v,
                                 "max_v": d.max_v, "n_samps": len(d.samps)}
        out["tests"][t] = ops
        print(t, len(ops), "ops")
    with open(os.path.join(HERE, "wisdom_digests.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))


if __name__ == "__main__":
    main()
