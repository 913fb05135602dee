#!/usr/bin/env python
"""Generates tests/golden/wisdom_merged_c3_subset.wis from the reference's published timing database (test/wisdom-merged.wis, read from
/root/reference in the build container only): the records of the ops of ops/c3-conv-ops-small.txt that the database holds (B = 20),
reduced to the runs on `nvrtc:GeForce GTX TITAN X` (7 Boda tunes + the cuDNN tune) and with each run's annotated op line replaced by the
plain op line (the analysis never reads it). Measurement data only -- no reference source. Re-run: python tests/golden/make_wisdom_subset.py"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/test/wisdom-merged.wis"
PLAT = "nvrtc:GeForce GTX TITAN X"


def main():
    want = [l.strip() for l in open(os.path.join(ROOT, "ops", "c3-conv-ops-small.txt")) if l.strip()]
    lines = open(REF).read().split("\n")
    out, i, kept = [], 0, 0
    while i < len(lines):
        if lines[i] != "op_wisdom_t":
            i += 1
            continue
        j = i
        while lines[j] != "/op_wisdom_t":
            j += 1
        rec, op = lines[i:j + 1], lines[i + 1]
        i = j + 1
        if op not in want:
            continue
        kept += 1
        out += ["op_wisdom_t", op]
        k = 2
        while k < len(rec) - 1:
            assert rec[k] == "op_tune_wisdom_t", rec[k]
            tune, k = rec[k + 1], k + 2
            runs = []
            while rec[k] != "/op_tune_wisdom_t":
                assert rec[k] == "op_run_t"
                plat, secs, err = rec[k + 1], rec[k + 2], rec[k + 3]
                k += 4 if err else 5
                if plat == PLAT:
                    runs += ["op_run_t", plat, secs, err] + ([] if err else [op])
            k += 1
            if runs:
                out += ["op_tune_wisdom_t", tune] + runs + ["/op_tune_wisdom_t"]
        out.append("/op_wisdom_t")
    dst = os.path.join(ROOT, "tests", "golden", "wisdom_merged_c3_subset.wis")
    open(dst, "w").write("\n".join(out) + "\n")
    print("kept %d of %d wanted ops -> %s (%d bytes)" % (kept, len(want), dst, os.path.getsize(dst)))


if __name__ == "__main__":
    main()
