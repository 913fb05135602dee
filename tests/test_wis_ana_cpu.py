"""Host-only tests of the wisdom reader + wis-ana restatement (boda_b200/csrc/wisdom.cu; reference: src/op-tuner.cc:17-93 reader, :135-143
filter_runs, :204-396 wis_ana_t). Golden data: tests/golden/wisdom_merged_c3_subset.wis, cut from the reference's published database
test/wisdom-merged.wis by tests/golden/make_wisdom_subset.py."""
import math
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDNN = "(use_be=nvrtc,use_culibs=1,MNt=8 8,MNb=8 16,tconv_max_ksz=11 11)"


@pytest.fixture(scope="module")
def bb():
    import boda_b200
    boda_b200.lib()
    return boda_b200


def _op(img, oc=16):
    return ("(str_vals=(type=Convolution),nda_vals=(biases=(dims=(out_chan=%d)),filts=(dims=(out_chan=%d,in_chan=8,y=3,x=3)),in=(dims=(img=%d,chan=8,y=10,x=10)),"
            "in_pad=(tn=none,dims=(y=1,x=1)),kern_sz=(tn=none,dims=(y=3,x=3)),out=(dims=(img=%d,chan=%d,y=10,x=10)),out_chans=(tn=uint32_t,v=%d),"
            "stride=(tn=none,dims=(y=1,x=1))))" % (oc, oc, img, img, oc, oc))


def _wis(op, runs):
    """runs: [(tune, plat, secs or None for an error run)]; one op_wisdom_t record in the reference's text format"""
    out = ["op_wisdom_t", op]
    by_tune = {}
    for t, p, s in runs:
        by_tune.setdefault(t, []).append((p, s))
    for t, rs in by_tune.items():
        out += ["op_tune_wisdom_t", t]
        for p, s in rs:
            out += ["op_run_t", p, "nan" if s is None else repr(s), "profile call failure: boom" if s is None else ""] + ([] if s is None else [op])
        out.append("/op_tune_wisdom_t")
    out.append("/op_wisdom_t")
    return "\n".join(out) + "\n"


def test_aom_pom_ref_semantics(bb):
    a, b, c = _op(1), _op(5), _op(5, oc=32)
    text = (_wis(a, [("(t=1)", "gpu:X", 3.0), ("(t=2)", "gpu:X", 2.0), ("(t=ref)", "gpu:X", 1.5), ("(t=1)", "gpu:Y", 0.1)]) +
            _wis(b, [("(t=1)", "gpu:X", 1.0), ("(t=2)", "gpu:X", None), ("(t=ref)", "gpu:X", 4.0)]) +
            _wis(c, [("(t=1)", "gpu:X", 5.0), ("(t=2)", "gpu:X", 1.0)]))
    r = bb.wis_ana(text, s_plat="gpu:X", ref_tune="(t=ref)")
    rows = {x["op"]: x for x in r["rows"]}
    # t=1 handled 3 ops (9.0 s), t=2 only 2 (one error run dropped): most-cases-first makes t=1 the best overall tune despite its larger total
    assert r["aom_tune"] == "(t=1)" and r["tot_runs"] == 5
    assert (rows[a]["aom"], rows[a]["pom"], rows[a]["ref"], rows[a]["pom_tune"]) == (3.0, 2.0, 1.5, "(t=2)")
    assert (rows[b]["aom"], rows[b]["pom"], rows[b]["ref"]) == (1.0, 1.0, 4.0)
    assert (rows[c]["aom"], rows[c]["pom"]) == (5.0, 1.0) and math.isnan(rows[c]["ref"])
    assert rows[a]["flops"] == 2 * 1 * 100 * 16 * 72 and rows[c]["flops"] == 2 * 5 * 100 * 32 * 72
    # platform regex: on gpu:Y only op a has a run
    ry = bb.wis_ana(text, s_plat="gpu:Y$")
    assert [x["pom"] for x in ry["rows"] if not math.isnan(x["pom"])] == [0.1] and ry["tot_runs"] == 1
    # s_img and min_flops filters are permanent (the op does not appear at all)
    assert [x["op"] for x in bb.wis_ana(text, s_img=5, s_plat="gpu:X")["rows"]] == sorted([b, c])
    assert [x["op"] for x in bb.wis_ana(text, s_plat="gpu:X", min_flops=rows[c]["flops"])["rows"]] == [c]
    # the csv wis-plot.py reads
    csv = bb.wis_ana(text, s_plat="gpu:X", ref_tune="(t=ref)", csv=True).splitlines()
    assert csv[0] == "OP FLOPS boda-manual-tune boda-autotuned REF" and len(csv) == 4
    assert csv[1 + sorted([a, b, c]).index(c)] == "%s %d 5 1 nan" % (c, rows[c]["flops"])


def test_reader_errors_and_records_written_by_this_repo(bb):
    a = _op(2)
    with pytest.raises(bb.RtException, match="unknown op_wisdom_t text format stream command"):
        bb.wis_ana("op_wisdom_t\n%s\nbogus\n/op_wisdom_t\n" % a)
    with pytest.raises(bb.RtException, match="expected a line with 'op_wisdom_t'"):
        bb.wis_ana("op_wisdumb_t\n")
    with pytest.raises(bb.RtException, match="got EOF"):
        bb.wis_ana("op_wisdom_t\n%s\nop_tune_wisdom_t\n(t=1)\nop_run_t\ngpu:X\n" % a)
    with pytest.raises(bb.RtException, match="duplicate op"):
        bb.wis_ana(_wis(a, [("(t=1)", "gpu:X", 1.0)]) * 2)
    with pytest.raises(bb.RtException):
        bb.wis_ana(_wis(a, [("(t=1)", "gpu:X", 1.0)]).replace("\n1.0\n", "\nfast\n"))  # time is not a number
    # records produced by b200_wisdom_record (incl. a known-good digest block, which the analysis skips) read back
    rec = bb.wisdom_record(a, [("out", "00AB")], "(use_be=b200,prec=fp32)", "b200:NVIDIA B200", 2.5e-5)
    r = bb.wis_ana(rec + bb.wisdom_record(_op(3), [], "(use_be=b200,prec=fp32)", "b200:NVIDIA B200", 0.0, err="unsupported: nope"), s_plat="b200:")
    assert [(x["op"], x["pom"]) for x in r["rows"] if not math.isnan(x["pom"])] == [(a, 2.5e-5)] and len(r["rows"]) == 2


def test_reference_database_subset(bb):
    """9 of the C3 ops (B=20) on the reference's Titan X: the analysis against an independent reading of the same records."""
    text = open(os.path.join(ROOT, "tests", "golden", "wisdom_merged_c3_subset.wis")).read()
    r = bb.wis_ana(text, s_img=20, s_plat="nvrtc:GeForce GTX TITAN X", ref_tune=CUDNN)
    assert len(r["rows"]) == 9 and r["tot_runs"] == 9 * 6  # per op: six Boda nvrtc tunes take part, the seventh (cuDNN) is REF
    assert r["aom_tune"] == "(use_be=nvrtc,MNt=8 8,MNb=8 16,k1conv=1,tconv=1,tconv_max_ksz=11 11)"  # the tests' "opt" tune (src/test_compute.cc:239)
    # independent parse: {op: {tune: secs}}
    lines, db, i = text.split("\n"), {}, 0
    while i < len(lines):
        if lines[i] == "op_wisdom_t":
            op = lines[i + 1]
            db[op] = {}
        elif lines[i] == "op_tune_wisdom_t":
            tune = lines[i + 1]
        elif lines[i] == "op_run_t" and lines[i + 3] == "":
            db[op][tune] = float(lines[i + 2])
        i += 1
    for row in r["rows"]:
        t = dict(db[row["op"]])
        assert row["ref"] == t.pop(CUDNN)
        assert row["pom"] == min(t.values()) and t[row["pom_tune"]] == row["pom"] and row["aom"] == t[r["aom_tune"]]
    # known answers: AlexNet conv1 at B=20 took 1.58 ms with cuDNN v5 and 1.46 ms with Boda's best tune on that Titan X (2.7 / 2.9 TF/s)
    conv1 = next(x for x in r["rows"] if "y=11,x=11" in x["op"])
    assert conv1["flops"] == 2 * 20 * 55 * 55 * 96 * 363 and abs(conv1["ref"] - 1.581e-3) < 1e-6 and abs(conv1["pom"] - 1.464e-3) < 1e-6
