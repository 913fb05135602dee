"""CPU tests of the wisdom-file writer (boda_b200/csrc/wisdom.cu): the product's own restatement of nda_digest_t
(std::hash seed, mt19937 + boost uniform_int offsets, float checksums, bwrite layout) against the reference's golden bytes and the oracle."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def bb():
    import boda_b200 as m
    m.lib()
    return m


def test_digest_writer_reproduces_the_reference_golden_bytes(bb, oracle, golden):
    """sgemm-gen600 has exact integer outputs (c[m,n] = 1000 m + n), so the digest the reference committed in
    test/good_tr/sgemm-gen600/wisdom.wis must come out BYTE FOR BYTE: seed = std::hash("c"), 60 offsets, min/max, every checksum."""
    rec = golden["tests"]["sgemm-gen600"][0]
    kg = rec["kgs"][0]
    c = oracle.sgemm(oracle.gen_sgemm_a(2048, 2048, 600), oracle.gen_sgemm_b(2048, 2048, 600))
    assert bb.nda_digest_hex(kg["var"], c, ["M", "N"]) == kg["digest_hex"].strip().upper()


@pytest.mark.parametrize("shape,names,var", [((3, 5, 7, 4), ["img", "chan", "y", "x"], "out"), ((1, 1, 1, 1), ["img", "chan", "y", "x"], "out"),
                                             ((70, 33), ["M", "N"], "c"), ((2, 96, 13, 13), ["img", "chan", "y", "x"], "out")])
def test_digest_writer_agrees_with_the_oracle_reader(bb, oracle, shape, names, var):
    rng = np.random.RandomState(sum(shape))
    a = (rng.randn(*shape) * 3).astype(np.float32)
    d = oracle.decode_digest(bb.nda_digest_hex(var, a, names))
    assert d.dim_names == names and d.sizes == list(shape)
    ref = oracle.make_digest(a, names, d.seed)
    assert d.min_v == ref.min_v and d.max_v == ref.max_v and d.samps == ref.samps  # same offsets (same RNG restatement), same float sums
    if var in ("c", "out"):  # the seeds the reference's golden files carry for these names
        assert d.seed == {"c": 10959529184379665549, "out": 470894893395316877}[var]


def test_wisdom_record_format(bb, oracle, golden):
    """One record, read back the way src/op-tuner.cc:70-96 (read_next_wisdom) does."""
    rec = golden["tests"]["conv-gen5"][0]
    hexd = rec["kgs"][0]["digest_hex"]
    txt = bb.wisdom_record(rec["op"], [("out", hexd)], "(use_be=b200,prec=fp32)", "b200:NVIDIA B200", 4.2e-05)
    lines = txt.split("\n")
    assert lines[0] == "op_wisdom_t" and lines[1] == rec["op"] and lines[2:5] == ["kg", "out", hexd]
    assert lines[5:7] == ["op_tune_wisdom_t", "(use_be=b200,prec=fp32)"] and lines[7:9] == ["op_run_t", "b200:NVIDIA B200"]
    assert abs(float(lines[9]) - 4.2e-05) < 1e-12 and lines[10] == "" and lines[11] == rec["op"]
    assert lines[12:14] == ["/op_tune_wisdom_t", "/op_wisdom_t"] and lines[14] == ""
    assert oracle.parse_op(lines[1]).type == "Convolution"
    err = bb.wisdom_record(rec["op"], [], "(use_be=b200)", "b200:x", 0.0, err="unsupported shape").split("\n")
    assert err[4:6] == ["op_run_t", "b200:x"] and err[7] == "unsupported shape" and err[8] == "/op_tune_wisdom_t"  # no op line after an error
