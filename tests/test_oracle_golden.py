"""CPU suite: the oracle against the reference's own golden vectors and known-answer properties (no GPU).

Golden source: test/good_tr/{sgemm-gen600,sgemm-gen5,conv-gen5,conv-debug,conv-full-gen5,ops-prof-conv-3x3-cudnn-boda}/wisdom.wis
decoded into tests/golden/wisdom_digests.json. Compare rule = the reference's nda_digest_t::mrd_comp at the ops-prof
tolerance 2e-4 (src/rtc_prof.cc:161, src/boda_base.cc:284-311).
"""
import numpy as np
import pytest

TOL = 2e-4


def _check_test(oracle, golden, name, mode):
    n = 0
    for o in golden["tests"][name]:
        op = oracle.parse_op(o["op"])
        outs = oracle.run_op(op, oracle.gen_op_inputs(op, mode))
        for kg in o["kgs"]:
            d = oracle.decode_digest(kg["digest_hex"])
            mine = oracle.make_digest(outs[kg["var"]], d.dim_names, d.seed)
            bad = oracle.digest_mrd_comp(d, mine, TOL)
            assert not bad, (name, o["op"], bad[:4])
            n += 1
    return n


def test_sgemm_gen600_exact(oracle, golden):
    """mode 600: a[k,m]=1000m+k, b=I => c[m,n]=1000m+n exactly; every checksum must be bit-identical."""
    (o,) = golden["tests"]["sgemm-gen600"]
    op = oracle.parse_op(o["op"])
    c = oracle.run_op(op, oracle.gen_op_inputs(op, 600))["c"]
    m, n = np.meshgrid(np.arange(2048), np.arange(2048), indexing="ij")
    assert np.array_equal(c, (1000 * m + n).astype(np.float32))
    d = oracle.decode_digest(o["kgs"][0]["digest_hex"])
    mine = oracle.make_digest(c, d.dim_names, d.seed)
    assert (mine.min_v, mine.max_v) == (d.min_v, d.max_v) == (0.0, 2049047.0)
    assert mine.samps == d.samps  # also pins the mt19937 + boost uniform_int offset restatement


def test_sgemm_gen5(oracle, golden):
    assert _check_test(oracle, golden, "sgemm-gen5", 5) == 1


def test_conv_gen5_and_debug(oracle, golden):
    assert _check_test(oracle, golden, "conv-gen5", 5) == 1
    assert _check_test(oracle, golden, "conv-debug", 5) == 2


def test_conv_full_gen5_all_204_ops(oracle, golden):
    assert _check_test(oracle, golden, "conv-full-gen5", 5) == 204


def test_conv_3x3_cudnn_boda_all_42_ops(oracle, golden):
    assert _check_test(oracle, golden, "ops-prof-conv-3x3-cudnn-boda", 5) == 42


def test_det_hash_rand_c_vs_numpy(oracle):
    ix = np.arange(0, 200000, 7, dtype=np.uint32)
    a = oracle.det_hash_rand_np(ix + np.uint32(234234567))
    b = np.array([oracle.det_hash_rand(int(i) + 234234567) for i in ix[:500]], np.float32)
    assert np.array_equal(a[:500], b)
    assert a.min() >= -5.0 and a.max() <= 5.0
    g = oracle.gen_conv_in(2, 3, 5, 7, 5, 0.0).ravel()
    assert np.array_equal(g, oracle.det_hash_rand_np(np.arange(g.size, dtype=np.uint32) + np.uint32(234234567)))


def test_conv_vs_float64_direct(oracle):
    """Small shapes incl. padding, stride, non-square kernels: oracle vs a float64 numpy direct conv."""
    rng = np.random.RandomState(1)
    for (N, C, H, W, OC, KH, KW, sy, sx, py, px) in [(2, 3, 9, 11, 5, 3, 3, 1, 1, 1, 1), (1, 4, 12, 12, 6, 5, 3, 2, 3, 2, 0),
                                                     (3, 2, 7, 7, 4, 7, 7, 1, 1, 0, 0), (2, 8, 6, 6, 3, 1, 1, 1, 1, 0, 0)]:
        x = rng.randn(N, C, H, W).astype(np.float32)
        w = rng.randn(OC, C, KH, KW).astype(np.float32)
        b = rng.randn(OC).astype(np.float32)
        got = oracle.conv_fwd(x, w, b, (sy, sx), (py, px), relu=False)
        xp = np.zeros((N, C, H + 2 * py, W + 2 * px)); xp[:, :, py:py + H, px:px + W] = x
        OH, OW = got.shape[2:]
        ref = np.zeros(got.shape)
        for oy in range(OH):
            for ox in range(OW):
                patch = xp[:, :, oy * sy:oy * sy + KH, ox * sx:ox * sx + KW]
                ref[:, :, oy, ox] = np.einsum("nchw,ochw->no", patch, w.astype(np.float64)) + b
        assert oracle.mrd_np(ref, got) < 1e-5
        assert np.array_equal(oracle.conv_fwd(x, w, b, (sy, sx), (py, px), relu=True), np.maximum(got, 0))


def test_pool_caffe_semantics(oracle):
    x = np.arange(2 * 3 * 7 * 7, dtype=np.float32).reshape(2, 3, 7, 7) - 50
    mx = oracle.pool_fwd(x, (3, 3), (2, 2), (0, 0), avg_pool=False)
    assert mx.shape == (2, 3, 3, 3)  # ceil((7-3)/2)+1
    assert mx[0, 0, 0, 0] == x[0, 0, :3, :3].max() and mx[1, 2, 2, 2] == x[1, 2, 4:7, 4:7].max()
    x8 = np.arange(8 * 8, dtype=np.float32).reshape(1, 1, 8, 8)
    mx8 = oracle.pool_fwd(x8, (3, 3), (2, 2), (0, 0))
    assert mx8.shape == (1, 1, 4, 4)  # partial last window (Caffe ceil rule, src/conv_util.cc:198-204)
    assert mx8[0, 0, 3, 3] == x8[0, 0, 6:8, 6:8].max()
    av = oracle.pool_fwd(x8, (3, 3), (2, 2), (1, 1), avg_pool=True)
    assert av[0, 0, 0, 0] == pytest.approx(x8[0, 0, 0:2, 0:2].mean())  # padding excluded from the count
    gl = oracle.pool_fwd(x8, None, avg_pool=True)
    assert gl.shape == (1, 1, 1, 1) and gl[0, 0, 0, 0] == pytest.approx(x8.mean())


def test_lrn_vs_closed_form(oracle):
    rng = np.random.RandomState(2)
    x = rng.randn(2, 11, 4, 5).astype(np.float32) * 3
    got = oracle.lrn_fwd(x, 5, 1e-4, 0.75, 1.0)
    sq = np.pad(x.astype(np.float64) ** 2, ((0, 0), (2, 2), (0, 0), (0, 0)))
    ssum = sum(sq[:, i:i + 11] for i in range(5))
    ref = x * (1.0 + 1e-4 / 5 * ssum) ** -0.75
    assert oracle.mrd_np(ref, got) < 1e-5


def test_softmax_relu_concat_reduce(oracle):
    rng = np.random.RandomState(3)
    x = rng.randn(2, 10, 2, 3).astype(np.float32)
    p = oracle.softmax(x)
    e = np.exp(x.astype(np.float64) - np.maximum(x.max(axis=1, keepdims=True), 0))
    assert oracle.mrd_np(e / e.sum(axis=1, keepdims=True), p) < 1e-6
    assert np.array_equal(oracle.relu(x), np.maximum(x, 0))
    a, b = rng.randn(2, 3, 4, 4).astype(np.float32), rng.randn(2, 5, 4, 4).astype(np.float32)
    assert np.array_equal(oracle.concat([a, b]), np.concatenate([a, b], axis=1))
    assert np.array_equal(oracle.reduce_sum([a, a, a]), (a + a) + a)


def test_mrd_metric(oracle):
    a = np.array([0.0, 0.5, 10.0, -100.0], np.float32)
    b = np.array([0.1, 0.5, 11.0, -100.0], np.float32)
    assert oracle.mrd(a, b) == pytest.approx(max(0.1 / 1.0, 1.0 / 11.0), rel=1e-6)
    assert oracle.mrd(a, a) == 0.0
    c = b.copy(); c[1] = np.nan
    assert oracle.mrd(a, c) == float("inf")
    assert oracle.mrd_np(a, b) == pytest.approx(oracle.mrd(a, b))


def test_op_line_grammar_both_syntaxes(oracle):
    cur = ("(str_vals=(type=Convolution),nda_vals=(biases=(dims=(out_chan=96)),filts=(dims=(out_chan=96,in_chan=3,y=11,x=11)),"
           "in=(dims=(img=20,chan=3,y=227,x=227)),in_pad=(tn=none,dims=(y=0,x=0)),kern_sz=(tn=none,dims=(y=11,x=11)),"
           "out=(dims=(img=20,chan=96,y=55,x=55)),out_chans=(tn=uint32_t,v=96),stride=(tn=none,dims=(y=4,x=4))))")
    stale = ("(type=Convolution,dims_vals=(biases=(out_chan=96),filts=(out_chan=96,in_chan=3,y=11,x=11),in=(img=20,chan=3,y=227,x=227),"
             "in_pad=(y=0,x=0),kern_sz=(y=11,x=11),out=(img=20,chan=96,y=55,x=55),stride=(y=4,x=4)),str_vals=(out_chans=96))")
    a, b = oracle.parse_op(cur), oracle.parse_op(stale)
    assert a.type == b.type == "Convolution"
    for k in a.nda_vals:
        assert a.nda_vals[k].dims == b.nda_vals[k].dims and a.nda_vals[k].tn == b.nda_vals[k].tn, k
    assert a.get_u32("out_chans") == b.get_u32("out_chans") == 96
    assert oracle.op_flops(a) == 2.0 * 20 * 96 * 55 * 55 * 3 * 11 * 11
    s = oracle.parse_op("(str_vals=(type=sgemm),nda_vals=(a=(dims=(K=128,M=128)),b=(dims=(K=128,N=128)),c=(dims=(M=128,N=128))))")
    assert oracle.op_flops(s) == 2.0 * 128 ** 3
    with pytest.raises(ValueError):
        oracle.parse_op("(str_vals=(type=sgemm),bogus=(x=1))")  # NESI rejects unused fields


def test_c1_sgemm_128_plumbing(oracle):
    """BASELINE config C1: the single 128^3 SGEMM of test/sgemm-ops-tiny.txt through the CPU comp_util path."""
    op = oracle.parse_op("(type=sgemm,dims_vals=(a=(K=128,M=128),b=(K=128,N=128),c=(M=128,N=128)))")
    ins = oracle.gen_op_inputs(op, 5)
    c = oracle.run_op(op, ins)["c"]
    ref = ins["a"].astype(np.float64).T @ ins["b"].astype(np.float64)
    assert oracle.mrd(ref.astype(np.float32), c) < 2e-4
    assert np.array_equal(ins["a"], ins["b"])  # same salt for a and b
