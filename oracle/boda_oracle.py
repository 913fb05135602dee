"""CPU oracle for Boda's rtc_fwd hot path -- TEST INFRASTRUCTURE ONLY.

Python face of oracle/boda_oracle.c (ctypes) plus the pieces of the reference's test harness that are pure
bookkeeping: the op-line text grammar, the `mrd` compare and the `nda_digest_t` golden-vector format.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

Parity status: PINNED -- tests/test_oracle_golden.py checks this oracle against every decodable golden digest
the reference's own tests hold for the path (tests/golden/wisdom_digests.json, generated from
test/good_tr/*/wisdom.wis by tests/golden/make_golden.py).

Reference citations are relative to the reference root (moskewcz/boda @ 27ec86e).
"""
from __future__ import annotations

import ctypes
import math
import os
import struct
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libboda_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64


def build(force: bool = False) -> str:
    """Compile oracle/boda_oracle.c -> oracle/libboda_oracle.so (gcc, OpenMP)."""
    src = os.path.join(_HERE, "boda_oracle.c")
    if force or (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libboda_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oracle_det_hash_rand.restype = ctypes.c_float
        _lib.oracle_det_hash_rand.argtypes = [_u32]
        _lib.oracle_strided_sum.restype = ctypes.c_float
        _lib.oracle_strided_sum.argtypes = [_f32p, _u64, _u64, _u64]
        _lib.oracle_conv_fwd.restype = ctypes.c_int
        _lib.oracle_conv_fwd_acc64.restype = ctypes.c_int
        _lib.oracle_num_threads.restype = ctypes.c_int
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    """OpenMP team size of the C restatement (bench.py's reference arm: all host cores, whatever OMP_NUM_THREADS the launcher exported)."""
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


# ---------------------------------------------------------------------------------------------------------------
# gen_data (test/rtc/gen-util.h, test/rtc/gen_data_*.cucl)
# ---------------------------------------------------------------------------------------------------------------

def det_hash_rand(ix: int) -> float:
    return float(lib().oracle_det_hash_rand(_u32(ix & 0xFFFFFFFF)))


def det_hash_rand_np(ix: np.ndarray) -> np.ndarray:
    """Vectorised numpy restatement of test/rtc/gen-util.h:1-9 (independent of the C one; cross-checked in tests)."""
    h = ix.astype(np.uint32)
    h ^= h >> np.uint32(16)
    h = (h * np.uint32(0x85EBCA6B)).astype(np.uint32)
    h ^= h >> np.uint32(13)
    h = (h * np.uint32(0xC2B2AE35)).astype(np.uint32)
    h ^= h >> np.uint32(16)
    scale = np.float64(np.float32(10.0) / np.float32(4294967295.0))  # == 10 * 2^-32 exactly
    # fused multiply-add semantics: the product and sum are exact in float64, so one rounding to float32 remains
    return (h.astype(np.float32).astype(np.float64) * scale - 5.0).astype(np.float32)


def gen_conv_in(img, chan, y, x, mode=5, vi=0.0) -> np.ndarray:
    a = np.empty((img, chan, y, x), np.float32)
    lib().oracle_gen_conv_in(_p(a), _u32(img), _u32(chan), _u32(y), _u32(x), _u32(mode), ctypes.c_float(vi))
    return a


def gen_conv_filts(out_chan, in_chan, y, x, mode=5, vi=0.0) -> np.ndarray:
    a = np.empty((out_chan, in_chan, y, x), np.float32)
    lib().oracle_gen_conv_filts(_p(a), _u32(out_chan), _u32(in_chan), _u32(y), _u32(x), _u32(mode), ctypes.c_float(vi))
    return a


def gen_conv_biases(out_chan, mode=5, vi=0.0) -> np.ndarray:
    a = np.empty((out_chan,), np.float32)
    lib().oracle_gen_conv_biases(_p(a), _u32(out_chan), _u32(mode), ctypes.c_float(vi))
    return a


def gen_sgemm_a(K, M, mode=5, vi=0.0) -> np.ndarray:
    a = np.empty((K, M), np.float32)
    lib().oracle_gen_sgemm_a(_p(a), _u32(K), _u32(M), _u32(mode), ctypes.c_float(vi))
    return a


def gen_sgemm_b(K, N, mode=5, vi=0.0) -> np.ndarray:
    a = np.empty((K, N), np.float32)
    lib().oracle_gen_sgemm_b(_p(a), _u32(K), _u32(N), _u32(mode), ctypes.c_float(vi))
    return a


# ---------------------------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------------------------

def conv_out_sz(in_sz, pad, stride, kern):
    """src/conv_util.cc:167-173 (conv_in_sz_to_out_sz); 0 if the padded input is smaller than the kernel."""
    p = in_sz + 2 * pad
    return 0 if p < kern else (p - kern) // stride + 1


def pool_out_sz(in_sz, pad, stride, kern):
    """src/conv_util.cc:198-204: Caffe convention, any partial window makes an output (ceil)."""
    p = in_sz + 2 * pad
    return 1 if p < kern else -((p - kern) // -stride) + 1


def conv_fwd(inp, filts, biases, stride=(1, 1), in_pad=(0, 0), relu=True, acc64=False) -> np.ndarray:
    """acc64=True: same algorithm and order with a double accumulator (the reference's math without fp32 accumulation noise)."""
    inp = np.ascontiguousarray(inp, np.float32)
    filts = np.ascontiguousarray(filts, np.float32)
    biases = np.ascontiguousarray(biases, np.float32)
    N, C, H, W = inp.shape
    OC, IC, KH, KW = filts.shape
    assert IC == C and biases.shape == (OC,)
    OH, OW = conv_out_sz(H, in_pad[0], stride[0], KH), conv_out_sz(W, in_pad[1], stride[1], KW)
    assert OH > 0 and OW > 0
    out = np.empty((N, OC, OH, OW), np.float32)
    r = (lib().oracle_conv_fwd_acc64 if acc64 else lib().oracle_conv_fwd)(_p(inp), _p(filts), _p(biases), _p(out), _u32(N), _u32(C), _u32(H), _u32(W), _u32(OC),
                              _u32(KH), _u32(KW), _u32(stride[0]), _u32(stride[1]), _u32(in_pad[0]), _u32(in_pad[1]),
                              ctypes.c_int(1 if relu else 0))
    assert r == 0
    return out


def sgemm(a, b, acc64=False) -> np.ndarray:
    """c[M,N] = a[K,M]^T . b[K,N] (test/rtc/sgemm.cucl:1-3: `a` is stored K:M)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    K, M = a.shape
    K2, N = b.shape
    assert K == K2
    c = np.empty((M, N), np.float32)
    (lib().oracle_sgemm_acc64 if acc64 else lib().oracle_sgemm)(_p(a), _p(b), _p(c), _u32(M), _u32(N), _u32(K))
    return c


def pool_fwd(inp, kern_sz=None, stride=(1, 1), in_pad=(0, 0), avg_pool=False) -> np.ndarray:
    inp = np.ascontiguousarray(inp, np.float32)
    N, C, H, W = inp.shape
    if kern_sz is None:  # global pooling: src/cnn_op.cc:39-45 sets kern_sz to the input size
        kern_sz = (H, W)
        OH = OW = 1
    else:
        OH, OW = pool_out_sz(H, in_pad[0], stride[0], kern_sz[0]), pool_out_sz(W, in_pad[1], stride[1], kern_sz[1])
    out = np.empty((N, C, OH, OW), np.float32)
    lib().oracle_pool_fwd(_p(inp), _p(out), _u32(N), _u32(C), _u32(H), _u32(W), _u32(OH), _u32(OW), _u32(kern_sz[0]),
                          _u32(kern_sz[1]), _u32(stride[0]), _u32(stride[1]), _u32(in_pad[0]), _u32(in_pad[1]),
                          ctypes.c_int(1 if avg_pool else 0))
    return out


def lrn_fwd(inp, local_size=5, alpha=1.0, beta=0.75, k=1.0) -> np.ndarray:
    inp = np.ascontiguousarray(inp, np.float32)
    N, C, H, W = inp.shape
    out = np.empty_like(inp)
    lib().oracle_lrn_fwd(_p(inp), _p(out), _u32(N), _u32(C), _u32(H), _u32(W), _u32(local_size), ctypes.c_float(alpha),
                         ctypes.c_float(beta), ctypes.c_float(k))
    return out


def relu(x) -> np.ndarray:
    out = np.array(x, np.float32, copy=True, order="C")
    lib().oracle_relu(_p(out), _u64(out.size))
    return out


def softmax(inp) -> np.ndarray:
    inp = np.ascontiguousarray(inp, np.float32)
    N, C, H, W = inp.shape
    out = np.empty_like(inp)
    lib().oracle_softmax(_p(inp), _p(out), _u32(N), _u32(C), _u32(H), _u32(W))
    return out


def concat(ins: List[np.ndarray]) -> np.ndarray:
    N, _, H, W = ins[0].shape
    out_C = sum(i.shape[1] for i in ins)
    out = np.zeros((N, out_C, H, W), np.float32)
    ocix = 0
    for i in ins:
        i = np.ascontiguousarray(i, np.float32)
        lib().oracle_concat_copy(_p(i), _p(out), _u32(N), _u32(i.shape[1]), _u32(H), _u32(W), _u32(out_C), _u32(ocix))
        ocix += i.shape[1]
    return out


def reduce_sum(ins: List[np.ndarray]) -> np.ndarray:
    ins = [np.ascontiguousarray(i, np.float32) for i in ins]
    out = np.empty_like(ins[0])
    arr = (_f32p * len(ins))(*[_p(i) for i in ins])
    lib().oracle_reduce_sum(arr, _u32(len(ins)), _p(out), _u64(out.size))
    return out


# ---------------------------------------------------------------------------------------------------------------
# compare metric (src/boda_base.cc:140-206, src/comp_util.cc:21-57)
# ---------------------------------------------------------------------------------------------------------------

def ssds_diff(o1: np.ndarray, o2: np.ndarray) -> Dict[str, float]:
    o1 = np.ascontiguousarray(o1, np.float32).ravel()
    o2 = np.ascontiguousarray(o2, np.float32).ravel()
    assert o1.size == o2.size
    res = (ctypes.c_double * 8)()
    lib().oracle_ssds_diff(_p(o1), _p(o2), _u64(o1.size), res)
    keys = ["ssds", "sds", "mad", "mrd", "sum1", "sum2", "num_diff", "has_nan"]
    return dict(zip(keys, [float(v) for v in res]))


def mrd(o1, o2) -> float:
    """max over elements of |a-b| / max(1,|a|,|b|); NaN-poisoned results return inf (src/comp_util.cc:38)."""
    d = ssds_diff(o1, o2)
    return float("inf") if d["has_nan"] or math.isnan(d["mrd"]) else d["mrd"]


def mrd_np(o1, o2) -> float:
    a = np.asarray(o1, np.float64).ravel()
    b = np.asarray(o2, np.float64).ravel()
    if np.isnan(a).any() or np.isnan(b).any():
        return float("inf")
    return float(np.max(np.abs(b - a) / np.maximum(1.0, np.maximum(np.abs(a), np.abs(b))))) if a.size else 0.0


# ---------------------------------------------------------------------------------------------------------------
# op-line text grammar (lexp; src/lexp.cc:22-30, src/nesi.cc:661-785, src/op_base.H:12-13)
# ---------------------------------------------------------------------------------------------------------------

def parse_lexp(s: str):
    """Parse `(k=v,k=(...),...)` into nested dicts (order kept); leaves are strings. `\\` escapes the next char."""
    pos = 0
    s = s.strip()

    def parse_val():
        nonlocal pos
        if pos < len(s) and s[pos] == "(":
            pos += 1
            d = {}
            while True:
                if pos >= len(s):
                    raise ValueError("lexp: unterminated list")
                if s[pos] == ")":
                    pos += 1
                    return d
                # key
                k = []
                while pos < len(s) and s[pos] not in "=,()":
                    if s[pos] == "\\":
                        pos += 1
                    k.append(s[pos]); pos += 1
                if pos >= len(s) or s[pos] != "=":
                    raise ValueError("lexp: expected '=' after key %r at %d" % ("".join(k), pos))
                pos += 1
                key = "".join(k)
                if key in d:
                    raise ValueError("lexp: duplicate key %r" % key)
                d[key] = parse_val()
                if pos < len(s) and s[pos] == ",":
                    pos += 1
        leaf = []
        while pos < len(s) and s[pos] not in ",()":
            if s[pos] == "\\":
                pos += 1
            leaf.append(s[pos]); pos += 1
        return "".join(leaf)

    v = parse_val()
    if pos != len(s):
        raise ValueError("lexp: trailing garbage at %d" % pos)
    return v


class Nda:
    """dims (ordered name->size), type name and optional values; mirrors nda_t for op descriptors."""

    def __init__(self, dims: Optional[Dict[str, int]] = None, tn: str = "float", v=None):
        self.dims = dict(dims or {})
        self.tn = tn
        self.v = v

    def dsz(self, name: str) -> int:
        return self.dims[name]

    def shape(self) -> Tuple[int, ...]:
        return tuple(self.dims.values())

    def __repr__(self):
        return "Nda(%s,tn=%s,v=%s)" % (self.dims, self.tn, self.v)


def _nda_from_lexp(d) -> Nda:
    """src/nesi.cc:720-785: (tn=..,dims=(..),v=a:b:c); tn defaults to float; dims may carry __tn__."""
    if not isinstance(d, dict):
        raise ValueError("nda must be a list")
    unknown = set(d) - {"tn", "dims", "v"}
    if unknown:
        raise ValueError("nda: unused fields %s" % sorted(unknown))
    tn = d.get("tn", "float")
    dims = {}
    for k, v in (d.get("dims") or {}).items():
        if k == "__tn__":
            tn = v
        else:
            dims[k] = int(v)
    vals = None
    if "v" in d:
        parts = d["v"].split(":")
        vals = [float(p) if tn in ("float", "double", "half") else int(p) for p in parts]
        if not dims and len(vals) == 1:
            vals = vals[0]
    return Nda(dims, tn, vals)


class Op:
    """op_base_t: str_vals + nda_vals (src/op_base.H:9-41)."""

    def __init__(self, str_vals: Dict[str, str], nda_vals: Dict[str, Nda]):
        self.str_vals = dict(str_vals)
        self.nda_vals = dict(nda_vals)

    @property
    def type(self) -> str:
        return self.str_vals["type"]

    def has(self, an):
        return an in self.nda_vals

    def get_dims(self, an) -> Nda:
        return self.nda_vals[an]

    def get_u32(self, an) -> int:
        return int(self.nda_vals[an].v)

    def pt(self, an, default):
        if an not in self.nda_vals:
            return default
        d = self.nda_vals[an].dims
        return (d["y"], d["x"])


def parse_op(line: str) -> Op:
    """Accepts the current syntax (str_vals/nda_vals) and the stale one (type/dims_vals/str_vals.out_chans)."""
    d = parse_lexp(line)
    if not isinstance(d, dict):
        raise ValueError("op line must be a list")
    if "dims_vals" in d or "type" in d:  # stale syntax, SURVEY Appendix A translation rule
        unknown = set(d) - {"type", "dims_vals", "str_vals"}
        if unknown:
            raise ValueError("op: unused fields %s" % sorted(unknown))
        str_vals = {"type": d["type"]}
        nda_vals = {}
        for k, v in (d.get("dims_vals") or {}).items():
            tn = "none" if k in ("in_pad", "kern_sz", "stride") else "float"
            dims = {}
            for dk, dv in v.items():
                if dk == "__tn__":
                    tn = dv
                else:
                    dims[dk] = int(dv)
            nda_vals[k] = Nda(dims, tn)
        for k, v in (d.get("str_vals") or {}).items():
            if k == "out_chans":
                nda_vals["out_chans"] = Nda({}, "uint32_t", int(v))
            else:
                str_vals[k] = v
        return Op(str_vals, nda_vals)
    unknown = set(d) - {"str_vals", "nda_vals"}
    if unknown:
        raise ValueError("op: unused fields %s" % sorted(unknown))  # NESI rejects unused fields (src/nesi.cc:25-35)
    str_vals = dict(d.get("str_vals") or {})
    nda_vals = {k: _nda_from_lexp(v) for k, v in (d.get("nda_vals") or {}).items()}
    return Op(str_vals, nda_vals)


def op_flops(op: Op) -> float:
    """Reference's FLOP convention: conv 2*B*OC*OH*OW*IC*KH*KW, sgemm 2*M*N*K (src/latex-util.H:116-133)."""
    if op.type == "Convolution":
        o, f = op.get_dims("out").dims, op.get_dims("filts").dims
        return 2.0 * o["img"] * o["chan"] * o["y"] * o["x"] * f["in_chan"] * f["y"] * f["x"]
    if op.type == "sgemm":
        a, b = op.get_dims("a").dims, op.get_dims("b").dims
        return 2.0 * a["M"] * b["N"] * a["K"]
    return 0.0


def gen_op_inputs(op: Op, mode: int = 5, vi: float = 0.0) -> Dict[str, np.ndarray]:
    """The ops-prof flow: every IN arg is filled by gen_data_<optype>_<arg> on its *_ref dims (src/rtc_prof.cc:73-90)."""
    if op.type == "Convolution":
        i, f = op.get_dims("in").dims, op.get_dims("filts").dims
        return {
            "in": gen_conv_in(i["img"], i["chan"], i["y"], i["x"], mode, vi),
            "filts": gen_conv_filts(f["out_chan"], f["in_chan"], f["y"], f["x"], mode, vi),
            "biases": gen_conv_biases(f["out_chan"], mode, vi),
        }
    if op.type == "sgemm":
        a, b = op.get_dims("a").dims, op.get_dims("b").dims
        return {"a": gen_sgemm_a(a["K"], a["M"], mode, vi), "b": gen_sgemm_b(b["K"], b["N"], mode, vi)}
    raise ValueError("gen_op_inputs: unhandled op type " + op.type)


def run_op(op: Op, ins: Dict[str, np.ndarray], acc64: bool = False) -> Dict[str, np.ndarray]:
    """Per-op flow of ops-prof: conv -> {"out"} with conv_has_relu forced to 1 (src/cnn_op.cc:337); sgemm -> {"c"}."""
    if op.type == "Convolution":
        out = conv_fwd(ins["in"], ins["filts"], ins["biases"], op.pt("stride", (1, 1)), op.pt("in_pad", (0, 0)), relu=True, acc64=acc64)
        if op.has("out"):
            assert out.shape == op.get_dims("out").shape(), (out.shape, op.get_dims("out").shape())
        return {"out": out}
    if op.type == "sgemm":
        return {"c": sgemm(ins["a"], ins["b"], acc64=acc64)}
    raise ValueError("run_op: unhandled op type " + op.type)


# ---------------------------------------------------------------------------------------------------------------
# nda_digest_t (src/boda_base.cc:210-383)
# ---------------------------------------------------------------------------------------------------------------

class MT19937:
    """32-bit Mersenne twister, seeded like boost::random::mt19937(value) (value truncated to 32 bits)."""

    def __init__(self, seed: int):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            self.mt[i] = (1812433253 * (self.mt[i - 1] ^ (self.mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.idx = 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def next(self) -> int:
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def _boost_uniform_u64(gen: MT19937, lo: int, hi: int) -> int:
    """boost::random::uniform_int_distribution<uint64_t>(lo,hi) over a 32-bit engine (boost/random/uniform_int_distribution.hpp,
    generate_uniform_int): range 0 consumes nothing; range < 2^32-1 uses bucketed rejection; larger ranges compose
    several engine calls (not needed for tensors < 2^32 elements, src/boda_base.H:608)."""
    rng = hi - lo
    brange = 0xFFFFFFFF
    if rng == 0:
        return lo
    if rng == brange:
        return gen.next() + lo
    if rng > brange:  # multi-call composition path: only for tensors >= 2^32 elements, which dims_t cannot hold
        raise NotImplementedError("uniform_int range > 2^32-1")
    bucket = brange // (rng + 1)
    if brange % (rng + 1) == rng:
        bucket += 1
    while True:
        r = gen.next() // bucket
        if r <= rng:
            return r + lo


def _floor_log2_u64(v: int) -> int:
    return v.bit_length() - 1


def row_major_strides(sizes: List[int]) -> List[int]:
    st = [1] * len(sizes)
    for i in range(len(sizes) - 2, -1, -1):
        st[i] = st[i + 1] * sizes[i + 1]
    return st


def digest_sample_infos(sizes: List[int], seed: int) -> List[Tuple[int, int, int]]:
    """(stride, offset, num_subsamps) list: src/boda_base.cc:221-252."""
    strides = row_major_strides(sizes)
    total = int(np.prod(sizes)) if sizes else 1
    ss = set(p for p in (1, 2, 3, 5, 7, 11, 13, 17, 19, 23, 29) if p <= total)
    ss.update(strides)
    ss.add(total)
    gen = MT19937(seed)
    sis = []
    for stride in sorted(ss):
        assert 0 < stride <= total
        num_offsets = _floor_log2_u64(stride + 1)
        seen = set()
        for _ in range(num_offsets):
            off = _boost_uniform_u64(gen, 0, stride - 1)
            if off in seen:
                continue
            seen.add(off)
            sis.append((stride, off, (total - off) // stride))
    return sis


class Digest:
    def __init__(self, tn, dim_names, sizes, strides, seed, min_v, max_v, samps, self_cmp_mrd=0.0):
        self.tn, self.dim_names, self.sizes, self.strides = tn, list(dim_names), list(sizes), list(strides)
        self.seed, self.min_v, self.max_v, self.samps, self.self_cmp_mrd = seed, min_v, max_v, list(samps), self_cmp_mrd

    def total(self) -> int:
        return int(np.prod(self.sizes))


def decode_digest(hexstr: str) -> Digest:
    """Binary layout: SURVEY Appendix D (src/boda_base.cc:329-363; bwrite of string/vector/dims_t in src/boda_base.H:319-417)."""
    b = bytes.fromhex(hexstr.strip())
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, b, pos)
        pos += struct.calcsize("<" + fmt)
        return v[0]

    def rd_str():
        nonlocal pos
        n = rd("I")
        s = b[pos:pos + n].decode()
        pos += n
        return s

    non_null = rd("B")
    assert non_null == 1
    tn = rd_str()
    assert tn == "float", tn
    magic = rd("I")
    assert magic == 0xDADA0101, hex(magic)
    self_cmp_mrd = rd("d")
    nd = rd("I")
    sizes, strides, names = [], [], []
    for _ in range(nd):
        sizes.append(rd("I")); strides.append(rd("I")); names.append(rd_str())
    tn2 = rd_str()
    assert tn2 == tn
    strides_sz = rd("Q")
    strides_valid = rd("B")
    assert strides_valid == 1 and strides_sz == int(np.prod(sizes))
    seed = rd("Q")
    min_v = rd("f")
    max_v = rd("f")
    n = rd("I")
    samps = [rd("f") for _ in range(n)]
    assert pos == len(b), (pos, len(b))
    return Digest(tn, names, sizes, strides, seed, min_v, max_v, samps, self_cmp_mrd)


def make_digest(arr: np.ndarray, dim_names: List[str], seed: int) -> Digest:
    """nda_digest_T::set_from_nda (src/boda_base.cc:249-262)."""
    a = np.ascontiguousarray(arr, np.float32)
    flat = a.ravel()
    mn, mx = ctypes.c_float(), ctypes.c_float()
    lib().oracle_min_max(_p(flat), _u64(flat.size), ctypes.byref(mn), ctypes.byref(mx))
    sis = digest_sample_infos(list(a.shape), seed)
    samps = [float(lib().oracle_strided_sum(_p(flat), _u64(flat.size), _u64(off), _u64(st))) for (st, off, _) in sis]
    return Digest("float", dim_names, list(a.shape), row_major_strides(list(a.shape)), seed, float(mn.value),
                  float(mx.value), samps)


def _rel_diff(v1, v2):
    a = max(1.0, abs(v1), abs(v2))
    return abs(v2 - v1) / a


def digest_mrd_comp(d1: Digest, d2: Digest, tol: float) -> List[str]:
    """mrd_comp (src/boda_base.cc:284-311): empty list == match. Checksums covering >1000 elements get sqrt(n/1000) slack."""
    if d1.sizes != d2.sizes or d1.dim_names != d2.dim_names:
        return ["dims mismatch: %s vs %s" % (d1.sizes, d2.sizes)]
    if d1.seed != d2.seed:
        return ["seed mismatch"]
    sis = digest_sample_infos(d1.sizes, d1.seed)
    if not (len(sis) == len(d1.samps) == len(d2.samps)):
        return ["sample-count mismatch: sis=%d d1=%d d2=%d" % (len(sis), len(d1.samps), len(d2.samps))]
    bad = []
    for tag, v1, v2 in (("min_v", d1.min_v, d2.min_v), ("max_v", d1.max_v, d2.max_v)):
        if math.isnan(v1) or math.isnan(v2) or _rel_diff(v1, v2) > tol:
            bad.append("[%s]: v1=%r v2=%r" % (tag, v1, v2))
    for (st, off, nss), v1, v2 in zip(sis, d1.samps, d2.samps):
        adj = tol * math.sqrt(nss / 1000.0) if nss > 1000 else tol
        if math.isnan(v1) or math.isnan(v2) or _rel_diff(v1, v2) > adj:
            bad.append("[stride=%d,offset=%d]: v1=%r v2=%r (rd=%.3g adj=%.3g)" % (st, off, v1, v2, _rel_diff(v1, v2), adj))
    return bad
