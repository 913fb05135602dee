"""boda_b200 -- Python (ctypes) face of libboda_b200.so, the B200-native back-end for Boda's rtc_fwd conv/SGEMM path.

The product is the C-ABI shared library (include/boda_b200.h) and the C++ classes behind it (boda_b200/csrc):
`b200_compute_t` mirrors Boda's `rtc_compute_t` (src/rtc_compute.H:35-97) and `b200_conv_fwd_t` mirrors
`has_conv_fwd_t` (src/has_conv_fwd.H:16-25). This module only marshals numpy buffers and strings onto that ABI with
the reference's method names, so tests read like the reference's own flows (ops-prof, test_compute).

There is no CPU fallback: if the library is missing or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libboda_b200.so")
_lib = None


class RtException(RuntimeError):
    """rt_exception (reference: rt_err, src/boda_base.H:98-105)."""


class UnsupException(RtException):
    """unsup_exception (reference: unsup_err): a shape / feature the back-end does not handle."""


_c = ctypes
_strp = _c.POINTER(_c.c_char_p)

_SIGS = {
    "b200_last_error": (_c.c_char_p, []),
    "b200_version": (_c.c_char_p, []),
    "b200_device_count": (_c.c_int, []),
    "b200_rtc_create": (_c.c_void_p, []),
    "b200_rtc_destroy": (None, [_c.c_void_p]),
    "b200_rtc_set_option": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_char_p]),
    "b200_rtc_init": (_c.c_int, [_c.c_void_p]),
    "b200_rtc_get_plat_tag": (_c.c_char_p, [_c.c_void_p]),
    "b200_rtc_create_var": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_char_p, _c.c_int, _strp, _c.POINTER(_c.c_uint32)]),
    "b200_rtc_create_view": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_char_p, _c.c_int, _strp, _c.POINTER(_c.c_uint32), _c.c_char_p]),
    "b200_rtc_release_var": (_c.c_int, [_c.c_void_p, _c.c_char_p]),
    "b200_rtc_get_var_dims": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_int, _c.POINTER(_c.c_uint32), _c.c_char_p, _c.c_int]),
    "b200_rtc_set_var_to_zero": (_c.c_int, [_c.c_void_p, _c.c_char_p]),
    "b200_rtc_compile": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_char_p]),
    "b200_rtc_release_func": (_c.c_int, [_c.c_void_p, _c.c_char_p]),
    "b200_rtc_run": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_int, _strp, _strp]),
    "b200_rtc_finish_and_sync": (_c.c_int, [_c.c_void_p]),
    "b200_rtc_release_per_call_id_data": (_c.c_int, [_c.c_void_p]),
    "b200_rtc_release_all_funcs": (_c.c_int, [_c.c_void_p]),
    "b200_rtc_get_dur": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_uint32, _c.POINTER(_c.c_float)]),
    "b200_rtc_copy_to_var": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.c_uint64]),
    "b200_rtc_copy_from_var": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_char_p, _c.c_uint64]),
    "b200_rtc_get_var_raw_native_pointer": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.POINTER(_c.c_void_p)]),
    "b200_rtc_launches": (_c.c_uint64, [_c.c_void_p]),
    "b200_pipe_describe": (_c.c_int64, [_c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_nda_digest_hex": (_c.c_int64, [_c.c_char_p, _c.c_int, _c.POINTER(_c.c_char_p), _c.POINTER(_c.c_uint32), _c.c_void_p, _c.c_char_p, _c.c_uint64]),
    "b200_wisdom_record": (_c.c_int64, [_c.c_char_p, _c.c_int, _c.POINTER(_c.c_char_p), _c.POINTER(_c.c_char_p), _c.c_char_p, _c.c_char_p, _c.c_double, _c.c_char_p,
                                        _c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_pipe_from_prototxt": (_c.c_int64, [_c.c_char_p, _c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_fwd_plan": (_c.c_int64, [_c.c_char_p, _c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_pipe_op_sigs": (_c.c_int64, [_c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_op_canonical": (_c.c_int64, [_c.c_char_p, _c.c_char_p, _c.c_uint64]),
    "b200_wis_ana": (_c.c_int64, [_c.c_char_p, _c.c_uint32, _c.c_char_p, _c.c_char_p, _c.c_double, _c.c_int, _c.c_char_p, _c.c_uint64]),
    "b200_fwd_create": (_c.c_void_p, [_c.c_char_p, _c.c_char_p]),
    "b200_fwd_destroy": (None, [_c.c_void_p]),
    "b200_fwd_set_param": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.c_uint64]),
    "b200_fwd_run": (_c.c_int, [_c.c_void_p, _c.c_int, _strp, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_uint64), _c.c_int, _strp,
                                _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_uint64)]),
    "b200_fwd_submit": (_c.c_int, [_c.c_void_p, _c.c_int, _strp, _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_uint64), _c.c_int, _strp,
                                   _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_uint64)]),
    "b200_fwd_wait": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "b200_fwd_run_device_only": (_c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(_c.c_float)]),
    "b200_fwd_set_det_drop_seed": (_c.c_int, [_c.c_void_p, _c.c_uint32]),
    "b200_fwd_get_info_log": (_c.c_char_p, [_c.c_void_p]),
    "b200_fwd_get_node_dims": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.POINTER(_c.c_uint32)]),
    "b200_fwd_num_calls": (_c.c_int, [_c.c_void_p]),
    "b200_fwd_launches": (_c.c_uint64, [_c.c_void_p]),
    "b200_fwd_profile": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_char_p, _c.c_int, _c.POINTER(_c.c_float), _c.POINTER(_c.c_float), _c.POINTER(_c.c_double), _c.c_int]),
    "b200_fwd_run_timed": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_uint64, _c.POINTER(_c.c_float)]),
    "b200_fwd_get_stream": (_c.c_int, [_c.c_void_p, _c.POINTER(_c.c_void_p)]),
    "b200_fwd_enqueue": (_c.c_int, [_c.c_void_p]),
    "b200_fwd_flush_l2": (_c.c_int, [_c.c_void_p, _c.c_uint64]),
    "b200_fwd_get_node_raw_native_pointer": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.POINTER(_c.c_void_p)]),
    "b200_rtc_get_kernel_dur": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.POINTER(_c.c_float)]),
    "b200_fwd_set_param_device": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.c_uint64]),
    "b200_shard_create": (_c.c_void_p, [_c.c_int, _c.c_int, _c.c_int]),
    "b200_shard_destroy": (None, [_c.c_void_p]),
    "b200_shard_nccl_unique_id": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "b200_shard_nccl_init": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "b200_shard_broadcast": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_int, _c.c_void_p]),
    "b200_shard_gather_export": (_c.c_int, [_c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "b200_shard_gather_import": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "b200_shard_gather_push": (_c.c_int64, [_c.c_void_p, _c.c_void_p, _c.c_void_p]),
    "b200_shard_gather_wait": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.c_void_p]),
    "b200_shard_gather_push_wait": (_c.c_int64, [_c.c_void_p, _c.c_void_p, _c.c_uint32, _c.c_void_p]),
    "b200_fwd_attach_gather": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_char_p]),
    "b200_shard_step_from_device": (_c.c_int64, [_c.c_void_p]),
    "b200_shard_gather_ptr": (_c.c_int, [_c.c_void_p, _c.c_uint32, _c.POINTER(_c.c_void_p)]),
    "b200_shard_all_gather_nccl": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_uint64, _c.c_void_p]),
    "b200_shard_launches": (_c.c_uint64, [_c.c_void_p]),
}
ABI_SYMBOLS = tuple(_SIGS)


def lib():
    """Load libboda_b200.so (built in-tree by boda_b200/build.py or __graft_entry__.build()). Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RtException("libboda_b200.so is not built (%s); run `python -m boda_b200.build` -- there is no fallback path" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _chk(rc: int) -> int:
    if rc < 0:
        msg = lib().b200_last_error().decode()
        raise (UnsupException if rc == -2 else RtException)(msg)
    return rc


def _b(s: str) -> bytes:
    return s.encode()


def pipe_from_prototxt(prototxt_text: str, in_dims: Optional[Dict[str, int]] = None, out_node_name: str = "", keep_softmax: bool = False) -> str:
    """Host-only (no GPU, no protobuf): Caffe prototxt -> the conv_pipe text B200ConvFwd takes; the TEST-phase translation of
    create_pipe_from_param (src/caffepb.cc:166-326). in_dims overrides source-node dims by name, e.g. {"img": 32}."""
    kv = ["%s=%d" % (k, v) for k, v in (in_dims or {}).items()]
    if out_node_name:
        kv.append("out_node_name=" + out_node_name)
    if keep_softmax:
        kv.append("keep_softmax=1")
    opts = "(" + ",".join(kv) + ")" if kv else ""
    need = _chk(lib().b200_pipe_from_prototxt(_b(prototxt_text), _b(opts), None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_pipe_from_prototxt(_b(prototxt_text), _b(opts), buf, need + 1))
    return buf.value.decode()


def nda_digest_hex(var_name: str, arr: np.ndarray, dim_names: Sequence[str]) -> str:
    """hex(bwrite(nda_digest_t)) of a float tensor, as Boda's ops-prof stores known-good outputs in wisdom files (src/boda_base.cc:210-383)."""
    a = np.ascontiguousarray(arr, np.float32)
    names = _str_array(list(dim_names))
    sizes = (_c.c_uint32 * a.ndim)(*a.shape)
    args = (_b(var_name), a.ndim, names, sizes, a.ctypes.data_as(_c.c_void_p))
    need = _chk(lib().b200_nda_digest_hex(*args, None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_nda_digest_hex(*args, buf, need + 1))
    return buf.value.decode()


def wisdom_record(op_text: str, kgs: Sequence[Tuple[str, str]], op_tune_text: str, be_plat_tag: str, rt_secs: float, err: str = "", run_op_text: str = "") -> str:
    """One op_wisdom_t text record (src/op-tuner.cc:98-130) for Boda's wisdom files / wis-ana."""
    kn, kh = _str_array([k for k, _ in kgs]), _str_array([h for _, h in kgs])
    args = (_b(op_text), len(kgs), kn, kh, _b(op_tune_text), _b(be_plat_tag), float(rt_secs), _b(err), _b(run_op_text or op_text))
    need = _chk(lib().b200_wisdom_record(*args, None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_wisdom_record(*args, buf, need + 1))
    return buf.value.decode()


def pipe_describe(pipe_text: str) -> Dict[str, object]:
    """Host-only (no GPU): parse a conv_pipe text with the C++ graph IR and return {"nodes": {name: ((dim, size), ...)}, "params": [names],
    "ops": n, "conv_flops": f} after the dims inference of conv_pipe_t::calc_dims (src/conv_util.cc:405-529)."""
    need = _chk(lib().b200_pipe_describe(_b(pipe_text), None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_pipe_describe(_b(pipe_text), buf, need + 1))
    nodes, params, res = {}, [], {}
    for line in buf.value.decode().splitlines():
        parts = line.split(" ")
        if parts[0] == "ops":
            res["ops"], res["conv_flops"] = int(parts[1]), int(parts[3])
            continue
        nodes[parts[0]] = tuple((d.split("=")[0], int(d.split("=")[1])) for d in parts[1].split(":"))
        if len(parts) > 2 and parts[2] == "param":
            params.append(parts[0])
    res["nodes"], res["params"] = nodes, params
    return res


def op_canonical(op_text: str) -> str:
    """Host-only: an op line in the current or the stale op-list syntax -> the canonical current-syntax line (the reference's printer)."""
    need = _chk(lib().b200_op_canonical(_b(op_text), None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_op_canonical(_b(op_text), buf, need + 1))
    return buf.value.decode()


def pipe_op_sigs(pipe_text: str) -> List[str]:
    """Host-only: the unique Convolution signatures of a pipe as canonical op lines (the reference's write_op_sigs, src/rtc_fwd.cc:246-264)."""
    need = _chk(lib().b200_pipe_op_sigs(_b(pipe_text), None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_pipe_op_sigs(_b(pipe_text), buf, need + 1))
    return buf.value.decode().splitlines()


def fwd_plan(pipe_text: str, opts: str = "") -> Dict[str, object]:
    """Host-only (no GPU): the forward as B200ConvFwd would plan it for this pipe and these options -- graph passes and per-layer launch
    plans for a 148-SM device (b200_fwd_plan). Returns {"calls": [(func_name, {arg: var-or-scalar})], "prep": [...], "alias": {node: (concat
    node, chan offset)}, "join": {conv tag: (join node, residual node)}, "absmax": {node: cell}}. A plan cannot compute anything."""
    need = _chk(lib().b200_fwd_plan(_b(pipe_text), _b(opts), None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_fwd_plan(_b(pipe_text), _b(opts), buf, need + 1))
    res: Dict[str, object] = {"calls": [], "prep": [], "alias": {}, "join": {}, "absmax": {}, "lrnpool": {}, "fcchain": []}
    for line in buf.value.decode().splitlines():
        parts = line.split(" ")
        if parts[0] in ("call", "prep"):
            res["calls" if parts[0] == "call" else "prep"].append((parts[1], dict(p.split("=", 1) for p in parts[2:])))  # incl. "plan:<key>" items
        elif parts[0] == "alias":
            res["alias"][parts[1]] = (parts[2], int(parts[3]))
        elif parts[0] == "join":
            res["join"][parts[1]] = (parts[2], parts[3])
        elif parts[0] == "absmax":
            res["absmax"][parts[1]] = int(parts[2])
        elif parts[0] == "fcchain":  # the convolutions one fc_chain call runs, first to last
            res["fcchain"].append(parts[1:])
        elif parts[0] == "lrnpool":  # pool tag -> (the LRN that runs inside its kernel, the node that kernel reads)
            res["lrnpool"][parts[1]] = (parts[2], parts[3])
    return res


def wis_ana(wisdom_text: str, s_img: int = 0, s_plat: str = ".*", ref_tune: str = "", min_flops: float = 0.0, csv: bool = False):
    """The reference's wis-ana analysis (src/op-tuner.cc:204-396) over wisdom text in its format. csv=True: the text wis-plot.py reads.
    Otherwise {"aom_tune": tune, "tot_runs": n, "rows": [{"op", "flops", "aom", "pom", "ref", "pom_tune"}]} (seconds; NaN = no such run)."""
    args = (_b(wisdom_text), s_img, _b(s_plat), _b(ref_tune), float(min_flops), 0 if csv else 1)
    need = _chk(lib().b200_wis_ana(*args, None, 0))
    buf = ctypes.create_string_buffer(need + 1)
    _chk(lib().b200_wis_ana(*args, buf, need + 1))
    text = buf.value.decode()
    if csv:
        return text
    lines = text.splitlines()
    head = lines[0].split("\t")
    res = {"aom_tune": head[1], "tot_runs": int(head[3]), "rows": []}
    for l in lines[1:]:
        fl, aom, pom, ref, tune, op = l.split("\t")
        res["rows"].append({"op": op, "flops": int(fl), "aom": float(aom), "pom": float(pom), "ref": float(ref), "pom_tune": tune})
    return res


def _str_array(items: Sequence[str]):
    arr = (_c.c_char_p * len(items))(*[_b(i) for i in items])
    return arr


def device_count() -> int:
    return int(lib().b200_device_count())


Dims = Sequence[Tuple[str, int]]


class B200Compute:
    """`be=b200`: same methods as rtc_compute_t (src/rtc_compute.H:35-97); ndas are numpy arrays + named dims."""

    def __init__(self, prec: str = "fp32", acc_chunk_kblks: Optional[int] = None, device: int = 0, use_clusters: Optional[int] = None, use_2cta: Optional[int] = None,
                 **opts):
        self._h = lib().b200_rtc_create()
        if not self._h:
            raise RtException(lib().b200_last_error().decode())
        _chk(lib().b200_rtc_set_option(self._h, b"prec", _b(prec)))
        _chk(lib().b200_rtc_set_option(self._h, b"device", _b(str(device))))
        if acc_chunk_kblks is not None:
            _chk(lib().b200_rtc_set_option(self._h, b"acc_chunk_kblks", _b(str(acc_chunk_kblks))))
        if use_clusters is not None:
            _chk(lib().b200_rtc_set_option(self._h, b"use_clusters", _b(str(use_clusters))))
        if use_2cta is not None:
            _chk(lib().b200_rtc_set_option(self._h, b"use_2cta", _b(str(use_2cta))))
        for k, v in opts.items():  # any other back-end option (use_taps, taps_2cta, acc_chunk_kblks_16, debug_flags ...); unknown keys are errors
            _chk(lib().b200_rtc_set_option(self._h, _b(k), _b(str(v))))
        self._dims: Dict[str, Dims] = {}

    def close(self):
        if self._h:
            lib().b200_rtc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        _chk(lib().b200_rtc_init(self._h))

    def get_plat_tag(self) -> str:
        return lib().b200_rtc_get_plat_tag(self._h).decode()

    def create_var_with_dims(self, vn: str, dims: Dims, tn: str = "float"):
        names = _str_array([d[0] for d in dims])
        sizes = (_c.c_uint32 * len(dims))(*[int(d[1]) for d in dims])
        _chk(lib().b200_rtc_create_var(self._h, _b(vn), _b(tn), len(dims), names, sizes))
        self._dims[vn] = tuple(dims)

    def create_var_with_dims_as_reshaped_view_of_var(self, vn: str, dims: Dims, src_vn: str, tn: str = "float"):
        names = _str_array([d[0] for d in dims])
        sizes = (_c.c_uint32 * len(dims))(*[int(d[1]) for d in dims])
        _chk(lib().b200_rtc_create_view(self._h, _b(vn), _b(tn), len(dims), names, sizes, _b(src_vn)))
        self._dims[vn] = tuple(dims)

    def release_var(self, vn: str):
        _chk(lib().b200_rtc_release_var(self._h, _b(vn)))
        self._dims.pop(vn, None)

    def get_var_dims(self, vn: str) -> List[Tuple[str, int]]:
        sizes = (_c.c_uint32 * 8)()
        buf = ctypes.create_string_buffer(256)
        n = _chk(lib().b200_rtc_get_var_dims(self._h, _b(vn), 8, sizes, buf, 256))
        names = buf.value.decode().split(":") if n else []
        return [(names[i], int(sizes[i])) for i in range(n)]

    def set_var_to_zero(self, vn: str):
        _chk(lib().b200_rtc_set_var_to_zero(self._h, _b(vn)))

    def compile(self, func_name: str, op_text: str):
        """rtc_compute_t::compile for one rtc_func_info_t{func_name, op}."""
        _chk(lib().b200_rtc_compile(self._h, _b(func_name), _b(op_text)))

    def release_func(self, func_name: str):
        _chk(lib().b200_rtc_release_func(self._h, _b(func_name)))

    def release_all_funcs(self):
        _chk(lib().b200_rtc_release_all_funcs(self._h))

    def run(self, func_name: str, arg_map: Dict[str, object]) -> int:
        """rtc_compute_t::run(rtc_func_call_t{func_name, arg_map}) -> call_id. Values: var name (str) or a scalar
        (int -> uint32_t nda, float -> float nda) passed by value."""
        names, vals = [], []
        for k, v in arg_map.items():
            names.append(k)
            if isinstance(v, str):
                vals.append(v)
            elif isinstance(v, (int, np.integer)):
                vals.append("(tn=uint32_t,v=%d)" % int(v))
            else:
                vals.append("(tn=float,v=%r)" % float(v))
        return _chk(lib().b200_rtc_run(self._h, _b(func_name), len(names), _str_array(names), _str_array(vals)))

    def finish_and_sync(self):
        _chk(lib().b200_rtc_finish_and_sync(self._h))

    def release_per_call_id_data(self):
        _chk(lib().b200_rtc_release_per_call_id_data(self._h))

    def get_kernel_dur(self, call_id: int) -> float:
        ms = _c.c_float()
        _chk(lib().b200_rtc_get_kernel_dur(self._h, call_id, ctypes.byref(ms)))
        return float(ms.value)

    def get_dur(self, b: int, e: int) -> float:
        ms = _c.c_float()
        _chk(lib().b200_rtc_get_dur(self._h, b, e, ctypes.byref(ms)))
        return float(ms.value)

    def copy_nda_to_var(self, vn: str, nda: np.ndarray):
        a = np.ascontiguousarray(nda)
        _chk(lib().b200_rtc_copy_to_var(self._h, _b(vn), a.ctypes.data_as(_c.c_void_p), a.nbytes))

    def copy_var_to_nda(self, vn: str, dtype=np.float32) -> np.ndarray:
        dims = self.get_var_dims(vn)
        out = np.empty([d[1] for d in dims], dtype)
        _chk(lib().b200_rtc_copy_from_var(self._h, out.ctypes.data_as(_c.c_void_p), _b(vn), out.nbytes))
        return out

    def create_var_from_nda(self, vn: str, nda: np.ndarray, dim_names: Sequence[str]):
        self.create_var_with_dims(vn, list(zip(dim_names, nda.shape)))
        self.copy_nda_to_var(vn, nda)

    def get_var_raw_native_pointer(self, vn: str) -> int:
        p = _c.c_void_p()
        _chk(lib().b200_rtc_get_var_raw_native_pointer(self._h, _b(vn), ctypes.byref(p)))
        return int(p.value or 0)

    def launches(self) -> int:
        return int(lib().b200_rtc_launches(self._h))


class B200ConvFwd:
    """`mode=b200`: has_conv_fwd_t (src/has_conv_fwd.H:16-25). `pipe_text` is the conv_pipe in op-line text form."""

    def __init__(self, pipe_text: str, opts: str = ""):
        self._h = lib().b200_fwd_create(_b(pipe_text), _b(opts))
        if not self._h:
            msg = lib().b200_last_error().decode()
            raise (UnsupException if msg.startswith("unsupported") else RtException)(msg)

    def close(self):
        if self._h:
            lib().b200_fwd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def node_dims(self, name: str) -> Tuple[int, ...]:
        d = (_c.c_uint32 * 4)()
        n = _chk(lib().b200_fwd_get_node_dims(self._h, _b(name), d))
        return tuple(int(d[i]) for i in range(n))

    def set_param(self, name: str, arr: np.ndarray):
        a = np.ascontiguousarray(arr, np.float32)
        _chk(lib().b200_fwd_set_param(self._h, _b(name), a.ctypes.data_as(_c.c_void_p), a.size))

    def set_param_device(self, name: str, dev_ptr: int, n_elems: int):
        """Parameter upload from a device buffer (a slice of the flat buffer the weights were broadcast in): no host round trip."""
        _chk(lib().b200_fwd_set_param_device(self._h, _b(name), _c.c_void_p(dev_ptr), n_elems))

    def attach_gather(self, shard, node: str) -> bool:
        """Multi-GPU: let the forward's last kernel (an fc_chain ending in `node`) gather the logits into every rank's buffer itself
        (b200_fwd_attach_gather). False: this net does not end that way -- keep calling B200Shard.gather_push_wait. shard=None detaches."""
        return bool(_chk(lib().b200_fwd_attach_gather(self._h, shard._h if shard is not None else None, _b(node))))

    def run_fwd(self, to_set: Dict[str, np.ndarray], to_get: Sequence[str], out_bufs: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """has_conv_fwd_t::run_fwd(to_set_vns, fwd, to_get_vns): host fp32 NCHW in, host out, synchronous."""
        sn = list(to_set)
        sa = [np.ascontiguousarray(to_set[k], np.float32) for k in sn]
        outs = {}
        for k in to_get:
            outs[k] = out_bufs[k] if out_bufs and k in out_bufs else np.empty(self.node_dims(k), np.float32)
        gn = list(to_get)
        sp = (_c.c_void_p * len(sn))(*[a.ctypes.data for a in sa])
        se = (_c.c_uint64 * len(sn))(*[a.size for a in sa])
        gp = (_c.c_void_p * len(gn))(*[outs[k].ctypes.data for k in gn])
        ge = (_c.c_uint64 * len(gn))(*[outs[k].size for k in gn])
        _chk(lib().b200_fwd_run(self._h, len(sn), _str_array(sn), sp, se, len(gn), _str_array(gn), gp, ge))
        return outs

    def run_fwd_ptrs(self, set_names, set_ptrs, set_elems, get_names, get_ptrs, get_elems):
        """Same call with raw host pointers (e.g. pinned torch tensors' data_ptr()) -- used by bench.py's e2e leg."""
        sp = (_c.c_void_p * len(set_names))(*set_ptrs)
        se = (_c.c_uint64 * len(set_names))(*set_elems)
        gp = (_c.c_void_p * len(get_names))(*get_ptrs)
        ge = (_c.c_uint64 * len(get_names))(*get_elems)
        _chk(lib().b200_fwd_run(self._h, len(set_names), _str_array(set_names), sp, se, len(get_names), _str_array(get_names), gp, ge))

    def submit_ptrs(self, set_names, set_ptrs, set_elems, get_names, get_ptrs, get_elems) -> int:
        """Pipelined run_fwd: enqueue H2D (overlapping the previous forward) + forward + D2H, return a ticket for wait()."""
        sp = (_c.c_void_p * len(set_names))(*set_ptrs)
        se = (_c.c_uint64 * len(set_names))(*set_elems)
        gp = (_c.c_void_p * len(get_names))(*get_ptrs)
        ge = (_c.c_uint64 * len(get_names))(*get_elems)
        return _chk(lib().b200_fwd_submit(self._h, len(set_names), _str_array(set_names), sp, se, len(get_names), _str_array(get_names), gp, ge))

    def wait(self, ticket: int):
        _chk(lib().b200_fwd_wait(self._h, ticket))

    def run_device_only(self, iters: int) -> float:
        ms = _c.c_float()
        _chk(lib().b200_fwd_run_device_only(self._h, iters, ctypes.byref(ms)))
        return float(ms.value)

    def set_det_drop_seed(self, seed: int):
        _chk(lib().b200_fwd_set_det_drop_seed(self._h, seed))

    def get_info_log(self) -> str:
        return lib().b200_fwd_get_info_log(self._h).decode()

    def num_calls(self) -> int:
        return int(lib().b200_fwd_num_calls(self._h))

    def launches(self) -> int:
        return int(lib().b200_fwd_launches(self._h))

    def profile(self, iters: int = 5) -> List[Tuple[str, float, float, float]]:
        """[(func_name, call_ms, contraction_kernel_ms, algorithmic_flops)] per forward call, averaged over `iters` eager runs."""
        n = self.num_calls()
        buf = ctypes.create_string_buffer(1 << 16)
        ms = (_c.c_float * max(n, 1))()
        kms = (_c.c_float * max(n, 1))()
        fl = (_c.c_double * max(n, 1))()
        got = _chk(lib().b200_fwd_profile(self._h, iters, buf, len(buf), ms, kms, fl, n))
        tags = buf.value.decode().split("\n") if got else []
        return [(tags[i], float(ms[i]), float(kms[i]), float(fl[i])) for i in range(got)]

    def run_timed(self, iters: int, l2_flush_bytes: int = 0) -> List[float]:
        ms = (_c.c_float * max(iters, 1))()
        _chk(lib().b200_fwd_run_timed(self._h, iters, l2_flush_bytes, ms))
        return [float(ms[i]) for i in range(iters)]

    def stream_ptr(self) -> int:
        """The back-end's cudaStream_t, e.g. for torch.cuda.ExternalStream."""
        p = _c.c_void_p()
        _chk(lib().b200_fwd_get_stream(self._h, ctypes.byref(p)))
        return int(p.value or 0)

    def enqueue(self):
        """Queue one forward on device-resident inputs; returns without synchronising."""
        _chk(lib().b200_fwd_enqueue(self._h))

    def flush_l2(self, nbytes: int):
        _chk(lib().b200_fwd_flush_l2(self._h, nbytes))

    def node_device_ptr(self, name: str) -> int:
        p = _c.c_void_p()
        _chk(lib().b200_fwd_get_node_raw_native_pointer(self._h, _b(name), ctypes.byref(p)))
        return int(p.value or 0)


class B200Shard:
    """One process's end of the batch-sharded forward (include/boda_b200.h, `b200_shard_*`): the NCCL communicator for the
    one weight broadcast and the peer-mapped gather buffers for the per-step logits gather. `exchange(obj) -> [obj of rank 0, 1, ...]` is the
    caller's transport for the rendezvous bytes (e.g. torch.distributed.all_gather_object)."""

    def __init__(self, device: int, rank: int, world: int):
        self.rank, self.world = rank, world
        self._h = lib().b200_shard_create(device, rank, world)
        if not self._h:
            raise RtException(lib().b200_last_error().decode())

    def close(self):
        if self._h:
            lib().b200_shard_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def nccl_init(self, exchange):
        idb = (_c.c_ubyte * 128)()
        if self.rank == 0:
            _chk(lib().b200_shard_nccl_unique_id(self._h, idb))
        got = exchange(bytes(idb))[0]
        buf = (_c.c_ubyte * 128).from_buffer_copy(got)
        _chk(lib().b200_shard_nccl_init(self._h, buf))

    def broadcast(self, dev_ptr: int, nbytes: int, root: int = 0, stream: int = 0):
        _chk(lib().b200_shard_broadcast(self._h, _c.c_void_p(dev_ptr), nbytes, root, _c.c_void_p(stream)))

    def gather_setup(self, bytes_per_rank: int, exchange):
        """Allocate the local gather buffer, swap IPC handles, map every peer's buffer."""
        h = (_c.c_ubyte * 64)()
        _chk(lib().b200_shard_gather_export(self._h, bytes_per_rank, h))
        allh = exchange(bytes(h))
        flat = (_c.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(allh))
        _chk(lib().b200_shard_gather_import(self._h, flat))

    def gather_ptr(self, step: int) -> int:
        """Device pointer of the local [world][bytes_per_rank] result of `step` (valid once gather_wait(step) has run on the stream)."""
        p = _c.c_void_p()
        _chk(lib().b200_shard_gather_ptr(self._h, step, _c.byref(p)))
        return int(p.value)

    def gather_push(self, dev_src: int, stream: int) -> int:
        return _chk(lib().b200_shard_gather_push(self._h, _c.c_void_p(dev_src), _c.c_void_p(stream)))

    def step_from_device(self) -> int:
        """The step the last fused forward (B200ConvFwd.attach_gather) published; re-synchronises the host-side counter."""
        return _chk(lib().b200_shard_step_from_device(self._h))

    def gather_push_wait(self, dev_src: int, wait_step: int, stream: int) -> int:
        """Push this step's logits and wait for `wait_step` (0 = none) in ONE kernel launch; returns the step it published."""
        return _chk(lib().b200_shard_gather_push_wait(self._h, _c.c_void_p(dev_src), int(wait_step), _c.c_void_p(stream)))

    def gather_wait(self, step: int, stream: int):
        _chk(lib().b200_shard_gather_wait(self._h, step, _c.c_void_p(stream)))

    def all_gather_nccl(self, dev_src: int, dev_dst: int, bytes_per_rank: int, stream: int):
        _chk(lib().b200_shard_all_gather_nccl(self._h, _c.c_void_p(dev_src), _c.c_void_p(dev_dst), bytes_per_rank, _c.c_void_p(stream)))

    def launches(self) -> int:
        return int(lib().b200_shard_launches(self._h))
