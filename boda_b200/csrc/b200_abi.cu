// b200_abi.cu -- extern "C" surface declared in include/boda_b200.h. No exceptions cross this boundary.
#include "../../include/boda_b200.h"
#include "b200_conv_fwd.h"
#include "caffe_prototxt.h"
#include "wisdom.h"
#include "b200_shard.h"
#include <cstdio>

using namespace boda;

struct b200_rtc { p_b200_compute_t rtc; string tmp; };
struct b200_fwd { shared_ptr<b200_conv_fwd_t> fwd; string tmp; };
struct b200_shard { shared_ptr<b200_shard_t> sh; };

namespace {
thread_local string g_last_error;

template <typename F> int guarded(F &&f) {
  try { return f(); }
  catch (unsup_exception const &e) { g_last_error = e.what(); return -2; }
  catch (std::exception const &e) { g_last_error = e.what(); return -1; }
  catch (...) { g_last_error = "unknown exception"; return -1; }
}

// entries that take an instance first make its device current for the calling thread (another instance, or the caller's own framework, may
// have switched it since): allocations, launches, copies and graph replays below all use the current device
template <typename F> int guarded_dev(b200_rtc *r, F &&f) { return guarded([&] { r->rtc->bind_device(); return f(); }); }
template <typename F> int guarded_dev(b200_fwd *f_, F &&f) { return guarded([&] { if (f_->fwd && f_->fwd->rtc) { f_->fwd->rtc->bind_device(); } return f(); }); }

dims_t make_dims(char const *tn, int ndims, char const *const *dim_names, uint32_t const *dim_sizes) {
  dims_t d;
  d.tn = tn ? tn : "float";
  for (int i = 0; i < ndims; ++i) { d.add_dim(dim_names[i], dim_sizes[i]); }
  d.calc_strides();
  return d;
}
}  // namespace

extern "C" {

#define B200_API __attribute__((visibility("default")))

B200_API const char *b200_last_error(void) { return g_last_error.c_str(); }
B200_API const char *b200_version(void) { return "boda_b200 0.1 (sm_100a; tcgen05+TMA igemm, fp16x2-split fp32 parity mode)"; }
B200_API int b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

B200_API b200_rtc *b200_rtc_create(void) {
  b200_rtc *r = nullptr;
  guarded([&] { r = new b200_rtc; r->rtc = std::make_shared<b200_compute_t>(); return 0; });
  return r;
}
B200_API void b200_rtc_destroy(b200_rtc *r) { guarded([&] { delete r; return 0; }); }
B200_API int b200_rtc_set_option(b200_rtc *r, const char *key, const char *val) {
  return guarded([&] {
    string const k = key, v = val;
    if (r->rtc->set_option(k, v)) {}
    else { rt_err("be=b200: unused option '" + k + "'"); }
    return 0;
  });
}
B200_API int b200_rtc_init(b200_rtc *r) { return guarded_dev(r, [&] { r->rtc->init(); return 0; }); }
B200_API const char *b200_rtc_get_plat_tag(b200_rtc *r) { r->tmp = r->rtc->get_plat_tag(); return r->tmp.c_str(); }
B200_API int b200_rtc_create_var(b200_rtc *r, const char *vn, const char *tn, int ndims, const char *const *dim_names, const uint32_t *dim_sizes) {
  return guarded_dev(r, [&] { r->rtc->create_var_with_dims(vn, make_dims(tn, ndims, dim_names, dim_sizes)); return 0; });
}
B200_API int b200_rtc_create_view(b200_rtc *r, const char *vn, const char *tn, int ndims, const char *const *dim_names, const uint32_t *dim_sizes, const char *src_vn) {
  return guarded_dev(r, [&] { r->rtc->create_var_with_dims_as_reshaped_view_of_var(vn, make_dims(tn, ndims, dim_names, dim_sizes), src_vn); return 0; });
}
B200_API int b200_rtc_release_var(b200_rtc *r, const char *vn) { return guarded_dev(r, [&] { r->rtc->release_var(vn); return 0; }); }
B200_API int b200_rtc_get_var_dims(b200_rtc *r, const char *vn, int max_dims, uint32_t *dim_sizes, char *names_buf, int names_buf_len) {
  return guarded_dev(r, [&] {
    dims_t const d = r->rtc->get_var_dims(vn);
    string names;
    for (size_t i = 0; i < d.size(); ++i) { if ((int)i < max_dims) { dim_sizes[i] = d[i].sz; } names += (i ? ":" : "") + d[i].name; }
    if (names_buf && names_buf_len > 0) { snprintf(names_buf, names_buf_len, "%s", names.c_str()); }
    return (int)d.size();
  });
}
B200_API int b200_rtc_set_var_to_zero(b200_rtc *r, const char *vn) { return guarded_dev(r, [&] { r->rtc->set_var_to_zero(vn); return 0; }); }
B200_API int b200_rtc_compile(b200_rtc *r, const char *func_name, const char *op_text) {
  return guarded_dev(r, [&] {
    rtc_func_info_t fi;
    fi.func_name = func_name;
    fi.op = *make_p_op_base_t_from_str(op_text);
    r->rtc->compile({fi}, rtc_compile_opts_t());
    return 0;
  });
}
B200_API int b200_rtc_release_func(b200_rtc *r, const char *func_name) { return guarded_dev(r, [&] { r->rtc->release_func(func_name); return 0; }); }
B200_API int b200_rtc_run(b200_rtc *r, const char *func_name, int nargs, const char *const *arg_names, const char *const *arg_vals) {
  return guarded_dev(r, [&] {
    rtc_func_call_t rfc;
    rfc.rtc_func_name = func_name;
    for (int i = 0; i < nargs; ++i) {
      string const v = arg_vals[i];
      if (!v.empty() && v[0] == '(') { rfc.arg_map[arg_names[i]] = rtc_arg_t(nda_from_lexp(*parse_lexp(v))); }
      else { rfc.arg_map[arg_names[i]] = rtc_arg_t(v); }
    }
    return (int)r->rtc->run(rfc);
  });
}
B200_API int b200_rtc_finish_and_sync(b200_rtc *r) { return guarded_dev(r, [&] { r->rtc->finish_and_sync(); return 0; }); }
B200_API int b200_rtc_release_per_call_id_data(b200_rtc *r) { return guarded_dev(r, [&] { r->rtc->release_per_call_id_data(); return 0; }); }
B200_API int b200_rtc_release_all_funcs(b200_rtc *r) { return guarded_dev(r, [&] { r->rtc->release_all_funcs(); return 0; }); }
B200_API int b200_rtc_get_dur(b200_rtc *r, uint32_t b, uint32_t e, float *ms_out) { return guarded_dev(r, [&] { *ms_out = r->rtc->get_dur(b, e); return 0; }); }
B200_API int b200_rtc_copy_to_var(b200_rtc *r, const char *vn, const void *host_src, uint64_t bytes) { return guarded_dev(r, [&] { r->rtc->copy_raw_to_var(vn, host_src, bytes); return 0; }); }
B200_API int b200_rtc_copy_from_var(b200_rtc *r, void *host_dst, const char *vn, uint64_t bytes) { return guarded_dev(r, [&] { r->rtc->copy_var_to_raw(host_dst, vn, bytes); return 0; }); }
B200_API int b200_rtc_get_var_raw_native_pointer(b200_rtc *r, const char *vn, void **dev_ptr_out) {
  return guarded_dev(r, [&] { *dev_ptr_out = r->rtc->get_var_raw_native_pointer(vn)->rp_elems(); return 0; });
}
B200_API uint64_t b200_rtc_launches(b200_rtc *r) { return r->rtc->launches(); }

// ---- host-only pipe description ----
B200_API int64_t b200_pipe_describe(const char *pipe_text, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    p_conv_pipe_t cp = make_conv_pipe_from_text(pipe_text);
    string out;
    for (auto const &kv : cp->nodes) {
      out += kv.first + " ";
      dims_t const &d = kv.second->dims;
      for (size_t i = 0; i < d.size(); ++i) { out += (i ? ":" : "") + d[i].name + "=" + str(d[i].sz); }
      if (kv.second->is_param) { out += " param"; }
      out += "\n";
    }
    out += "ops " + str(cp->ops.size()) + " conv_flops " + str(cp->total_conv_flops()) + "\n";
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_op_canonical(const char *op_text, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    string const out = op_base_text(*make_p_op_base_t_from_str(op_text));
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_pipe_op_sigs(const char *pipe_text, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    string const out = conv_pipe_op_sigs_text(*make_conv_pipe_from_text(pipe_text));
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_fwd_plan(const char *pipe_text, const char *opts, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    string o = (opts && opts[0]) ? string(opts) : string("()");
    size_t const close = o.rfind(')');
    if (o.empty() || o[0] != '(' || close == string::npos) { rt_err("b200_fwd_plan: opts must be a (key=value,...) list"); }
    o = o.substr(0, close) + (close > 1 ? "," : "") + "plan_only=1)";
    b200_conv_fwd_t fwd;
    fwd.init(make_conv_pipe_from_text(pipe_text), o);
    string const out = fwd.plan_text();
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_pipe_from_prototxt(const char *prototxt_text, const char *opts, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    prototxt_opts_t po;
    if (opts && *opts) {
      p_lexp_t l = parse_lexp(opts);
      if (!l->is_leaf) {
        for (auto const &kv : l->kids) {
          string const &k = kv.first, &v = kv.second->leaf;
          if (k == "out_node_name") { po.out_node_name = v; }
          else if (k == "keep_softmax") { po.keep_softmax = std::stoi(v) != 0; }
          else { po.in_dims[k] = (uint32_t)std::stoul(v); }
        }
      }
    }
    string const out = conv_pipe_text_from_prototxt(prototxt_text, po);
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_nda_digest_hex(const char *var_name, int ndims, const char *const *dim_names, const uint32_t *dim_sizes, const float *host_data,
                                     char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    string const out = nda_digest_hex(var_name, make_dims("float", ndims, dim_names, dim_sizes), host_data);
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}
B200_API int64_t b200_wisdom_record(const char *op_text, int n_kgs, const char *const *kg_names, const char *const *kg_hex, const char *op_tune_text,
                                    const char *be_plat_tag, double rt_secs, const char *err, const char *run_op_text, char *buf, uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    vector<std::pair<string, string>> kgs;
    for (int i = 0; i < n_kgs; ++i) { kgs.push_back({kg_names[i], kg_hex[i]}); }
    wisdom_run_t r;
    r.op_tune_text = op_tune_text; r.be_plat_tag = be_plat_tag; r.rt_secs = rt_secs; r.err = err ? err : ""; r.run_op_text = run_op_text ? run_op_text : "";
    string const out = wisdom_record_text(op_text, kgs, {r});
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

B200_API int64_t b200_wis_ana(const char *wisdom_text, uint32_t s_img, const char *s_plat, const char *ref_tune, double min_flops, int detail, char *buf,
                              uint64_t buf_len) {
  int64_t need = -1;
  int const rc = guarded([&] {
    wis_ana_opts_t o;
    o.s_img = s_img; o.min_flops = min_flops;
    if (s_plat && s_plat[0]) { o.s_plat = s_plat; }
    if (ref_tune) { o.ref_tune = ref_tune; }
    wis_ana_res_t const res = wis_ana(read_wisdom_text(wisdom_text), o);
    string out;
    if (!detail) { out = wis_ana_csv(res, o); }
    else {
      auto num = [](double v) { char b[64]; snprintf(b, sizeof(b), "%.9g", v); return string(b); };
      out = "#aom_tune\t" + res.aom_tune + "\ttot_runs\t" + str(res.tot_runs) + "\n";
      for (auto const &r : res.rows) { out += str(r.flops) + "\t" + num(r.aom) + "\t" + num(r.pom) + "\t" + num(r.ref) + "\t" + r.pom_tune + "\t" + r.op_text + "\n"; }
    }
    need = (int64_t)out.size();
    if (buf && buf_len) { snprintf(buf, buf_len, "%s", out.c_str()); }
    return 0;
  });
  return rc == 0 ? need : rc;
}

// ---- tier B ----
B200_API b200_fwd *b200_fwd_create(const char *pipe_text, const char *opts) {
  b200_fwd *f = nullptr;
  int const rc = guarded([&] {
    f = new b200_fwd;
    f->fwd = std::make_shared<b200_conv_fwd_t>();
    f->fwd->init(make_conv_pipe_from_text(pipe_text), opts ? opts : "");
    return 0;
  });
  if (rc != 0) { delete f; return nullptr; }
  return f;
}
B200_API void b200_fwd_destroy(b200_fwd *f) { guarded([&] { delete f; return 0; }); }
B200_API int b200_fwd_set_param(b200_fwd *f, const char *node_name, const float *host_src, uint64_t n_elems) {
  return guarded_dev(f, [&] { f->fwd->set_param(node_name, host_src, n_elems); return 0; });
}
B200_API int b200_fwd_run(b200_fwd *f, int n_set, const char *const *set_names, const float *const *set_bufs, const uint64_t *set_elems, int n_get,
                          const char *const *get_names, float *const *get_bufs, const uint64_t *get_elems) {
  return guarded_dev(f, [&] { f->fwd->run_fwd_raw(n_set, set_names, set_bufs, set_elems, n_get, get_names, get_bufs, get_elems); return 0; });
}
B200_API int b200_fwd_submit(b200_fwd *f, int n_set, const char *const *set_names, const float *const *set_bufs, const uint64_t *set_elems, int n_get,
                             const char *const *get_names, float *const *get_bufs, const uint64_t *get_elems) {
  return guarded_dev(f, [&] { return f->fwd->submit(n_set, set_names, set_bufs, set_elems, n_get, get_names, get_bufs, get_elems); });
}
B200_API int b200_fwd_wait(b200_fwd *f, int ticket) { return guarded_dev(f, [&] { f->fwd->wait(ticket); return 0; }); }
B200_API int b200_fwd_run_device_only(b200_fwd *f, int iters, float *ms_per_iter_out) {
  return guarded_dev(f, [&] { *ms_per_iter_out = f->fwd->run_device_only(iters); return 0; });
}
B200_API int b200_fwd_set_det_drop_seed(b200_fwd *f, uint32_t seed) { return guarded_dev(f, [&] { f->fwd->set_det_drop_seed(seed); return 0; }); }
B200_API const char *b200_fwd_get_info_log(b200_fwd *f) { f->tmp = f->fwd->get_info_log(); return f->tmp.c_str(); }
B200_API int b200_fwd_get_node_dims(b200_fwd *f, const char *node_name, uint32_t *dims4) {
  return guarded_dev(f, [&] {
    dims_t const &d = f->fwd->cp->must_get_node(node_name)->dims;
    for (size_t i = 0; i < d.size() && i < 4; ++i) { dims4[i] = d[i].sz; }
    return (int)d.size();
  });
}
B200_API int b200_fwd_num_calls(b200_fwd *f) { return (int)f->fwd->fwd_calls.size(); }
B200_API uint64_t b200_fwd_launches(b200_fwd *f) { return f->fwd->launches(); }
B200_API int b200_fwd_profile(b200_fwd *f, int iters, char *tags_buf, int tags_buf_len, float *call_ms_out, float *kernel_ms_out, double *flops_out, int max_calls) {
  return guarded_dev(f, [&] {
    auto res = f->fwd->profile(iters);
    string tags;
    int n = 0;
    for (auto const &r : res) {
      if (n >= max_calls) { break; }
      tags += (n ? "\n" : "") + r.func_name;
      call_ms_out[n] = r.call_ms; kernel_ms_out[n] = r.kernel_ms; flops_out[n] = r.flops;
      ++n;
    }
    if (tags_buf && tags_buf_len > 0) { snprintf(tags_buf, tags_buf_len, "%s", tags.c_str()); }
    return n;
  });
}
B200_API int b200_fwd_run_timed(b200_fwd *f, int iters, uint64_t l2_flush_bytes, float *ms_each_out) {
  return guarded_dev(f, [&] {
    vector<float> ms = f->fwd->run_timed(iters, l2_flush_bytes);
    for (int i = 0; i < iters; ++i) { ms_each_out[i] = ms[i]; }
    return 0;
  });
}
B200_API int b200_fwd_attach_gather(b200_fwd *f, b200_shard *s, const char *node_name) {
  return guarded_dev(f, [&] {
    if (!s) { f->fwd->attach_gather(string(), nullptr); return 0; }
    b200_gather_desc_t d;
    s->sh->gather_desc(d);
    return f->fwd->attach_gather(node_name, &d) ? 1 : 0;
  });
}
B200_API int64_t b200_shard_step_from_device(b200_shard *s) {
  int64_t step = 0;
  int const rc = guarded([&] { step = s->sh->step_from_device(); return 0; });
  return rc < 0 ? rc : step;
}
B200_API int b200_fwd_enqueue(b200_fwd *f) { return guarded_dev(f, [&] { f->fwd->enqueue_fwd(); return 0; }); }
B200_API int b200_fwd_flush_l2(b200_fwd *f, uint64_t bytes) { return guarded_dev(f, [&] { f->fwd->flush_l2(bytes); return 0; }); }
B200_API int b200_fwd_get_stream(b200_fwd *f, void **stream_out) { return guarded_dev(f, [&] { *stream_out = f->fwd->stream(); return 0; }); }
B200_API int b200_fwd_get_node_raw_native_pointer(b200_fwd *f, const char *node_name, void **dev_ptr_out) {
  return guarded_dev(f, [&] { *dev_ptr_out = f->fwd->rtc->get_var_raw_native_pointer(node_name)->rp_elems(); return 0; });
}
B200_API int b200_rtc_get_kernel_dur(b200_rtc *r, uint32_t id, float *ms_out) { return guarded_dev(r, [&] { *ms_out = r->rtc->get_kernel_dur(id); return 0; }); }


/* ---- batch sharding over the GPUs of one box ---- */
B200_API b200_shard *b200_shard_create(int device, int rank, int world) {
  b200_shard *s = nullptr;
  guarded([&] { s = new b200_shard; try { s->sh = std::make_shared<b200_shard_t>(device, rank, world); } catch (...) { delete s; s = nullptr; throw; } return 0; });
  return s;
}
B200_API void b200_shard_destroy(b200_shard *s) { guarded([&] { delete s; return 0; }); }
B200_API int b200_shard_nccl_unique_id(b200_shard *s, void *id_out_128) { return guarded([&] { s->sh->nccl_unique_id(id_out_128); return 0; }); }
B200_API int b200_shard_nccl_init(b200_shard *s, const void *id_128) { return guarded([&] { s->sh->nccl_init(id_128); return 0; }); }
B200_API int b200_shard_broadcast(b200_shard *s, void *dev_buf, uint64_t bytes, int root, void *stream) { return guarded([&] { s->sh->broadcast(dev_buf, bytes, root, stream); return 0; }); }
B200_API int b200_shard_all_gather_nccl(b200_shard *s, const void *dev_src, void *dev_dst, uint64_t bytes_per_rank, void *stream) {
  return guarded([&] { s->sh->all_gather_nccl(dev_src, dev_dst, bytes_per_rank, stream); return 0; });
}
B200_API int b200_shard_gather_export(b200_shard *s, uint64_t bytes_per_rank, void *ipc_handle_out_64) { return guarded([&] { s->sh->gather_export(bytes_per_rank, ipc_handle_out_64); return 0; }); }
B200_API int b200_shard_gather_import(b200_shard *s, const void *ipc_handles) { return guarded([&] { s->sh->gather_import(ipc_handles); return 0; }); }
B200_API int64_t b200_shard_gather_push(b200_shard *s, const void *dev_src, void *stream) {
  int64_t step = -1;
  int const rc = guarded([&] { step = s->sh->gather_push(dev_src, stream); return 0; });
  return rc < 0 ? rc : step;
}
B200_API int64_t b200_shard_gather_push_wait(b200_shard *s, const void *dev_src, uint32_t wait_step, void *stream) {
  int64_t step = 0;
  int const rc = guarded([&] { step = s->sh->gather_push_wait(dev_src, wait_step, stream); return 0; });
  return rc < 0 ? rc : step;
}
B200_API int b200_shard_gather_wait(b200_shard *s, uint32_t step, void *stream) { return guarded([&] { s->sh->gather_wait(step, stream); return 0; }); }
B200_API int b200_shard_gather_ptr(b200_shard *s, uint32_t step, void **dev_ptr_out) { return guarded([&] { *dev_ptr_out = s->sh->gather_ptr(step); if (!*dev_ptr_out) { rt_err("b200_shard: no gather buffer yet"); } return 0; }); }
B200_API uint64_t b200_shard_launches(b200_shard *s) { return s->sh->n_launches; }
/* parameter upload from a DEVICE buffer (the slice of the broadcast flat buffer): no host round trip */
B200_API int b200_fwd_set_param_device(b200_fwd *f, const char *node_name, const void *dev_src, uint64_t n_elems) {
  return guarded_dev(f, [&] { f->fwd->set_param_device(node_name, dev_src, n_elems); return 0; });
}
}  // extern "C"
