// b200_compute.h -- `be=b200`: a B200-native implementation of Boda's run-time compute interface.
//
// Mirrors rtc_compute_t (src/rtc_compute.H:35-97) method-for-method, with the same argument meaning and error
// behaviour as the NVRTC back-end it replaces (nvrtc_compute_t, src/nvrtc_util.cc:174-395):
//   * vars are owned by the back-end, keyed by unique name, zero-filled on creation (nvrtc_util.cc:81-84);
//     duplicate names and unknown names are errors; reshaped views share storage (:136-138)
//   * copy_nda_to_var / copy_var_to_nda require exact dims_t equality incl. names (:302-303)
//   * compile() registers functions, run() launches one and returns a call_id usable with get_dur() (ms)
//   * unsupported shapes throw unsup_exception, everything else rt_exception
// What differs by design: compile() does not JIT generated source. It binds rtc_func_info_t.op (function name or op
// type + dims + params) to a table of precompiled hand-written sm_100a kernels, and run() ignores tpb/blks: the
// B200 kernels pick their own geometry (as the culibs escape hatch does, src/nvrtc_util.cc:369-373).
#pragma once
#include "boda_base.h"
#include "b200_shard.h"
#include <cuda.h>
#include <cuda_runtime.h>

namespace boda {

struct rtc_compile_opts_t {  // src/rtc_compute.H:10-22
  uint32_t show_compile_log = 0, enable_lineinfo = 0, show_func_attrs = 0, show_rtc_calls = 0;
};

struct rtc_func_info_t {  // src/rtc_compute.H:24-29
  string func_name;       // unique (generated) name the call will use
  string func_src;        // ignored by this back-end (no run-time compilation)
  vect_string arg_names;  // ignored: the kernel table knows its arguments
  op_base_t op;
};
typedef vector<rtc_func_info_t> vect_rtc_func_info_t;

struct rtc_arg_t {  // src/rtc_compute.H:103-115: a var name, or a by-value nda
  rtc_arg_t() {}
  rtc_arg_t(string const &n_) : n(n_) {}
  rtc_arg_t(char const *n_) : n(n_) {}
  rtc_arg_t(p_nda_t const &v_) : v(v_) {}
  string n;
  p_nda_t v;
  bool is_valid() const { return bool(v) != (!n.empty()); }
  bool is_var() const { assert_st(is_valid()); return !n.empty(); }
  bool is_nda() const { assert_st(is_valid()); return bool(v); }
  string const &get_var() const { assert_st(is_var()); return n; }
  p_nda_t const &get_nda() const { assert_st(is_nda()); return v; }
};
typedef map<string, rtc_arg_t> map_str_rtc_arg_t;

struct rtc_func_call_t {  // src/rtc_compute.H:120-126
  string rtc_func_name;
  map_str_rtc_arg_t arg_map;
  uint32_t tpb = 0, blks = 0;  // accepted for signature compatibility; unused
};

struct b200_impl_t;  // kernel launchers + per-function derived state (packed operands, tensor maps)

// numeric mode of the contraction kernels
enum b200_prec_t {
  B200_PREC_FP32_SPLIT = 0,  // fp32 parity: fp16 hi/lo planes, 3 tcgen05.mma per k-step (default)
  B200_PREC_FP16 = 1,        // operands rounded to fp16, fp32 accumulate (BASELINE config C3)
  B200_PREC_BF16 = 2,        // operands rounded to bf16, fp32 accumulate (BASELINE config C4)
};

struct b200_compute_t {
  string be = "b200";
  int device = 0;
  b200_prec_t prec = B200_PREC_FP32_SPLIT;
  int fc_l2_ahead = 0;      // fc_chain kernel: the first layer's filter tiles are requested into L2 this many k-blocks ahead of the shared-memory ring (0 = off; measured: slower, r02)
  int fc_l2_next = 0;       // fc_chain kernel: request the next layer's filter tiles into L2 while the current layer drains and reduces (measured: no net gain, r02)
  int input_pack_ctas_per_sm = 2;  // absmax_pack_smallc_kernel: CTAs per SM the grid is sized for
  int fuse_input_pack = 1;  // fp32-parity / fp16 modes, few-channel network inputs: max|x|, scale and the row-merged planes in one kernel (absmax_pack_smallc_kernel)
  int debug_flags = 0;      // timing experiments on the 1-CTA kernel: 1 = skip TMA loads, 2 = skip MMA issue (results are garbage)
  int use_2cta = 1;         // CTA pairs: one tcgen05.mma.cta_group::2 per 256 x BN tile (igemm2.cuh)
  int use_clusters = 0;     // 1-CTA tiles only: let CTAs that share an operand tile form a cluster and TMA-multicast it (choose_cluster)
  int acc_chunk_kblks = 4;  // drain TMEM accumulators into fp32 registers every this many 64-wide k-blocks (fp32-parity mode)
  int acc_chunk_kblks_16 = 8;   // same for the fp16 / bf16 storage modes (32 measured 1.9e-3 mrd at K = 3456 in fp16 mode: the truncating in-TMEM accumulation drifts)
  int use_pdl = 1;          // programmatic dependent launch between the kernels of a forward pass (pdl.cuh)
  int use_taps = 0;         // stride-1 KHxKW convs: tap-reuse kernel (igemm3.cuh): one activation halo tile feeds every filter tap.
                            // Off by default: measured slower than the CTA-pair im2col kernel on the AlexNet layers (profiles/), kept for A/B runs
  int taps_max_b_stages = 8, taps_max_a_stages = 3;  // experiments: cap the tap-reuse kernel's ring depths
  int taps_2cta = -1;       // tap-reuse kernel: -1 = cost model picks single CTAs or CTA pairs, 0 / 1 = force
  int use_sk4 = 0;          // the round-2 contraction kernel (igemm4.cuh: persistent CTA pairs, halo operand mode, stage records, stream-K, two epilogue
                            // warpgroups) for every layer with >= 2 row tiles. OFF by default: parity-green, and on par with the round-1 pair kernel in the
                            // 16-bit modes (halo mode moves 2.8x fewer bytes), but measured slower in fp32-parity mode, which is MMA-bound at 3 passes and
                            // only pays the halo layout's dropped virtual pixels (profiles/diag_r02*)
  int fuse_splitk_reduce = 0;  // split-K layers (inner-product shapes): the last split CTA of a tile sums the partial tiles inside the contraction kernel
                               // (same order of additions as splitk_reduce_kernel, which then is not launched). Off by default: bit-identical and three
                               // launches fewer per AlexNet forward, but every split CTA then waits for its slowest sibling inside the kernel --
                               // measured neutral (0.394 vs 0.389 ms per step), so the two-kernel form stays
  int use_halo = 1;         // igemm4: one activation halo tile per channel block feeds every filter tap of a stride-1 KHxKW convolution
  int use_streamk = 1;      // igemm4: cut the (tile, k-block) space into equal contiguous ranges per CTA pair when whole tiles would leave > 8 % of a round idle
  int sk4_max_b_stages = 0; // experiments: cap igemm4's filter ring depth (0 = as many as fit)
  int plan_only = 0;        // host-side planning / inspection only: init() touches no device (plans for plan_num_sms SMs), vars carry dims but no
  int plan_num_sms = 148;   // storage, compile() plans as usual and every call that would compute or copy throws. Never a compute path.

  b200_compute_t();
  ~b200_compute_t();

  // One option parser for both ABI tiers (b200_rtc_set_option, the opts of b200_fwd_create). Returns false for a key this class does not know
  // (the caller may know it, else rejects it the way NESI rejects unused keys, src/nesi.cc:25-35). The reference's own op_tune_t knob names
  // (src/cnn_op.H:10-32: use_be, use_culibs, MNt, MNb, Kb, use_local_mem, prof_variant, vw, k1conv, tconv, tconv_max_ksz, ipconv) are accepted
  // and ignored -- they select among CUCL variants that do not exist here -- so existing --op-tune=(...) command lines keep working
  // (SURVEY section 5 "Config/flags"); they are listed in ignored_tune_knobs() so that a caller can log what was dropped.
  bool set_option(string const &key, string const &val);
  string const &ignored_tune_knobs() const { return ignored_knobs; }
  string ignored_knobs;

  void init();
  void bind_device();  // cudaSetDevice(device) for the calling thread (no-op before init() and on plan-only instances)
  string get_plat_tag();

  void create_var_with_dims(string const &vn, dims_t const &dims);
  void create_var_with_dims_as_reshaped_view_of_var(string const &vn, dims_t const &dims, string const &src_vn);
  void release_var(string const &vn);
  dims_t get_var_dims(string const &vn);
  void set_var_to_zero(string const &vn);

  void compile(vect_rtc_func_info_t const &func_infos, rtc_compile_opts_t const &opts);
  void release_func(string const &func_name);
  uint32_t run(rtc_func_call_t const &rfc);
  void finish_and_sync();
  void release_per_call_id_data();
  void release_all_funcs();
  float get_dur(uint32_t const &b, uint32_t const &e);  // ms, start of call b to end of call e
  float get_kernel_dur(uint32_t const &id);             // ms of the call's main (contraction) kernel alone, without operand packing
  void profile_start();
  void profile_stop();

  void copy_var_to_nda(p_nda_t const &nda, string const &vn);
  void copy_nda_to_var(string const &vn, p_nda_t const &nda);
  p_nda_t get_var_raw_native_pointer(string const &vn);
  // layered helpers (src/rtc_compute.cc:30-97)
  void create_var_from_nda(p_nda_t const &nda, string const &vn) { create_var_with_dims(vn, nda->dims); copy_nda_to_var(vn, nda); }
  p_nda_t create_nda_from_var(string const &vn) { p_nda_t r = std::make_shared<nda_t>(get_var_dims(vn)); copy_var_to_nda(r, vn); return r; }

  // --- extensions used by the C ABI / whole-net driver ---
  // can the convolution described by `op` (type + in/filts/out dims + params) also write the consumer's NHWC 16-bit plane(s) ("out_pack")? Static:
  // depends only on its plan (not an inner-product-shaped / split-K layer, out_chans a multiple of 8). Every producer of a destination var must
  // be able to, or none may be asked to -- the whole-net driver decides per destination with this.
  bool conv_plane_writable(op_base_t const &op, bool dst_is_concat);
  // whether this convolution's launch can take a residual input ("res" argument): pixel-major, un-split launches only
  bool conv_res_fusable(op_base_t const &op);
  bool conv_halo_pad(op_base_t const &op, int &py, int &px);  // halo-mode convolutions read their input planes in the shared-padding layout (py, px)
  bool conv_uses_sk4(op_base_t const &op);
  // inner-product shaped at batch <= 32 (weights as the 128-row operand, one 32-image tile, all (tile, split) units resident): a run of such
  // conv functions, each reading the previous one's output, can be one "fc_chain" function (fcchain.cuh): str parameter "layers" = their names
  // joined by ':'; call arguments "in", and per layer i "filts<i>", "biases<i>", "out<i>" (+ the abs-max cell arguments of "in" / "out<i>")
  bool func_fc_chainable(string const &fn) const;  // (of a compiled conv function)
  // multi-GPU: an fc_chain call whose last layer writes the var `out_vn` also stores those values into every rank's gather buffer and publishes
  // the step from inside its kernel (fcchain.cuh: FcGather). nullptr = off. Returns false when the descriptor does not fit (world > 8).
  bool set_tail_gather(string const &out_vn, b200_gather_desc_t const *d);
  bool has_var(string const &vn) const;
  bool has_func(string const &fn) const;
  // host-only: the launch plan compile() made for a convolution function, as "kernel=pair|single bn=<N tile> kblks=<64-wide k-blocks>
  // splits=<split-K> swapped=0|1 rowmerge=0|1 im2col=0|1 grid=<x>x<y>x<z>" ("" for other function kinds). pair = the persistent CTA-pair
  // kernel (igemm2.cuh), single = one CTA per tile (igemm.cuh). What run() launches, without launching it.
  string func_plan_text(string const &fn) const;
  void copy_raw_to_var(string const &vn, void const *src, uint64_t bytes);   // host -> device (async on the stream, then sync)
  void copy_var_to_raw(void *dst, string const &vn, uint64_t bytes);
  void copy_raw_to_var_async(string const &vn, void const *src, uint64_t bytes);
  void copy_var_to_raw_async(void *dst, string const &vn, uint64_t bytes);
  void copy_device_to_var_async(string const &vn, void const *dev_src, uint64_t bytes);  // device -> var on the back-end's stream
  cudaStream_t stream() const;
  uint64_t launches() const;  // number of kernels this back-end has launched so far (claimed in bench.py's gpu_launches)
  void set_timing(bool on);   // per-call event recording on/off (off inside CUDA-graph capture)

  b200_impl_t *impl;
};
typedef shared_ptr<b200_compute_t> p_b200_compute_t;

void rtc_reshape_check(dims_t const &dims, dims_t const &src_dims);  // src/rtc_compute.cc:24-27

}  // namespace boda
