// b200_compute.cu -- implementation of `be=b200` (see b200_compute.h) over the CUDA runtime + driver tensor-map API.
// Replaces nvrtc_compute_t (src/nvrtc_util.cc:174-395) and the culibs escape hatch (src/culibs-wrap.cc) for the
// rtc_fwd path. There is deliberately NO CPU fallback: every function either launches sm_100a kernels or throws.
#include "b200_compute.h"
#include <climits>
#include "igemm.cuh"
#include "igemm2.cuh"
#include "igemm3.cuh"
#include "igemm4.cuh"
#include "pointwise.cuh"
#include "fcchain.cuh"
#include <cudaTypedefs.h>
#include <cuda_profiler_api.h>
#include <algorithm>
#include <cmath>

namespace boda {

#define CU_CHK(x) do { cudaError_t const e_ = (x); if (e_ != cudaSuccess) { rt_err(string("CUDA error: ") + cudaGetErrorString(e_) + " in " #x " at " + __FILE__ + ":" + std::to_string(__LINE__)); } } while (0)

void rtc_reshape_check(dims_t const &dims, dims_t const &src_dims) {
  if (dims.bytes_sz() != src_dims.bytes_sz()) { rt_err("reshape: view dims " + dims.pretty() + " and source dims " + src_dims.pretty() + " differ in size"); }
}

namespace {

struct dev_buf_t {
  void *p = nullptr;
  uint64_t bytes = 0;
  explicit dev_buf_t(uint64_t b) : bytes(b) { CU_CHK(cudaMalloc(&p, std::max<uint64_t>(b, 16))); }
  ~dev_buf_t() { if (p) { cudaFree(p); } }
  dev_buf_t(dev_buf_t const &) = delete;
};
typedef shared_ptr<dev_buf_t> p_dev_buf_t;

struct var_info_t {
  p_dev_buf_t buf;
  dims_t dims;
  shared_ptr<uint64_t> gen;  // bumped on every write; shared between reshaped views of the same storage
};

// a 16-bit K-major operand in hi (+lo) planes with its power-of-two scale
struct packed_t {
  p_dev_buf_t hi, lo, scale2, absmax_bits;
  uint64_t src_gen = ~0ull;
  void const *src_ptr = nullptr;
  long long rows = 0, row_stride = 0;  // elements
  uint64_t layout_key = 0;  // geometry the planes were written for (pack_layout_key): reshaped views share storage AND generation, so a
                            // consumer reading the same storage through other dims must not reuse planes packed for this geometry
};

// FNV-1a over the arguments that determine where pack() puts every element
uint64_t pack_layout_key(std::initializer_list<long long> v) {
  uint64_t h = 1469598103934665603ull;
  for (long long x : v) { for (int i = 0; i < 8; ++i) { h ^= (uint64_t)(x >> (8 * i)) & 0xff; h *= 1099511628211ull; } }
  return h ? h : 1;
}

enum func_kind_t { FK_CONV, FK_SGEMM, FK_POOL, FK_LRN, FK_RELU, FK_SOFTMAX, FK_COPY, FK_REDUCE, FK_GEN_DATA, FK_BN_FOLD, FK_FC_CHAIN };

struct conv_plan_t {
  int N, C, H, W, OC, KH, KW, sy, sx, py, px, OH, OW;
  int Cpad, cblks;
  bool im2col, full_kernel, swapped;
  bool rowmerge;  // small-chan convs (conv1): K = (ky) x [(kx,chan) run of 64 contiguous NHWC elements], see plan_conv
  int Wp;         // rowmerge: pixel pitch of a packed image row (>= W + 2*px, and long enough for the last 64-element run)
  bool taps;      // stride-1 KHxKW conv eligible for the tap-reuse kernel (igemm3.cuh): shared-padding NHWC layout, virtual-pixel GEMM rows
  int tHp, tWp;   // taps: padded image height / row pitch (H + pad_y, W + pad_x)
  int BN, splits, kblks_total, kblks_per_split;
  int kb_mod, ksteps_last;  // K tail: see IgemmParams
  long long a_rows, a_row_stride;  // activation matrix view (2-d modes)
  long long w_tap_stride, w_row_stride;
  int relu, has_bias;
};

struct func_t {
  func_kind_t kind;
  op_base_t op;
  string gen_arg;  // gen_data: name of the output argument
  conv_plan_t cp;
  packed_t w_pack, a_pack;  // conv: filts / in ; sgemm: b / a
  p_dev_buf_t w_l1max;      // conv: max over out chans of sum |w| (bound of the outputs, for producer-written fp16 planes)
  uint64_t w_l1max_gen = ~0ull;
  p_dev_buf_t splitk_ws, splitk_tickets;
  // fc_chain: the compiled conv functions it runs as one kernel (their plans and filter packs are used), the planes between the layers, the barrier counters
  vector<string> chain_funcs;
  vector<packed_t> chain_planes;
  p_dev_buf_t chain_sync;
};

struct call_ev_t { cudaEvent_t b = nullptr, e = nullptr, kb = nullptr, ke = nullptr; };  // whole call; its main (contraction) kernel

PFN_cuTensorMapEncodeTiled_v12000 g_encode_tiled = nullptr;
PFN_cuTensorMapEncodeIm2col_v12000 g_encode_im2col = nullptr;

void load_driver_entry_points() {
  if (g_encode_tiled) { return; }
  cudaDriverEntryPointQueryResult qres;
  void *fn = nullptr;
  CU_CHK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { rt_err("cuTensorMapEncodeTiled not available from the driver"); }
  g_encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  fn = nullptr;
  CU_CHK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) { rt_err("cuTensorMapEncodeIm2col not available from the driver"); }
  g_encode_im2col = reinterpret_cast<PFN_cuTensorMapEncodeIm2col_v12000>(fn);
}

// 2-d K-major matrix [rows][row_stride] of 16-bit elements; box = 64 (K) x box_rows, 128-byte swizzle, zero OOB fill.
CUtensorMap make_tiled_map(void const *base, bool bf16, uint64_t k_extent, uint64_t rows, uint64_t row_stride_elems, uint32_t box_rows) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {k_extent, rows};
  cuuint64_t gstride[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult const r = g_encode_tiled(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim,
                                    gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rt_err("cuTensorMapEncodeTiled failed with code " + str(int(r)) + " (k=" + str(k_extent) + " rows=" + str(rows) + " stride=" + str(row_stride_elems) + ")"); }
  return m;
}

// k-block-major packed filters [n_kb][rows][64] of 16-bit elements as a 3-d tensor; box = 64 x box_rows x box_kb, 128-byte swizzle
CUtensorMap make_tiled_map3(void const *base, bool bf16, uint64_t rows, uint64_t n_kb, uint32_t box_rows, uint32_t box_kb) {
  CUtensorMap m;
  cuuint64_t gdim[3] = {64, rows, n_kb};
  cuuint64_t gstride[2] = {128, rows * 128};
  cuuint32_t box[3] = {64, box_rows, box_kb};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult const r = g_encode_tiled(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), gdim, gstride, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rt_err("cuTensorMapEncodeTiled (3-d) failed with code " + str(int(r)) + " (rows=" + str(rows) + " kblks=" + str(n_kb) + ")"); }
  return m;
}

// NHWC activation tensor [N][H][W][Cpad] read in im2col mode: 128 output pixels x 64 channels per load.
CUtensorMap make_im2col_map(void const *base, bool bf16, conv_plan_t const &cp, uint32_t pixels_per_col = 128) {
  CUtensorMap m;
  cuuint64_t gdim[4] = {(cuuint64_t)cp.Cpad, (cuuint64_t)cp.W, (cuuint64_t)cp.H, (cuuint64_t)cp.N};
  cuuint64_t gstride[3] = {(cuuint64_t)cp.Cpad * 2, (cuuint64_t)cp.W * cp.Cpad * 2, (cuuint64_t)cp.H * cp.W * cp.Cpad * 2};
  int lower[2] = {-cp.px, -cp.py};                                // {w, h}: footprint corner of the first output pixel
  int upper[2] = {cp.px - (cp.KW - 1), cp.py - (cp.KH - 1)};      // ... of the last one, relative to the far image edge
  cuuint32_t estr[4] = {1, (cuuint32_t)cp.sx, (cuuint32_t)cp.sy, 1};
  if (cp.rowmerge) {
    // virtual tensor: "channel" = run of 64 contiguous elements of a packed image row, "w" = output column (runs overlap: stride sx*Cpad
    // elements), x padding is materialised in the packed row, y padding is the im2col corner as usual; the filter is 1 wide, KH tall.
    gdim[0] = 64; gdim[1] = (cuuint64_t)cp.OW;
    gstride[0] = (cuuint64_t)cp.sx * cp.Cpad * 2; gstride[1] = (cuuint64_t)cp.Wp * cp.Cpad * 2; gstride[2] = (cuuint64_t)cp.H * cp.Wp * cp.Cpad * 2;
    lower[0] = 0; upper[0] = 0;
    estr[1] = 1;
  }
  CUresult const r = g_encode_im2col(&m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), gdim, gstride,
                                     lower, upper, 64 /*channelsPerPixel*/, pixels_per_col /*pixelsPerColumn*/, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { rt_err("cuTensorMapEncodeIm2col failed with code " + str(int(r))); }
  // driver <= 13.1 quirk for small tensors (same workaround CUTLASS applies when it builds im2col descriptors)
  int drv = 0;
  cudaDriverGetVersion(&drv);
  if (drv <= 13010 && (uint64_t)cp.N * cp.H * (cp.rowmerge ? cp.Wp : cp.W) * cp.Cpad * 2 < 131072) { reinterpret_cast<uint64_t *>(&m)[1] &= ~(1ull << 21); }
  return m;
}

// Every kernel of this back-end asks for the maximum shared-memory carve-out: the contraction kernels need ~200 KB of dynamic shared
// memory, and letting the bandwidth kernels between them run with the default (L1-heavy) split makes the SMs re-partition L1/shared
// memory at every layer boundary, which drains the SM and costs microseconds per switch.
template <typename F> void prefer_max_smem(F *func) {
  cudaFuncSetAttribute(reinterpret_cast<void const *>(func), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
// cudaFuncSetAttribute is per device (context): the "already done" state is a bit per device ordinal, so a second instance on another GPU of
// the same process gets its own opt-ins (carve-out preference, > 48 KB dynamic shared memory)
inline bool first_use_on_device(uint64_t &mask, int dev) { uint64_t const b = 1ull << (dev & 63); if (mask & b) { return false; } mask |= b; return true; }
#define B200_CARVEOUT_ONCE(...) do { static uint64_t done_ = 0; if (first_use_on_device(done_, rtc.device)) { prefer_max_smem(__VA_ARGS__); } } while (0)

int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }
long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

// CTA-cluster shape for the contraction kernel. The kernel is bound by L2->SM operand traffic long before the tensor pipe saturates
// (a 128 x BN tile needs planes*128*(128+BN) bytes per k-block for planes^2-1|1 MMA passes), so CTAs that share an operand tile fetch one
// slice each and multicast it: cm CTAs along P-tiles share every Q tile, cn CTAs along Q-tiles share every P tile.
// Cost model: waves x max(MMA cycles, bytes / measured L2 delivery per SM) per k-block; clusters must fit whole GPCs.
struct cluster_choice_t { int cm = 1, cn = 1; };
cluster_choice_t choose_cluster(int p_tiles, int q_tiles, int BN, int planes, bool allow) {
  cluster_choice_t best;
  if (!allow) { return best; }
  double best_cost = 1e30;
  for (int cm : {1, 2, 4}) {
    for (int cn : {1, 2, 4}) {
      int const sz = cm * cn;
      if (sz > 8 || (BN / cm) % 8 != 0 || BN % cm != 0) { continue; }
      if (round_up(p_tiles, cm) * 3 > (long long)p_tiles * 4 || round_up(q_tiles, cn) * 3 > (long long)q_tiles * 4) { continue; }  // <= 33 % padding CTAs
      int const per_wave = sz == 1 ? 148 : sz == 2 ? 74 : sz == 4 ? 33 : 16;
      long long const clusters = (long long)ceil_div(p_tiles, cm) * ceil_div(q_tiles, cn);
      double const waves = std::ceil((double)clusters / per_wave);
      double const bytes = planes * 128.0 * (128.0 / cn + (double)BN / cm);
      double const mma = (planes == 2 ? 12.0 : 4.0) * BN / 2.0;
      double const cost = waves * std::max(mma, bytes / 33.0) + 0.01 * sz;
      if (cost < best_cost) { best_cost = cost; best.cm = cm; best.cn = cn; }
    }
  }
  return best;
}


// ---- launch plan of the round-2 kernel (igemm4.cuh) ----------------------------------------------------------------------------------------
struct sk4_plan_t {
  bool use = false, halo = false, sk = false;
  int BN = 0, a_stages = 0, b_stages = 0;
  int halo_rows = 0, a_loads = 0, a_box_rows = 0, Hp = 0, Wp = 0;
  long long m_rows = 0;  // halo: virtual pixels that may hold an output
  int m_pair_tiles = 0, q_tiles = 0, n_tiles = 0, ukb = 0, n_pairs = 0;
  int ku = 0;  // k-blocks per stage (ku * planes operand slots)
  size_t smem = 0;
};

sk4_plan_t plan_sk4(conv_plan_t const &cp, int planes, int num_sms, b200_compute_t const &rtc) {
  sk4_plan_t sp;
  if (!rtc.use_sk4 || !rtc.use_2cta || num_sms < 2) { return sp; }
  long long const pixels = (long long)cp.N * cp.OH * cp.OW;
  // halo mode: stride 1, window > 1x1, padding no larger than the window overhang, an acceptable share of dropped virtual pixels
  if (rtc.use_halo && cp.im2col && !cp.rowmerge && !cp.swapped && cp.sx == 1 && cp.sy == 1 && cp.KH * cp.KW > 1 && cp.px <= cp.KW - 1 && cp.py <= cp.KH - 1) {
    int const Hp = cp.H + cp.py, Wp = cp.W + cp.px;
    double const waste = (double)Hp * Wp / ((double)cp.OH * cp.OW);
    int halo_rows = (int)round_up(128 + (long long)(cp.KH - 1) * Wp + cp.KW - 1, 8);
    if (waste <= 1.5 && halo_rows <= 768 && (long long)cp.N * Hp * Wp < (1ll << 31)) {
      sp.halo = true; sp.Hp = Hp; sp.Wp = Wp;
      sp.a_loads = ceil_div(halo_rows, 256); sp.a_box_rows = (int)round_up(ceil_div(halo_rows, sp.a_loads), 8);
      sp.halo_rows = sp.a_loads * sp.a_box_rows;
      sp.m_rows = (long long)(cp.N - 1) * Hp * Wp + (long long)(cp.OH - 1) * Wp + cp.OW;
    }
  }
  for (int attempt = 0; attempt < 2; ++attempt) {  // second attempt: without the halo mode (its rings did not fit shared memory)
    long long const p_rows = cp.swapped ? cp.OC : (sp.halo ? sp.m_rows : pixels), q_rows = cp.swapped ? pixels : cp.OC;
    int const p_tiles = ceil_div(p_rows, b200::IGEMM_BM);
    if (p_tiles < 2) { return sk4_plan_t(); }
    if (cp.swapped) { sp.BN = cp.BN; }
    else {  // tile width over out_chans: least padding, where an MMA narrower than 64 columns costs the issue slot of a 64-wide one; ties -> wider
      long long best_cost = 0;
      for (int bn : {128, 96, 64, 32}) {
        long long const cost = (long long)ceil_div(cp.OC, bn) * std::max(bn, 64);
        if (!sp.BN || cost < best_cost) { sp.BN = bn; best_cost = cost; }
      }
    }
    // k-blocks per stage: 4 operand slots in halo mode (16 single-plane / 24 fp32-parity MMAs per barrier round trip -- the MMA warp's per-stage
    // cost is paid once per stage), 2 in the per-k-block operand modes (their stages also carry 16 KB activation tiles)
    sp.ku = (sp.halo ? 4 : 2) / planes;
    long long const avail = 225 * 1024 - 1024 - b200::SK4_BAR_BYTES, b_stage = (long long)sp.ku * planes * (sp.BN / 2) * 128;
    if (sp.halo) {
      long long const a_stage = (long long)planes * sp.halo_rows * 128;
      sp.a_stages = 0;
      for (int as : {2, 1}) {
        long long const bs = std::min<long long>((avail - as * a_stage) / b_stage, b200::SK4_MAX_B_STAGES);
        if (bs >= (as == 2 ? 3 : 2)) { sp.a_stages = as; sp.b_stages = (int)bs; break; }
      }
      if (sp.a_stages == 2 && (avail - 3 * a_stage) / b_stage >= 5) { sp.a_stages = 3; sp.b_stages = (int)std::min<long long>((avail - 3 * a_stage) / b_stage, b200::SK4_MAX_B_STAGES); }
      if (rtc.sk4_max_b_stages > 0) { sp.b_stages = std::min(sp.b_stages, rtc.sk4_max_b_stages); }
      if (!sp.a_stages) { sp.halo = false; sp.BN = 0; continue; }
      sp.smem = (size_t)(sp.a_stages * a_stage + sp.b_stages * b_stage + 1024 + b200::SK4_BAR_BYTES);
    } else {
      long long const stage = (long long)sp.ku * planes * b200::IGEMM_BM * 128 + b_stage;
      sp.b_stages = (int)std::min<long long>(avail / stage, b200::SK4_MAX_A_STAGES);  // modes 0/1: the P slots ride on the Q stages (ring depths are equal)
      sp.a_stages = sp.b_stages;
      if (sp.b_stages < 2) { return sk4_plan_t(); }
      sp.smem = (size_t)(sp.b_stages * stage + 1024 + b200::SK4_BAR_BYTES);
    }
    sp.m_pair_tiles = ceil_div(p_tiles, 2); sp.q_tiles = ceil_div(q_rows, sp.BN);
    sp.n_tiles = sp.m_pair_tiles * sp.q_tiles;
    int const nkb = sp.halo ? cp.cblks * cp.KH * cp.KW : cp.kblks_total;
    sp.ukb = ceil_div(nkb, sp.ku);
    int const P = num_sms / 2;
    long long const U = (long long)sp.n_tiles * sp.ukb;
    // whole tiles cost ceil(tiles / pairs) rounds of ukb stage-units; stream-K costs U / pairs units (>= 2 per pair) plus about two units' worth of
    // partial-tile hand-off (a 64 KB write + read per split tile). Stream-K when that is at least 8 % faster.
    int const pairs_sk = (int)std::min<long long>(P, U / 2);
    double const dp_time = (double)ceil_div(sp.n_tiles, P) * sp.ukb, sk_time = pairs_sk > 0 ? (double)ceil_div(U, pairs_sk) + 2.0 : 1e30;
    if (U * (P + 1) >= (1ll << 31)) { return sk4_plan_t(); }  // the kernel's range arithmetic is 32-bit
    sp.sk = rtc.use_streamk && pairs_sk >= 2 && (sk_time < 0.92 * dp_time || rtc.use_streamk == 2) && sp.BN <= 128;  // (use_streamk=2: always, for tests)
    sp.n_pairs = sp.sk ? pairs_sk : std::min(sp.n_tiles, P);
    sp.use = true;
    return sp;
  }
  return sk4_plan_t();
}

}  // namespace

struct b200_impl_t {
  map<string, var_info_t> vars;
  map<string, func_t> funcs;
  bool tail_gather_on = false;  // set_tail_gather
  string tail_gather_vn;
  b200_gather_desc_t tail_gather;
  // NHWC 16-bit planes of activation vars, shared by every Convolution that reads the same var (the four branches of an inception module, the
  // shortcut + first 1x1 of a ResNet block): the first consumer of a write generation packs, the others reuse. Keyed by the storage pointer.
  // (second key: 0 = plain NHWC [pixel][chan]; pad_tag(py, px) = the shared-padding layout a halo-mode convolution reads, igemm4.cuh)
  map<std::pair<void const *, uint32_t>, packed_t> act_packs;
  static uint32_t pad_tag(int py, int px) { return 1u + (uint32_t)py * 256u + (uint32_t)px; }
  p_dev_buf_t sk_ws, sk_flags;  // stream-K partial-tile workspace + hand-off flags of igemm_sk4_kernel (one launch at a time per stream)
  vector<call_ev_t> calls;
  cudaStream_t stream = nullptr;
  bool inited = false, timing = true;
  uint64_t n_launches = 0;
  int num_sms = 148;
  string plat_tag;

  var_info_t &must_var(string const &vn) {
    auto i = vars.find(vn);
    if (i == vars.end()) { rt_err("var '" + vn + "' not found"); }
    return i->second;
  }
  void bump(var_info_t &v) { ++(*v.gen); }
};

b200_compute_t::b200_compute_t() : impl(new b200_impl_t) {}
b200_compute_t::~b200_compute_t() {
  if (impl) {
    if (impl->inited && !plan_only) { cudaStreamSynchronize(impl->stream); }
    release_per_call_id_data();
    impl->funcs.clear();
    impl->vars.clear();
    if (impl->stream) { cudaStreamDestroy(impl->stream); }
    delete impl;
  }
}

bool b200_compute_t::set_option(string const &k, string const &v) {
  if (k == "prec") { prec = (v == "fp32") ? B200_PREC_FP32_SPLIT : (v == "fp16") ? B200_PREC_FP16 : (v == "bf16") ? B200_PREC_BF16 : (rt_err("unknown prec '" + v + "'"), B200_PREC_FP32_SPLIT); }
  else if (k == "acc_chunk_kblks") { acc_chunk_kblks = std::stoi(v); }
  else if (k == "acc_chunk_kblks_16") { acc_chunk_kblks_16 = std::stoi(v); }
  else if (k == "use_taps") { use_taps = std::stoi(v); }
  else if (k == "use_pdl") { use_pdl = std::stoi(v); }
  else if (k == "taps_2cta") { taps_2cta = std::stoi(v); }
  else if (k == "taps_max_b_stages") { taps_max_b_stages = std::stoi(v); }
  else if (k == "taps_max_a_stages") { taps_max_a_stages = std::stoi(v); }
  else if (k == "use_clusters") { use_clusters = std::stoi(v); }
  else if (k == "use_2cta") { use_2cta = std::stoi(v); }
  else if (k == "debug_flags") { debug_flags = std::stoi(v); }
  else if (k == "input_pack_ctas_per_sm") { input_pack_ctas_per_sm = std::max(1, std::stoi(v)); }
  else if (k == "fuse_input_pack") { fuse_input_pack = std::stoi(v); }
  else if (k == "fc_l2_ahead") { fc_l2_ahead = std::stoi(v); }
  else if (k == "fc_l2_next") { fc_l2_next = std::stoi(v); }
  else if (k == "use_sk4") { use_sk4 = std::stoi(v); }
  else if (k == "fuse_splitk_reduce") { fuse_splitk_reduce = std::stoi(v); }
  else if (k == "use_halo") { use_halo = std::stoi(v); }
  else if (k == "use_streamk") { use_streamk = std::stoi(v); }
  else if (k == "sk4_max_b_stages") { sk4_max_b_stages = std::stoi(v); }
  else if (k == "device") { device = std::stoi(v); }
  else if (k == "plan_only") { plan_only = std::stoi(v); }
  else if (k == "plan_num_sms") { plan_num_sms = std::stoi(v); }
  else if (k == "use_be") {  // src/cnn_op.H:13: "if non-empty, use this tune only with the specific named backend"
    if (!v.empty() && v != "b200") { rt_err("op_tune use_be='" + v + "' names another back-end; this is be=b200"); }
    ignored_knobs += (ignored_knobs.empty() ? "" : ",") + k;
  }
  else if (k == "use_culibs" || k == "MNt" || k == "MNb" || k == "Kb" || k == "use_local_mem" || k == "prof_variant" || k == "vw" || k == "k1conv" || k == "tconv" ||
           k == "tconv_max_ksz" || k == "ipconv") {
    ignored_knobs += (ignored_knobs.empty() ? "" : ",") + k;  // CUCL variant selection: one hand-written kernel family here, chosen by plan_conv
  }
  else { return false; }
  return true;
}

void b200_compute_t::init() {
  if (impl->inited) { return; }
  if (plan_only) {  // no device: launch plans only (b200_fwd_plan); run() and every copy refuse
    impl->num_sms = plan_num_sms;
    impl->plat_tag = "b200:plan-only";
    impl->inited = true;
    return;
  }
  int ndev = 0;
  cudaError_t const e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { rt_err(string("be=b200 needs a CUDA device and there is no CPU fallback: ") + cudaGetErrorString(e)); }
  CU_CHK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_CHK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) { unsup_err("be=b200 is built for sm_100a only; device is sm_" + str(prop.major) + str(prop.minor)); }
  impl->num_sms = prop.multiProcessorCount;
  impl->plat_tag = string("b200:") + prop.name;
  CU_CHK(cudaStreamCreateWithFlags(&impl->stream, cudaStreamNonBlocking));
  load_driver_entry_points();
  impl->inited = true;
}
void b200_compute_t::bind_device() { if (impl->inited && !plan_only) { CU_CHK(cudaSetDevice(device)); } }
string b200_compute_t::get_plat_tag() { return impl->plat_tag.empty() ? string("b200:uninit") : impl->plat_tag; }
cudaStream_t b200_compute_t::stream() const { return impl->stream; }
uint64_t b200_compute_t::launches() const { return impl->n_launches; }
void b200_compute_t::set_timing(bool on) { impl->timing = on; }
bool b200_compute_t::has_var(string const &vn) const { return impl->vars.count(vn) != 0; }
bool b200_compute_t::has_func(string const &fn) const { return impl->funcs.count(fn) != 0; }

// ---- vars -----------------------------------------------------------------------------------------------------
void b200_compute_t::create_var_with_dims(string const &vn, dims_t const &dims) {
  assert_st(impl->inited);
  if (impl->vars.count(vn)) { rt_err("var '" + vn + "' already exists"); }
  if (dims.tn == "none") { rt_err("can't create var '" + vn + "' with no element type"); }
  var_info_t v;
  v.dims = dims;
  v.dims.calc_strides();
  v.gen = std::make_shared<uint64_t>(1);
  if (!plan_only) {
    v.buf = std::make_shared<dev_buf_t>(dims.bytes_sz());
    CU_CHK(cudaMemsetAsync(v.buf->p, 0, std::max<uint64_t>(dims.bytes_sz(), 16), impl->stream));  // new vars are zero-filled
  }
  impl->vars[vn] = v;
}
void b200_compute_t::create_var_with_dims_as_reshaped_view_of_var(string const &vn, dims_t const &dims, string const &src_vn) {
  if (impl->vars.count(vn)) { rt_err("var '" + vn + "' already exists"); }
  var_info_t const &src = impl->must_var(src_vn);
  rtc_reshape_check(dims, src.dims);
  var_info_t v = src;  // shares buffer + generation counter
  v.dims = dims;
  v.dims.calc_strides();
  impl->vars[vn] = v;
}
void b200_compute_t::release_var(string const &vn) {
  var_info_t &v = impl->must_var(vn);
  if (v.buf && v.buf.use_count() == 1) {  // last name of this storage: drop its packed planes too (every layout)
    for (auto i = impl->act_packs.begin(); i != impl->act_packs.end();) { if (i->first.first == v.buf->p) { i = impl->act_packs.erase(i); } else { ++i; } }
  }
  impl->vars.erase(vn);
}
dims_t b200_compute_t::get_var_dims(string const &vn) { return impl->must_var(vn).dims; }
void b200_compute_t::set_var_to_zero(string const &vn) {
  if (plan_only) { rt_err("plan-only instance (no device): vars have no storage; there is no CPU fallback"); }
  var_info_t &v = impl->must_var(vn);
  CU_CHK(cudaMemsetAsync(v.buf->p, 0, v.dims.bytes_sz(), impl->stream));
  impl->bump(v);
}
void b200_compute_t::copy_nda_to_var(string const &vn, p_nda_t const &nda) {
  var_info_t &v = impl->must_var(vn);
  if (!(nda->dims == v.dims)) { rt_err("copy_nda_to_var: dims mismatch for '" + vn + "': nda " + nda->dims.pretty() + " vs var " + v.dims.pretty()); }
  copy_raw_to_var(vn, nda->rp_elems(), v.dims.bytes_sz());
}
void b200_compute_t::copy_var_to_nda(p_nda_t const &nda, string const &vn) {
  var_info_t &v = impl->must_var(vn);
  if (!(nda->dims == v.dims)) { rt_err("copy_var_to_nda: dims mismatch for '" + vn + "': nda " + nda->dims.pretty() + " vs var " + v.dims.pretty()); }
  copy_var_to_raw(nda->rp_elems(), vn, v.dims.bytes_sz());
}
void b200_compute_t::copy_raw_to_var_async(string const &vn, void const *src, uint64_t bytes) {
  if (plan_only) { rt_err("plan-only instance (no device): vars have no storage; there is no CPU fallback"); }
  var_info_t &v = impl->must_var(vn);
  if (bytes != v.dims.bytes_sz()) { rt_err("copy to var '" + vn + "': got " + str(bytes) + " bytes, var holds " + str(v.dims.bytes_sz())); }
  CU_CHK(cudaMemcpyAsync(v.buf->p, src, bytes, cudaMemcpyHostToDevice, impl->stream));
  impl->bump(v);
}
void b200_compute_t::copy_var_to_raw_async(void *dst, string const &vn, uint64_t bytes) {
  if (plan_only) { rt_err("plan-only instance (no device): vars have no storage; there is no CPU fallback"); }
  var_info_t &v = impl->must_var(vn);
  if (bytes != v.dims.bytes_sz()) { rt_err("copy from var '" + vn + "': asked " + str(bytes) + " bytes, var holds " + str(v.dims.bytes_sz())); }
  CU_CHK(cudaMemcpyAsync(dst, v.buf->p, bytes, cudaMemcpyDeviceToHost, impl->stream));
}
void b200_compute_t::copy_device_to_var_async(string const &vn, void const *dev_src, uint64_t bytes) {
  if (plan_only) { rt_err("plan-only instance (no device): vars have no storage; there is no CPU fallback"); }
  var_info_t &v = impl->must_var(vn);
  if (bytes != v.dims.bytes_sz()) { rt_err("copy to var '" + vn + "': got " + str(bytes) + " bytes, var holds " + str(v.dims.bytes_sz())); }
  CU_CHK(cudaMemcpyAsync(v.buf->p, dev_src, bytes, cudaMemcpyDeviceToDevice, impl->stream));
  impl->bump(v);
}
void b200_compute_t::copy_raw_to_var(string const &vn, void const *src, uint64_t bytes) {
  copy_raw_to_var_async(vn, src, bytes);
  CU_CHK(cudaStreamSynchronize(impl->stream));
}
void b200_compute_t::copy_var_to_raw(void *dst, string const &vn, uint64_t bytes) {
  copy_var_to_raw_async(dst, vn, bytes);
  CU_CHK(cudaStreamSynchronize(impl->stream));
}
p_nda_t b200_compute_t::get_var_raw_native_pointer(string const &vn) {
  if (plan_only) { rt_err("plan-only instance (no device): vars have no storage; there is no CPU fallback"); }
  var_info_t &v = impl->must_var(vn);
  impl->bump(v);  // the caller may write through the pointer: invalidate anything derived from this var
  p_nda_t r = std::make_shared<nda_t>(v.dims, false);
  r->rp = v.buf->p;
  return r;
}

// ---- functions ------------------------------------------------------------------------------------------------
namespace {

bool starts_with(string const &s, string const &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }

func_kind_t resolve_kind(op_base_t const &op, string &gen_arg) {
  string fn = op.has_func_name() ? op.get_func_name() : string();
  string const type = op.has_type() ? op.get_type() : string();
  if (starts_with(fn, "gen_data_")) {
    size_t const us = fn.rfind('_');
    gen_arg = fn.substr(us + 1);
    return FK_GEN_DATA;
  }
  if (fn.empty()) {
    if (type == "Convolution") { fn = "conv"; } else if (type == "sgemm") { fn = "sgemm"; } else if (type == "Pooling") { fn = "pool"; }
    else if (type == "LRN") { fn = "lrn"; } else if (type == "ReLU") { fn = "relu"; } else if (type == "Softmax") { fn = "softmax"; }
    else if (type == "Concat") { fn = "copy"; } else if (type == "Reduce" || type == "Eltwise") { fn = "reduce"; }
    else { unsup_err("be=b200: no function for op of type '" + type + "'"); }
  }
  // every reference conv variant computes the same function; ours takes the reference-layout (NCHW / OIHW) arguments,
  // i.e. the same call contract as the `cudnn_conv` variant (src/cnn_op.cc:46-50, src/culibs-wrap.cc:94-212)
  if (fn == "conv" || fn == "tconv" || fn == "k1conv" || fn == "ipconv" || fn == "conv_simd" || fn == "k1conv_simd" || fn == "cudnn_conv" || fn == "b200_conv") { return FK_CONV; }
  if (fn == "sgemm" || fn == "sgemm_simd" || fn == "sgemm_no_local" || fn == "sgemm_simd_local" || fn == "cublas_sgemm" || fn == "b200_sgemm") { return FK_SGEMM; }
  if (fn == "pool") { return FK_POOL; }
  if (fn == "lrn") { return FK_LRN; }
  if (fn == "relu") { return FK_RELU; }
  if (fn == "softmax") { return FK_SOFTMAX; }
  if (fn == "copy") { return FK_COPY; }
  if (fn == "reduce") { return FK_REDUCE; }
  if (fn == "bn_fold") { return FK_BN_FOLD; }
  if (fn == "fc_chain") { return FK_FC_CHAIN; }
  unsup_err("be=b200: unknown function '" + fn + "'");
}

void check_nchw(dims_t const &d, char const *what) {
  if (d.size() != 4 || d[0].name != "img" || d[1].name != "chan" || d[2].name != "y" || d[3].name != "x" || d.tn != "float") {
    unsup_err(string("be=b200 conv expects reference-layout float img:chan:y:x for '") + what + "', got " + d.pretty());
  }
}

void plan_conv(conv_plan_t &cp, op_base_t const &op, int num_sms) {
  dims_t const &din = op.get_dims("in"), &df = op.get_dims("filts"), &dout = op.get_dims("out");
  check_nchw(din, "in");
  check_nchw(dout, "out");
  if (df.size() != 4 || df[0].name != "out_chan" || df[1].name != "in_chan" || df[2].name != "y" || df[3].name != "x" || df.tn != "float") {
    unsup_err("be=b200 conv expects reference-layout float out_chan:in_chan:y:x filts, got " + df.pretty());
  }
  cp.N = din.dsz("img"); cp.C = din.dsz("chan"); cp.H = din.dsz("y"); cp.W = din.dsz("x");
  cp.OC = df.dsz("out_chan"); cp.KH = df.dsz("y"); cp.KW = df.dsz("x");
  if ((int)df.dsz("in_chan") != cp.C) { unsup_err("conv: filts in_chan != in chan (groups are not supported by the reference either)"); }
  cp.sy = op.yx("stride", "y", 1); cp.sx = op.yx("stride", "x", 1);
  cp.py = op.yx("in_pad", "y", 0); cp.px = op.yx("in_pad", "x", 0);
  if (op.has("kern_sz") && ((int)op.get_dims("kern_sz").dsz("y") != cp.KH || (int)op.get_dims("kern_sz").dsz("x") != cp.KW)) { rt_err("conv: kern_sz disagrees with filts dims"); }
  if (cp.H + 2 * cp.py < cp.KH || cp.W + 2 * cp.px < cp.KW) { rt_err("conv: padded input smaller than kernel"); }
  cp.OH = (cp.H + 2 * cp.py - cp.KH) / cp.sy + 1;  // src/conv_util.cc:167-173
  cp.OW = (cp.W + 2 * cp.px - cp.KW) / cp.sx + 1;
  if ((int)dout.dsz("img") != cp.N || (int)dout.dsz("chan") != cp.OC || (int)dout.dsz("y") != cp.OH || (int)dout.dsz("x") != cp.OW) {
    rt_err("conv: out dims " + dout.pretty() + " do not match computed img=" + str(cp.N) + ",chan=" + str(cp.OC) + ",y=" + str(cp.OH) + ",x=" + str(cp.OW));
  }
  if (cp.KW > 255 || cp.KH > 255 || cp.px > 127 || cp.py > 127) { unsup_err("conv: kernel/padding too large for the im2col TMA path"); }
  cp.relu = op.has("conv_has_relu") ? op.get_u32("conv_has_relu") : 0;
  cp.has_bias = 1;
  cp.Cpad = (int)round_up(cp.C, 8);
  bool const k1 = (cp.KH == 1 && cp.KW == 1 && cp.sy == 1 && cp.sx == 1 && cp.py == 0 && cp.px == 0);
  cp.full_kernel = (!k1 && cp.KH == cp.H && cp.KW == cp.W && cp.py == 0 && cp.px == 0);  // inner-product shaped (OH=OW=1)
  cp.im2col = !(k1 || cp.full_kernel);
  long long const pixels = (long long)cp.N * cp.OH * cp.OW;
  // Row-merged path for few-channel inputs (AlexNet/NiN conv1 11x11 s4, GoogLeNet conv1 7x7 s2): with chan padded to only 4 (or 8) an NHWC
  // image row is [x][chan] contiguous, so the (kx, chan) taps of one filter row are ONE contiguous run of <= 64 elements starting at pixel
  // ox*sx. The K loop then has KH k-blocks of 64 instead of KH*KW k-blocks that are mostly zero padding (chan 3 of 64).
  cp.rowmerge = false;
  cp.Wp = 0;
  if (cp.im2col) {
    int const cpx = (cp.C <= 4 && (cp.sx % 2) == 0) ? 4 : (cp.C <= 8 ? 8 : 0);  // x stride in bytes (sx*cpx*2) must be a multiple of 16 for TMA
    if (cpx && cp.KW * cpx <= 64 && cp.KW > 1) {
      cp.rowmerge = true;
      cp.Cpad = cpx;
      cp.Wp = (int)round_up(std::max(cp.W + 2 * cp.px, (cp.OW - 1) * cp.sx + 64 / cpx), 16 / cpx);
    }
  }
  if (cp.rowmerge) {
    cp.cblks = 1;
    cp.w_tap_stride = cp.Cpad;  // (kx,chan) packed densely inside a 64-element filter row
    cp.kblks_total = cp.KH;
    cp.a_rows = 0; cp.a_row_stride = 0;
  } else if (cp.im2col) {
    cp.cblks = ceil_div(cp.Cpad, 64);
    cp.w_tap_stride = (long long)cp.cblks * 64;
    cp.kblks_total = cp.KH * cp.KW * cp.cblks;
    cp.a_rows = 0; cp.a_row_stride = 0;
  } else if (k1) {
    cp.cblks = ceil_div(cp.Cpad, 64);
    cp.w_tap_stride = cp.Cpad;
    cp.kblks_total = cp.cblks;
    cp.a_rows = (long long)cp.N * cp.H * cp.W; cp.a_row_stride = cp.Cpad;
  } else {
    cp.cblks = 0;
    cp.w_tap_stride = cp.Cpad;
    cp.kblks_total = ceil_div((long long)cp.KH * cp.KW * cp.Cpad, 64);
    cp.a_rows = cp.N; cp.a_row_stride = (long long)cp.H * cp.W * cp.Cpad;
  }
  cp.w_row_stride = cp.rowmerge ? (long long)cp.KH * 64 : round_up((long long)cp.KH * cp.KW * cp.w_tap_stride, 64);
  {  // real data in the last k-block of a group (rest: zero padding in the packed operands / TMA out-of-bounds fill)
    long long const k_grp = cp.rowmerge ? (long long)cp.KW * cp.Cpad : (cp.im2col || k1) ? (long long)cp.Cpad : (long long)cp.KH * cp.KW * cp.Cpad;
    cp.kb_mod = cp.rowmerge ? 1 : ceil_div(k_grp, 64);
    cp.ksteps_last = ceil_div(k_grp - (long long)(cp.kb_mod - 1) * 64, 16);
  }
  cp.swapped = (!cp.im2col) && pixels <= 64 && cp.OC >= 128;
  if (cp.swapped) { cp.BN = pixels <= 32 ? 32 : 64; }
  else {  // tile width over out_chans: per k-block a tile costs max(MMA cycles ~ BN, issue overhead) + the activation tile it re-reads, and the
          // persistent pair kernel runs ceil(tiles / SM pairs) rounds of tiles -- so a narrower tile can win by filling the chip (AlexNet
          // conv5: 44 tiles of 128 on 74 pairs vs 66 tiles of 96) as well as by removing padding; ties go to less padding, then wider tiles
    int best = 0;
    long long best_cost = 0, best_cols = 0;
    long long const m_pairs = ceil_div(ceil_div(pixels, b200::IGEMM_BM), 2), slots = std::max(num_sms / 2, 1);
    // (measured, r02: counting the fp32-parity mode's three MMAs per k-step here -- which moves AlexNet conv2 from 184 tiles of 128 = 3 rounds
    // to 368 tiles of 64 = 5 half rounds -- made the step 8 % SLOWER: the tensor pipe does not run 64-wide MMAs at twice the rate of 128-wide ones)
    for (int bn : {128, 96, 64, 32}) {
      long long const tiles_n = ceil_div(cp.OC, bn), cols = tiles_n * bn;
      long long const cost = ceil_div(m_pairs * tiles_n, slots) * (std::max(4 * bn, 300) + 64);
      if (!best || cost < best_cost || (cost == best_cost && cols < best_cols)) { best = bn; best_cost = cost; best_cols = cols; }
    }
    cp.BN = best;
  }
  long long const p_rows = cp.swapped ? cp.OC : pixels, q_rows = cp.swapped ? pixels : cp.OC;
  long long const tiles = (long long)ceil_div(p_rows, b200::IGEMM_BM) * ceil_div(q_rows, cp.BN);
  cp.splits = 1;
  if (tiles * 2 <= num_sms && cp.kblks_total >= 16) {  // too few CTAs to fill the chip: split K (deterministic two-pass reduce)
    int s = (int)std::min<long long>((long long)(num_sms / tiles), (long long)(cp.kblks_total / 8));
    cp.splits = std::max(1, std::min(s, 16));
  }
  cp.kblks_per_split = ceil_div(cp.kblks_total, cp.splits);
  cp.splits = ceil_div(cp.kblks_total, cp.kblks_per_split);
  // tap-reuse kernel: stride 1, window > 1x1, padding no larger than the window overhang, and an acceptable share of dropped virtual pixels
  cp.taps = false; cp.tHp = cp.H + cp.py; cp.tWp = cp.W + cp.px;
  if (cp.im2col && !cp.rowmerge && cp.sx == 1 && cp.sy == 1 && cp.KH * cp.KW > 1 && cp.px <= cp.KW - 1 && cp.py <= cp.KH - 1 && cp.splits == 1 && !cp.swapped) {
    double const waste = (double)cp.tHp * cp.tWp / ((double)cp.OH * cp.OW);
    long long const halo = round_up(128 + (long long)(cp.KH - 1) * cp.tWp + cp.KW - 1, 8);
    if (waste <= 1.5 && halo <= 768 && (long long)cp.N * cp.tHp * cp.tWp < (1ll << 31)) { cp.taps = true; }
  }
}

// a layer fc_chain_kernel can run: weights as the 128-row operand, all images in one 32-wide tile, plain 2-d K-major operands, one CTA per
// (tile, split) unit with all units resident at once
bool fc_chainable_plan(conv_plan_t const &cp, int num_sms) {
  long long const pixels = (long long)cp.N * cp.OH * cp.OW;
  long long const tiles = ceil_div(cp.OC, b200::IGEMM_BM);
  return cp.swapped && !cp.im2col && !cp.rowmerge && cp.BN == b200::FC_BN && pixels <= b200::FC_BN && cp.OH == 1 && cp.OW == 1 && tiles * cp.splits <= num_sms &&
         (cp.OC % 4) == 0 && cp.splits <= 16;
}

}  // namespace

void b200_compute_t::compile(vect_rtc_func_info_t const &func_infos, rtc_compile_opts_t const &) {
  assert_st(impl->inited);
  for (auto const &fi : func_infos) {
    if (impl->funcs.count(fi.func_name)) { rt_err("function '" + fi.func_name + "' already compiled"); }
    func_t f;
    f.op = fi.op;
    f.kind = resolve_kind(fi.op, f.gen_arg);
    if (f.kind == FK_CONV) { plan_conv(f.cp, fi.op, impl->num_sms); }
    if (f.kind == FK_FC_CHAIN) {  // "layers" = the names of already compiled conv functions, first to last, separated by ':'
      string const &ls = fi.op.get_str("layers");
      for (size_t b = 0; b <= ls.size();) {
        size_t e = ls.find(':', b);
        if (e == string::npos) { e = ls.size(); }
        string const n = ls.substr(b, e - b);
        auto li = impl->funcs.find(n);
        if (li == impl->funcs.end() || li->second.kind != FK_CONV) { rt_err("fc_chain '" + fi.func_name + "': layer '" + n + "' is not a compiled conv function"); }
        if (!fc_chainable_plan(li->second.cp, impl->num_sms)) { unsup_err("fc_chain '" + fi.func_name + "': layer '" + n + "' is not inner-product shaped at batch <= 32 (conv_fc_chainable)"); }
        f.chain_funcs.push_back(n);
        b = e + 1;
      }
      if (f.chain_funcs.size() < 2 || f.chain_funcs.size() > (size_t)b200::FC_MAX_LAYERS) { unsup_err("fc_chain '" + fi.func_name + "': 2.." + str(b200::FC_MAX_LAYERS) + " layers"); }
      f.chain_planes.resize(f.chain_funcs.size());
    }
    impl->funcs[fi.func_name] = f;
  }
}
string b200_compute_t::func_plan_text(string const &fn) const {
  auto fi = impl->funcs.find(fn);
  if (fi == impl->funcs.end()) { rt_err("func_plan_text: function '" + fn + "' was not compiled"); }
  if (fi->second.kind != FK_CONV) { return string(); }
  conv_plan_t const &cp = fi->second.cp;
  long long const pixels = (long long)cp.N * cp.OH * cp.OW;
  sk4_plan_t const sp = plan_sk4(cp, prec == B200_PREC_FP32_SPLIT ? 2 : 1, impl->num_sms, *this);
  if (sp.use) {  // the round-2 kernel: what run() launches, without launching it
    return string("kernel=sk4 mode=") + (sp.halo ? "halo" : cp.im2col ? "im2col" : "2d") + " bn=" + str(sp.BN) + " kblks=" + str(sp.halo ? cp.cblks * cp.KH * cp.KW : cp.kblks_total) +
           " tiles=" + str(sp.n_tiles) + " streamk=" + str((int)sp.sk) + " a_stages=" + str(sp.a_stages) + " b_stages=" + str(sp.b_stages) + " splits=1 swapped=" + str((int)cp.swapped) +
           " rowmerge=" + str((int)cp.rowmerge) + " im2col=" + str((int)cp.im2col) + " grid=" + str(2 * sp.n_pairs) + "x1x1";
  }
  int const p_rows_n = cp.swapped ? cp.OC : (int)pixels, q_rows_n = cp.swapped ? (int)pixels : cp.OC;
  int const p_tiles = ceil_div(p_rows_n, b200::IGEMM_BM), q_tiles = ceil_div(q_rows_n, cp.BN);
  bool const two_cta = use_2cta && !cp.swapped && cp.splits == 1 && p_tiles >= 2;  // the rule of run_conv
  string grid;
  if (two_cta) { grid = str(2 * std::min(ceil_div(p_tiles, 2) * q_tiles, impl->num_sms / 2)) + "x1x1"; }
  else { grid = str(p_tiles) + "x" + str(q_tiles) + "x" + str(cp.splits); }
  return string("kernel=") + (two_cta ? "pair" : "single") + " bn=" + str(cp.BN) + " kblks=" + str(cp.kblks_total) + " splits=" + str(cp.splits) + " swapped=" + str((int)cp.swapped) +
         " rowmerge=" + str((int)cp.rowmerge) + " im2col=" + str((int)cp.im2col) + " grid=" + grid;
}
// the shared-padding plane layout (py, px) this convolution's halo mode reads its input in; false = it reads plain NHWC planes
bool b200_compute_t::conv_halo_pad(op_base_t const &op, int &py, int &px) {
  conv_plan_t cp;
  plan_conv(cp, op, impl->num_sms);
  sk4_plan_t const sp = plan_sk4(cp, prec == B200_PREC_FP32_SPLIT ? 2 : 1, impl->num_sms, *this);
  if (!sp.use || !sp.halo) { return false; }
  py = cp.py; px = cp.px;
  return true;
}
// whether this convolution is run by the round-2 kernel (which can write its consumers' planes in the shared-padding layout)
bool b200_compute_t::conv_uses_sk4(op_base_t const &op) {
  conv_plan_t cp;
  plan_conv(cp, op, impl->num_sms);
  return plan_sk4(cp, prec == B200_PREC_FP32_SPLIT ? 2 : 1, impl->num_sms, *this).use;
}
bool b200_compute_t::conv_plane_writable(op_base_t const &op, bool dst_is_concat) {
  conv_plan_t cp;
  plan_conv(cp, op, impl->num_sms);
  // the fp16 planes of the fp32-parity / fp16 modes carry a per-tensor scale that each producer derives from its own output bound: several
  // producers of one Concat output would not agree on it, so those modes write planes for a convolution's own output node only
  if (prec != B200_PREC_BF16 && dst_is_concat) { return false; }
  return !cp.swapped && cp.splits == 1 && (cp.OC % 8) == 0;
}
bool b200_compute_t::set_tail_gather(string const &out_vn, b200_gather_desc_t const *d) {
  impl->tail_gather_on = false;
  if (!d) { return true; }
  if (d->world < 1 || d->world > 8) { return false; }
  impl->tail_gather = *d;
  impl->tail_gather_vn = out_vn;
  impl->tail_gather_on = true;
  return true;
}
bool b200_compute_t::func_fc_chainable(string const &fn) const {
  auto fi = impl->funcs.find(fn);
  return fi != impl->funcs.end() && fi->second.kind == FK_CONV && use_sk4 == 0 && fc_chainable_plan(fi->second.cp, impl->num_sms);
}
bool b200_compute_t::conv_res_fusable(op_base_t const &op) {
  conv_plan_t cp;
  plan_conv(cp, op, impl->num_sms);
  return !cp.swapped && cp.splits == 1;
}
void b200_compute_t::release_func(string const &func_name) {
  if (!impl->funcs.erase(func_name)) { rt_err("release_func: '" + func_name + "' not found"); }
}
void b200_compute_t::release_all_funcs() { impl->funcs.clear(); }
void b200_compute_t::finish_and_sync() { if (!plan_only) { CU_CHK(cudaStreamSynchronize(impl->stream)); } }
void b200_compute_t::release_per_call_id_data() {
  for (auto &c : impl->calls) { for (cudaEvent_t ev : {c.b, c.e, c.kb, c.ke}) { if (ev) { cudaEventDestroy(ev); } } }
  impl->calls.clear();
}
float b200_compute_t::get_dur(uint32_t const &b, uint32_t const &e) {
  if (b >= impl->calls.size() || e >= impl->calls.size()) { rt_err("get_dur: invalid call id"); }
  if (!impl->calls[b].b || !impl->calls[e].e) { rt_err("get_dur: call was not timed"); }
  CU_CHK(cudaEventSynchronize(impl->calls[e].e));
  float ms = 0;
  CU_CHK(cudaEventElapsedTime(&ms, impl->calls[b].b, impl->calls[e].e));
  return ms;
}
float b200_compute_t::get_kernel_dur(uint32_t const &id) {
  if (id >= impl->calls.size()) { rt_err("get_kernel_dur: invalid call id"); }
  call_ev_t const &c = impl->calls[id];
  if (!c.kb || !c.ke) { return get_dur(id, id); }  // single-kernel functions: the call is the kernel
  CU_CHK(cudaEventSynchronize(c.ke));
  float ms = 0;
  CU_CHK(cudaEventElapsedTime(&ms, c.kb, c.ke));
  return ms;
}
// src/nvrtc_util.cc:388-389 (cuProfilerStart / cuProfilerStop): brackets the region a profiler launched with --profile-from-start off captures
void b200_compute_t::profile_start() { if (!plan_only) { CU_CHK(cudaSetDevice(device)); CU_CHK(cudaProfilerStart()); } }
void b200_compute_t::profile_stop() { if (!plan_only) { CU_CHK(cudaSetDevice(device)); CU_CHK(cudaProfilerStop()); } }

// ---- run ------------------------------------------------------------------------------------------------------
namespace {

struct run_ctx_t {
  b200_compute_t &rtc;
  b200_impl_t &im;
  func_t &f;
  rtc_func_call_t const &rfc;
  cudaStream_t st;
  call_ev_t *ev;
  void mark_kernel_begin() { if (ev && im.timing) { CU_CHK(cudaEventCreate(&ev->kb)); CU_CHK(cudaEventRecord(ev->kb, st)); } }
  void mark_kernel_end() { if (ev && im.timing) { CU_CHK(cudaEventCreate(&ev->ke)); CU_CHK(cudaEventRecord(ev->ke, st)); } }

  var_info_t &var(string const &an) {
    auto i = rfc.arg_map.find(an);
    if (i == rfc.arg_map.end() || !i->second.is_var()) { rt_err("call to '" + rfc.rtc_func_name + "': missing var argument '" + an + "'"); }
    return im.must_var(i->second.get_var());
  }
  bool has_arg(string const &an) { return rfc.arg_map.count(an) != 0; }
  double scalar(string const &an, bool has_default = false, double dflt = 0) {
    auto i = rfc.arg_map.find(an);
    if (i != rfc.arg_map.end() && i->second.is_nda()) { return nda_scalar_as_double(*i->second.get_nda()); }
    if (f.op.has(an) && f.op.get(an)->has_data()) { return nda_scalar_as_double(*f.op.get(an)); }
    if (has_default) { return dflt; }
    rt_err("call to '" + rfc.rtc_func_name + "': missing by-value argument '" + an + "'");
  }
  void launched(int n = 1) {
    im.n_launches += n;
    cudaError_t const e = cudaGetLastError();
    if (e != cudaSuccess) { rt_err(string("kernel launch failed in '") + rfc.rtc_func_name + "': " + cudaGetErrorString(e)); }
  }
  // optional abs-max side channel: args "<which>_absmax_cells" (uint32_t var) + by-value "<which>_absmax_ix"
  unsigned int *absmax_cell(string const &which) {
    string const an = which + "_absmax_cells";
    if (!has_arg(an)) { return nullptr; }
    var_info_t &v = var(an);
    if (v.dims.tn != "uint32_t") { rt_err("call to '" + rfc.rtc_func_name + "': '" + an + "' must be a uint32_t var"); }
    uint64_t const ix = (uint64_t)scalar(which + "_absmax_ix");
    if (ix >= v.dims.dims_prod()) { rt_err("call to '" + rfc.rtc_func_name + "': '" + which + "_absmax_ix' out of range"); }
    return static_cast<unsigned int *>(v.buf->p) + ix;
  }
  float *fptr(var_info_t &v) { if (v.dims.tn != "float") { unsup_err("be=b200: only float vars are supported, got " + v.dims.pretty()); } return static_cast<float *>(v.buf->p); }

  // Every kernel goes through here: programmatic dependent launch (pdl.cuh) lets it be scheduled while its predecessor drains.
  template <typename... KArgs, typename... Args>
  void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rtc.use_pdl ? 1 : 0;
    CU_CHK(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  }

  // src [B][R][Cc] fp32 -> planes [B][Cc][..R..] 16-bit, with abs-max scaling. Cached on (pointer, generation).
  void pack(packed_t &pk, var_info_t &src, int B, int R, int Cc, int Rpad, long long dst_c_stride, long long dst_b_stride, long long total_elems, bool want_lo, bool bf16,
            int c_inner = 0, long long dst_chi_stride = 0, long long dst_base = 0, unsigned int const *absmax_src = nullptr, int smallc_W = 0, long long kmajor_rows = 0, int tap_minor = 0) {
    if (c_inner <= 0) { c_inner = std::max(Cc, 1); }
    uint64_t const lkey = pack_layout_key({B, R, Cc, Rpad, dst_c_stride, dst_b_stride, c_inner, dst_chi_stride, dst_base, want_lo, bf16, smallc_W, kmajor_rows, tap_minor});
    if (pk.src_gen == *src.gen && pk.src_ptr == src.buf->p && pk.hi && pk.layout_key == lkey && (!want_lo || pk.lo)) { return; }
    if (pk.hi && pk.layout_key != lkey && pk.hi->bytes >= (uint64_t)total_elems * 2 && (!want_lo || pk.lo)) {  // same storage, other geometry: padding positions must be zero again
      CU_CHK(cudaMemsetAsync(pk.hi->p, 0, pk.hi->bytes, st));
      if (pk.lo) { CU_CHK(cudaMemsetAsync(pk.lo->p, 0, pk.lo->bytes, st)); }
    }
    pk.layout_key = lkey;
    if (!pk.hi || pk.hi->bytes < (uint64_t)total_elems * 2 || (want_lo && !pk.lo)) {
      pk.hi = std::make_shared<dev_buf_t>(total_elems * 2);
      CU_CHK(cudaMemsetAsync(pk.hi->p, 0, total_elems * 2, st));
      if (want_lo) { pk.lo = std::make_shared<dev_buf_t>(total_elems * 2); CU_CHK(cudaMemsetAsync(pk.lo->p, 0, total_elems * 2, st)); }
      pk.scale2 = std::make_shared<dev_buf_t>(8);
      pk.absmax_bits = std::make_shared<dev_buf_t>(16);  // {max bits, blocks-done counter, blocks-left counter (absmax_pack_smallc_kernel)}
      CU_CHK(cudaMemsetAsync(pk.absmax_bits->p, 0, 16, st));
      if (bf16) { static float const ones[2] = {1.0f, 1.0f}; CU_CHK(cudaMemcpyAsync(pk.scale2->p, ones, 8, cudaMemcpyHostToDevice, st)); }  // bf16 has fp32's exponent range: no scaling, ever
    }
    long long const n = (long long)B * R * Cc;
    int const blocks = (int)std::min<long long>((n + 4095) / 4096, 148 * 8);
    bool const use_scale = !bf16;  // bf16 has fp32's exponent range: no scaling needed
    if (!use_scale) { absmax_src = nullptr; }
    uint16_t *hi0 = static_cast<uint16_t *>(pk.hi->p), *lo0 = want_lo ? static_cast<uint16_t *>(pk.lo->p) : nullptr;
    if (!absmax_src && use_scale && rtc.fuse_input_pack && smallc_W > 0 && (Rpad == 4 || Rpad == 8) && dst_base % Rpad == 0 &&
        (reinterpret_cast<uintptr_t>(fptr(src)) & 15) == 0) {
      // network input (few channels, row-merged layout): max|x|, the scale and the planes in ONE kernel -- every CTA keeps its rows in shared
      // memory across a grid-wide barrier, so the input is read once and the step has a launch fewer (absmax_pack_smallc_kernel)
      int const H = Cc / smallc_W, Wp = (int)(dst_chi_stride / Rpad), px_off = (int)(dst_base / Rpad);
      // row groups per image: the fewest that put input_pack_ctas_per_sm CTAs on every SM (every CTA arrives on ONE counter: with 1056 CTAs the
      // barrier's serialised atomics were a quarter of the kernel), more only when the rows would not fit shared memory
      int upi_first = 1;
      while (upi_first < H && (long long)B * upi_first < (long long)im.num_sms * rtc.input_pack_ctas_per_sm) { ++upi_first; }
      for (int upi = upi_first; upi <= H; ++upi) {
        int const rows = ceil_div(H, upi);
        if (ceil_div(H, rows) != upi) { continue; }
        size_t const smem = (size_t)R * b200::smallc_row_stride(rows, smallc_W) * 4;
        if (smem > 200 * 1024) { continue; }
        auto kern = Rpad == 4 ? b200::absmax_pack_smallc_kernel<4> : b200::absmax_pack_smallc_kernel<8>;
        static uint64_t attr4_ = 0, attr8_ = 0;
        if (first_use_on_device(Rpad == 4 ? attr4_ : attr8_, rtc.device)) { CU_CHK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); prefer_max_smem(kern); }
        int per_sm = 0;
        CU_CHK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
        if ((long long)B * upi > (long long)im.num_sms * per_sm) { break; }  // (more groups only shrink per_sm's slack further)
        launch_k(kern, dim3(B * upi), dim3(256), smem, fptr(src), hi0, lo0, static_cast<float *>(pk.scale2->p), static_cast<unsigned int *>(pk.absmax_bits->p), R, H, smallc_W, Wp, px_off, rows, upi,
                 (long long)B * R * Cc);
        launched();
        pk.src_gen = *src.gen;
        pk.src_ptr = src.buf->p;
        return;
      }
    }
    if (!absmax_src) {  // nobody published max|x| for this tensor: reduce it here
      if (use_scale) {  // max|x| and the scale in one launch (the last block finalises)
        B200_CARVEOUT_ONCE(b200::absmax_kernel); launch_k(b200::absmax_kernel, dim3(std::max(blocks, 1)), dim3(256), 0, fptr(src), n, static_cast<unsigned int *>(pk.absmax_bits->p), static_cast<float *>(pk.scale2->p));
        launched();
      }  // bf16: scale2 = {1, 1} was written when the planes were allocated
    }
    dim3 grid(ceil_div(Rpad, 64), ceil_div(Cc, 32), B);
    uint16_t *hi = static_cast<uint16_t *>(pk.hi->p), *lo = want_lo ? static_cast<uint16_t *>(pk.lo->p) : nullptr;
    if (smallc_W > 0 && (Rpad == 4 || Rpad == 8)) {  // row-merged conv input: one thread per pixel, vector stores
      int const H = Cc / smallc_W, Wp = (int)(dst_chi_stride / Rpad), px_off = (int)(dst_base / Rpad);
      long long const n_pix = (long long)B * Cc;
      float *sc = static_cast<float *>(pk.scale2->p);
      int const nb = ceil_div(n_pix, 256);
      if (Rpad == 4 && !bf16) { B200_CARVEOUT_ONCE(b200::pack_smallc_kernel<4, false>); launch_k(b200::pack_smallc_kernel<4, false>, dim3(nb), dim3(256), 0, fptr(src), hi, lo, sc, R, H, smallc_W, Wp, px_off, n_pix, absmax_src); }
      else if (Rpad == 4) { B200_CARVEOUT_ONCE(b200::pack_smallc_kernel<4, true>); launch_k(b200::pack_smallc_kernel<4, true>, dim3(nb), dim3(256), 0, fptr(src), hi, lo, sc, R, H, smallc_W, Wp, px_off, n_pix, absmax_src); }
      else if (!bf16) { B200_CARVEOUT_ONCE(b200::pack_smallc_kernel<8, false>); launch_k(b200::pack_smallc_kernel<8, false>, dim3(nb), dim3(256), 0, fptr(src), hi, lo, sc, R, H, smallc_W, Wp, px_off, n_pix, absmax_src); }
      else { B200_CARVEOUT_ONCE(b200::pack_smallc_kernel<8, true>); launch_k(b200::pack_smallc_kernel<8, true>, dim3(nb), dim3(256), 0, fptr(src), hi, lo, sc, R, H, smallc_W, Wp, px_off, n_pix, absmax_src); }
      launched();
      pk.src_gen = *src.gen;
      pk.src_ptr = src.buf->p;
      return;
    }
    if (Cc == 1 && dst_base == 0) {  // rows are already K-major: elementwise scale + split
      long long const nn = (long long)B * R;
      if (bf16) { B200_CARVEOUT_ONCE(b200::pack_rows_split_kernel<true>); launch_k(b200::pack_rows_split_kernel<true>, dim3(ceil_div(nn, 256)), dim3(256), 0, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, dst_b_stride, nn, absmax_src, kmajor_rows); }
      else { B200_CARVEOUT_ONCE(b200::pack_rows_split_kernel<false>); launch_k(b200::pack_rows_split_kernel<false>, dim3(ceil_div(nn, 256)), dim3(256), 0, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, dst_b_stride, nn, absmax_src, kmajor_rows); }
      launched();
      pk.src_gen = *src.gen;
      pk.src_ptr = src.buf->p;
      return;
    }
    if (kmajor_rows == 0 && dst_chi_stride == 0 && dst_base == 0 && (Cc % 4) == 0 && Cc >= 512 && c_inner >= Cc) {  // big plain-NHWC activations: wide tiles
      dim3 const g4(ceil_div(Rpad, 64), ceil_div(Cc, 128), B);
      size_t const smem4 = 64 * 129 * sizeof(float);
      if (bf16) { B200_CARVEOUT_ONCE(b200::pack_xpose_split_v4_kernel<true>); launch_k(b200::pack_xpose_split_v4_kernel<true>, g4, dim3(256), smem4, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, Cc, Rpad, dst_c_stride, dst_b_stride, absmax_src); }
      else { B200_CARVEOUT_ONCE(b200::pack_xpose_split_v4_kernel<false>); launch_k(b200::pack_xpose_split_v4_kernel<false>, g4, dim3(256), smem4, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, Cc, Rpad, dst_c_stride, dst_b_stride, absmax_src); }
      launched();
      pk.src_gen = *src.gen;
      pk.src_ptr = src.buf->p;
      return;
    }
    if (bf16) { B200_CARVEOUT_ONCE(b200::pack_xpose_split_kernel<true>); launch_k(b200::pack_xpose_split_kernel<true>, dim3(grid), dim3(256), 0, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, Cc, Rpad, dst_c_stride, dst_b_stride, c_inner, dst_chi_stride, dst_base, absmax_src, kmajor_rows, tap_minor); }
    else { B200_CARVEOUT_ONCE(b200::pack_xpose_split_kernel<false>); launch_k(b200::pack_xpose_split_kernel<false>, dim3(grid), dim3(256), 0, fptr(src), hi, lo, static_cast<float *>(pk.scale2->p), R, Cc, Rpad, dst_c_stride, dst_b_stride, c_inner, dst_chi_stride, dst_base, absmax_src, kmajor_rows, tap_minor); }
    launched();
    pk.src_gen = *src.gen;
    pk.src_ptr = src.buf->p;
  }

  template <int BN, int kPlanes>
  void launch_igemm_t(dim3 grid, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::IgemmParams const &prm) {
    using Cfg = b200::IgemmCfg<BN, kPlanes>;
    static uint64_t attr_set = 0;
    if (first_use_on_device(attr_set, rtc.device)) {
      CU_CHK(cudaFuncSetAttribute(b200::igemm_umma_kernel<BN, kPlanes>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(b200::IGEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (prm.cm * prm.cn > 1) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = prm.cm; attr[na].val.clusterDim.y = prm.cn; attr[na].val.clusterDim.z = 1; ++na; }
    if (rtc.use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CU_CHK(cudaLaunchKernelEx(&cfg, b200::igemm_umma_kernel<BN, kPlanes>, ph, pl, qh, ql, prm));
    launched();
  }
  template <int BN, int kPlanes>
  void launch_igemm2_t(dim3 grid, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::IgemmParams const &prm) {
    using Cfg = b200::Igemm2Cfg<BN, kPlanes>;
    static uint64_t attr_set = 0;
    if (first_use_on_device(attr_set, rtc.device)) {
      CU_CHK(cudaFuncSetAttribute(b200::igemm_umma_2cta_kernel<BN, kPlanes>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(b200::IGEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rtc.use_pdl ? 2 : 1;
    CU_CHK(cudaLaunchKernelEx(&cfg, b200::igemm_umma_2cta_kernel<BN, kPlanes>, ph, pl, qh, ql, prm));
    launched();
  }
  template <int BN, int kPlanes, int kEpi>
  void launch_sk4_t(sk4_plan_t const &sp, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::Sk4Params const &prm) {
    static uint64_t attr_set = 0;
    if (first_use_on_device(attr_set, rtc.device)) {
      CU_CHK(cudaFuncSetAttribute(b200::igemm_sk4_kernel<BN, kPlanes, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * sp.n_pairs, 1, 1);
    cfg.blockDim = dim3(b200::SK4_THREADS);
    cfg.dynamicSmemBytes = sp.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rtc.use_pdl ? 2 : 1;
    CU_CHK(cudaLaunchKernelEx(&cfg, b200::igemm_sk4_kernel<BN, kPlanes, kEpi>, ph, pl, qh, ql, prm));
    launched();
  }
  template <int kPlanes>
  void launch_sk4_p(int BN, sk4_plan_t const &sp, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::Sk4Params const &prm) {
    // the epilogue variant compiled into the kernel that is launched (igemm4.cuh: instruction-cache footprint)
    if (prm.g.swapped) {
      if (BN == 64) { launch_sk4_t<64, kPlanes, b200::SK4_EPI_SWAPPED>(sp, ph, pl, qh, ql, prm); } else if (BN == 32) { launch_sk4_t<32, kPlanes, b200::SK4_EPI_SWAPPED>(sp, ph, pl, qh, ql, prm); }
      else { rt_err("igemm4: swapped launches use 32- or 64-wide tiles"); }
    } else if (prm.g.out16 || prm.g.res) {
      if (BN == 128) { launch_sk4_t<128, kPlanes, b200::SK4_EPI_FULL>(sp, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_sk4_t<96, kPlanes, b200::SK4_EPI_FULL>(sp, ph, pl, qh, ql, prm); }
      else if (BN == 64) { launch_sk4_t<64, kPlanes, b200::SK4_EPI_FULL>(sp, ph, pl, qh, ql, prm); } else { launch_sk4_t<32, kPlanes, b200::SK4_EPI_FULL>(sp, ph, pl, qh, ql, prm); }
    } else {
      if (BN == 128) { launch_sk4_t<128, kPlanes, b200::SK4_EPI_LEAN>(sp, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_sk4_t<96, kPlanes, b200::SK4_EPI_LEAN>(sp, ph, pl, qh, ql, prm); }
      else if (BN == 64) { launch_sk4_t<64, kPlanes, b200::SK4_EPI_LEAN>(sp, ph, pl, qh, ql, prm); } else { launch_sk4_t<32, kPlanes, b200::SK4_EPI_LEAN>(sp, ph, pl, qh, ql, prm); }
    }
  }
  void launch_sk4(int BN, int planes, sk4_plan_t const &sp, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::Sk4Params const &prm) {
    if (planes == 2) { launch_sk4_p<2>(BN, sp, ph, pl, qh, ql, prm); } else { launch_sk4_p<1>(BN, sp, ph, pl, qh, ql, prm); }
  }
  void launch_igemm2(int BN, int planes, dim3 grid, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::IgemmParams const &prm) {
    if (planes == 2) {
      if (BN == 128) { launch_igemm2_t<128, 2>(grid, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_igemm2_t<96, 2>(grid, ph, pl, qh, ql, prm); } else if (BN == 64) { launch_igemm2_t<64, 2>(grid, ph, pl, qh, ql, prm); } else { launch_igemm2_t<32, 2>(grid, ph, pl, qh, ql, prm); }
    } else {
      if (BN == 128) { launch_igemm2_t<128, 1>(grid, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_igemm2_t<96, 1>(grid, ph, pl, qh, ql, prm); } else if (BN == 64) { launch_igemm2_t<64, 1>(grid, ph, pl, qh, ql, prm); } else { launch_igemm2_t<32, 1>(grid, ph, pl, qh, ql, prm); }
    }
  }
  void launch_igemm(int BN, int planes, dim3 grid, CUtensorMap const &ph, CUtensorMap const &pl, CUtensorMap const &qh, CUtensorMap const &ql, b200::IgemmParams const &prm) {
    if (planes == 2) {
      if (BN == 128) { launch_igemm_t<128, 2>(grid, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_igemm_t<96, 2>(grid, ph, pl, qh, ql, prm); } else if (BN == 64) { launch_igemm_t<64, 2>(grid, ph, pl, qh, ql, prm); } else { launch_igemm_t<32, 2>(grid, ph, pl, qh, ql, prm); }
    } else {
      if (BN == 128) { launch_igemm_t<128, 1>(grid, ph, pl, qh, ql, prm); } else if (BN == 96) { launch_igemm_t<96, 1>(grid, ph, pl, qh, ql, prm); } else if (BN == 64) { launch_igemm_t<64, 1>(grid, ph, pl, qh, ql, prm); } else { launch_igemm_t<32, 1>(grid, ph, pl, qh, ql, prm); }
    }
  }

  // ---- tap-reuse path (igemm3.cuh): returns false when the shape does not fit its shared-memory budget (caller falls back to im2col) ----
  template <int BN, int kPlanes, bool k2>
  void launch_taps_t(dim3 grid, size_t smem, CUtensorMap const &ah, CUtensorMap const &al, CUtensorMap const &wh, CUtensorMap const &wl, b200::TapsParams const &prm) {
    static uint64_t attr_set = 0;
    if (first_use_on_device(attr_set, rtc.device)) {
      CU_CHK(cudaFuncSetAttribute(b200::igemm_taps_kernel<BN, kPlanes, k2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(b200::IGEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (k2) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na; }
    if (rtc.use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    CU_CHK(cudaLaunchKernelEx(&cfg, b200::igemm_taps_kernel<BN, kPlanes, k2>, ah, al, wh, wl, prm));
    launched();
  }
  template <int BN>
  void launch_taps_bn(int planes, bool k2, dim3 grid, size_t smem, CUtensorMap const &ah, CUtensorMap const &al, CUtensorMap const &wh, CUtensorMap const &wl, b200::TapsParams const &prm) {
    if (planes == 2) { if (k2) { launch_taps_t<BN, 2, true>(grid, smem, ah, al, wh, wl, prm); } else { launch_taps_t<BN, 2, false>(grid, smem, ah, al, wh, wl, prm); } }
    else { if (k2) { launch_taps_t<BN, 1, true>(grid, smem, ah, al, wh, wl, prm); } else { launch_taps_t<BN, 1, false>(grid, smem, ah, al, wh, wl, prm); } }
  }

  struct taps_stage_t { int a_stages = 0, b_stages = 0; size_t smem = 0; };
  taps_stage_t taps_staging(int halo_rows, int planes, int bq) {
    taps_stage_t r;
    long long const avail = 224 * 1024 - 1024 - b200::TAPS_BAR_BYTES, a_stage = (long long)planes * halo_rows * 128, b_stage = (long long)planes * bq * 128;
    for (int as : {3, 2, 1}) {
      if (as > rtc.taps_max_a_stages) { continue; }
      long long const bs = std::min<long long>(std::min<long long>((avail - as * a_stage) / b_stage, b200::TAPS_MAX_B_STAGES), rtc.taps_max_b_stages);
      if (bs >= (as == 3 ? 4 : 3)) { r.a_stages = as; r.b_stages = (int)bs; r.smem = (size_t)(as * a_stage + bs * b_stage + 1024 + b200::TAPS_BAR_BYTES); return r; }
    }
    return r;
  }

  bool run_conv_taps(var_info_t &vin, var_info_t &vf, var_info_t &vout, float const *bias, bool bf16, int planes) {
    conv_plan_t const &cp = f.cp;
    int const Hp = cp.tHp, Wp = cp.tWp, taps = cp.KH * cp.KW, BN = cp.BN;
    int halo_rows = (int)round_up(128 + (long long)(cp.KH - 1) * Wp + cp.KW - 1, 8);
    int const a_loads = ceil_div(halo_rows, 256), a_box_rows = (int)round_up(ceil_div(halo_rows, a_loads), 8);
    halo_rows = a_loads * a_box_rows;
    long long const m_rows = (long long)(cp.N - 1) * Hp * Wp + (long long)(cp.OH - 1) * Wp + cp.OW;
    int const p_tiles = ceil_div(m_rows, b200::IGEMM_BM), q_tiles = ceil_div(cp.OC, BN);
    // single CTAs or CTA pairs: waves x max(MMA cycles, operand bytes / sustained L2->SM rate) per 64-channel block
    taps_stage_t const st1 = taps_staging(halo_rows, planes, BN), st2 = (BN >= 32) ? taps_staging(halo_rows, planes, BN / 2) : taps_stage_t();
    double const mma = (planes == 2 ? 3.0 : 1.0) * taps * 2.0 * BN;
    double const cost1 = st1.a_stages ? std::ceil((double)p_tiles * q_tiles / im.num_sms) * std::max(mma, planes * 128.0 * (halo_rows + (double)taps * BN) / 36.0) : 1e30;
    double const cost2 = (st2.a_stages && p_tiles >= 2) ? std::ceil((double)ceil_div(p_tiles, 2) * q_tiles / (im.num_sms / 2)) * std::max(mma, planes * 128.0 * (halo_rows + (double)taps * BN / 2) / 36.0) : 1e30;
    bool k2 = cost2 < cost1;
    if (rtc.taps_2cta == 0 && st1.a_stages) { k2 = false; }
    if (rtc.taps_2cta == 1 && st2.a_stages && p_tiles >= 2) { k2 = true; }
    taps_stage_t const stg = k2 ? st2 : st1;
    if (!stg.a_stages) { return false; }

    // filters: OIHW -> [OC][tap][chan] K-major rows, once per weight version; activations: NCHW -> shared-padding NHWC (see igemm3.cuh)
    long long const oc_pad = round_up(cp.OC, 128);  // filters are packed k-block-major: [k-block][out_chan padded][64]
    pack(f.w_pack, vf, cp.OC, cp.C, taps, cp.Cpad, cp.w_tap_stride, cp.w_row_stride, oc_pad * cp.w_row_stride, planes == 2, bf16, 0, 0, 0, nullptr, 0, oc_pad);
    long long const img_elems = (long long)Hp * Wp * cp.Cpad;
    pack(f.a_pack, vin, cp.N, cp.C, cp.H * cp.W, cp.Cpad, cp.Cpad, img_elems, (long long)cp.N * img_elems, planes == 2, bf16, cp.W, (long long)Wp * cp.Cpad,
         ((long long)cp.py * Wp + cp.px) * cp.Cpad, absmax_cell("in"));
    CUtensorMap const a_hi = make_tiled_map(f.a_pack.hi->p, bf16, cp.Cpad, (uint64_t)cp.N * Hp * Wp, cp.Cpad, a_box_rows);
    CUtensorMap const a_lo = planes == 2 ? make_tiled_map(f.a_pack.lo->p, bf16, cp.Cpad, (uint64_t)cp.N * Hp * Wp, cp.Cpad, a_box_rows) : a_hi;
    uint32_t const w_box = k2 ? BN / 2 : BN;
    uint64_t const w_rows = (uint64_t)(cp.w_row_stride / 64) * oc_pad;
    CUtensorMap const w_hi = make_tiled_map(f.w_pack.hi->p, bf16, 64, w_rows, 64, w_box);
    CUtensorMap const w_lo = planes == 2 ? make_tiled_map(f.w_pack.lo->p, bf16, 64, w_rows, 64, w_box) : w_hi;

    b200::TapsParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.m_rows = (int)m_rows; prm.q_rows = cp.OC;
    prm.cblks = cp.cblks; prm.taps = taps; prm.kw = cp.KW;
    prm.ksteps_last = cp.ksteps_last; prm.w_kb_rows = (int)oc_pad;
    prm.Wp = Wp; prm.HpWp = Hp * Wp; prm.OH = cp.OH; prm.OW = cp.OW;
    prm.halo_rows = halo_rows; prm.a_loads = a_loads; prm.a_box_rows = a_box_rows;
    prm.a_stages = stg.a_stages; prm.b_stages = stg.b_stages;
    prm.chunk_kblks = std::max(1, planes == 2 ? rtc.acc_chunk_kblks : rtc.acc_chunk_kblks_16);
    prm.out_chans = cp.OC; prm.out_hw = cp.OH * cp.OW;
    prm.relu = cp.relu; prm.has_bias = bias ? 1 : 0; prm.bias = bias;
    prm.out = fptr(vout);
    prm.p_scale = static_cast<float *>(f.a_pack.scale2->p);
    prm.q_scale = static_cast<float *>(f.w_pack.scale2->p);
    prm.out_absmax = absmax_cell("out");
    prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, k2 ? 256 : 128, BN);
    prm.debug = rtc.debug_flags & 11;
    long long *ts_dev = nullptr;
    if (rtc.debug_flags & 4) { CU_CHK(cudaMalloc(&ts_dev, 16 * sizeof(long long))); CU_CHK(cudaMemsetAsync(ts_dev, 0, 16 * sizeof(long long), st)); prm.ts = ts_dev; }
    dim3 const grid((unsigned)round_up(p_tiles, k2 ? 2 : 1), (unsigned)q_tiles, 1);
    mark_kernel_begin();
    if (BN == 128) { launch_taps_bn<128>(planes, k2, grid, stg.smem, a_hi, a_lo, w_hi, w_lo, prm); }
    else if (BN == 96) { launch_taps_bn<96>(planes, k2, grid, stg.smem, a_hi, a_lo, w_hi, w_lo, prm); }
    else if (BN == 64) { launch_taps_bn<64>(planes, k2, grid, stg.smem, a_hi, a_lo, w_hi, w_lo, prm); }
    else { launch_taps_bn<32>(planes, k2, grid, stg.smem, a_hi, a_lo, w_hi, w_lo, prm); }
    mark_kernel_end();
    if (ts_dev) {  // experiments: phase stamps of CTA (0,0) in SM cycles relative to kernel entry
      long long ts[16];
      CU_CHK(cudaStreamSynchronize(st));
      CU_CHK(cudaMemcpy(ts, ts_dev, sizeof(ts), cudaMemcpyDeviceToHost));
      cudaFree(ts_dev);
      fprintf(stderr, "taps stamps (cycles from entry): setup %lld  first_b_full %lld  mma_issue_done %lld  epi_first_chunk %lld  epi_last_chunk %lld  epi_drained %lld  stores_done %lld  end %lld  [steps %d a_stages %d b_stages %d k2 %d]\n",
              ts[1] - ts[0], ts[2] - ts[0], ts[3] - ts[0], ts[4] - ts[0], ts[5] - ts[0], ts[6] - ts[0], ts[7] - ts[0], ts[8] - ts[0], prm.cblks * prm.taps, prm.a_stages, prm.b_stages, (int)k2);
    }
    im.bump(vout);
    return true;
  }

  // experiments (debug_flags bit 4): role stall counters of the persistent CTA-pair kernel, median / max over clusters, in SM cycles
  void print_role_stamps(long long *ts_dev, int n_clusters, int BN, int planes, int kblks) {
    std::vector<long long> ts((size_t)n_clusters * 16);
    CU_CHK(cudaStreamSynchronize(st));
    CU_CHK(cudaMemcpy(ts.data(), ts_dev, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(ts_dev);
    static char const *names[16] = {"prod_total", "prod_wait_empty", "mma_total", "mma_wait_full", "mma_wait_tmem_empty", "mma_first_full", "mma_stages", "mma_wait_afull", "epi_total", "epi_wait_tmem_full", "epi_drain", "epi_store", "epi_tiles", "prod_issue", "", ""};
    string line = "role stamps '" + rfc.rtc_func_name + "' bn=" + str(BN) + " planes=" + str(planes) + " kblks/tile=" + str(kblks) + " clusters=" + str(n_clusters) + " (median/max cycles):";
    for (int k = 0; k < 14; ++k) {
      if (!names[k][0]) { continue; }
      std::vector<long long> v;
      for (int c = 0; c < n_clusters; ++c) { v.push_back(ts[(size_t)c * 16 + k]); }
      std::sort(v.begin(), v.end());
      line += string(" ") + names[k] + "=" + str(v[v.size() / 2]) + "/" + str(v.back());
    }
    fprintf(stderr, "%s\n", line.c_str());
  }

  void run_conv() {
    conv_plan_t const &cp = f.cp;
    var_info_t &vin = var("in"), &vf = var("filts"), &vout = var("out");
    if (!(vin.dims == f.op.get_dims("in")) || !(vf.dims == f.op.get_dims("filts")) || !(vout.dims == f.op.get_dims("out"))) {
      rt_err("conv call '" + rfc.rtc_func_name + "': var dims differ from the dims the function was compiled for");
    }
    float const *bias = nullptr;
    if (has_arg("biases")) { var_info_t &vb = var("biases"); if ((int)vb.dims.dims_prod() != cp.OC) { rt_err("conv: biases size mismatch"); } bias = fptr(vb); }
    bool const bf16 = (rtc.prec == B200_PREC_BF16);
    int const planes = (rtc.prec == B200_PREC_FP32_SPLIT) ? 2 : 1;
    // optional residual input (IgemmParams::res): out = relu?((conv + bias) + res)
    float const *res = nullptr;
    if (has_arg("res")) {
      var_info_t &vr = var("res");
      if (!(vr.dims == vout.dims)) { rt_err("conv: res dims differ from out"); }
      if (cp.swapped || cp.splits != 1 || has_arg("out_concat")) { unsup_err("conv: a residual input needs a pixel-major, un-split launch without out_concat (conv_res_fusable)"); }
      res = fptr(vr);
    }
    // the round-2 kernel (igemm4.cuh: persistent CTA pairs, halo operand mode, stream-K) takes every layer with at least two 128-row tiles
    sk4_plan_t const sp = plan_sk4(cp, planes, im.num_sms, rtc);
    if (!sp.use && cp.taps && rtc.use_taps && !res && !has_arg("out_concat") && run_conv_taps(vin, vf, vout, bias, bf16, planes)) { return; }
    int const BN = sp.use ? sp.BN : cp.BN;
    // filters: OIHW -> K-major rows (k = tap x chan), stored k-block-major [k / 64][OC padded][64] so that every TMA tile is contiguous
    // (once per weight version; the reference's xpose_filts, src/rtc_fwd.cc:310-313)
    long long const oc_pad = round_up(cp.OC, 128);
    // activations: NCHW -> NHWC (chan padded to a multiple of 8; row-merged path: chan padded to 4|8 and image rows at pitch Wp with x padding;
    // halo mode: the shared-padding layout [n][H + py][W + px][chan], pixel (y, x) at (y + py, x + px), see igemm3.cuh / igemm4.cuh)
    packed_t *a_pack_p = nullptr;
    if (cp.rowmerge) {
      pack(f.w_pack, vf, cp.OC, cp.C, cp.KH * cp.KW, cp.Cpad, cp.Cpad, cp.w_row_stride, oc_pad * cp.w_row_stride, planes == 2, bf16, cp.KW, 64, 0, nullptr, 0, oc_pad);
      long long const img_elems = (long long)cp.H * cp.Wp * cp.Cpad;
      pack(f.a_pack, vin, cp.N, cp.C, cp.H * cp.W, cp.Cpad, cp.Cpad, img_elems, (long long)cp.N * img_elems + 64, planes == 2, bf16, cp.W, (long long)cp.Wp * cp.Cpad, (long long)cp.px * cp.Cpad, absmax_cell("in"), cp.W);
      a_pack_p = &f.a_pack;  // row-merged planes are private to this function
    } else {
      pack(f.w_pack, vf, cp.OC, cp.C, cp.KH * cp.KW, cp.Cpad, cp.w_tap_stride, cp.w_row_stride, oc_pad * cp.w_row_stride, planes == 2, bf16, 0, 0, 0, nullptr, 0, oc_pad,
           sp.halo ? cp.KH * cp.KW : 0);  // halo mode: k-blocks channel block major, tap minor (one 3-d TMA box per stage)
      if (sp.halo) {
        long long const img_elems = (long long)sp.Hp * sp.Wp * cp.Cpad;
        a_pack_p = &im.act_packs[{vin.buf->p, b200_impl_t::pad_tag(cp.py, cp.px)}];
        pack(*a_pack_p, vin, cp.N, cp.C, cp.H * cp.W, cp.Cpad, cp.Cpad, img_elems, (long long)cp.N * img_elems, planes == 2, bf16, cp.W, (long long)sp.Wp * cp.Cpad,
             ((long long)cp.py * sp.Wp + cp.px) * cp.Cpad, absmax_cell("in"));
      } else {
        long long const act_elems = (long long)cp.N * cp.H * cp.W * cp.Cpad;
        a_pack_p = &im.act_packs[{vin.buf->p, 0u}];  // plain NHWC planes are shared by every consumer of the node
        pack(*a_pack_p, vin, cp.N, cp.C, cp.H * cp.W, cp.Cpad, cp.Cpad, (long long)cp.H * cp.W * cp.Cpad, act_elems, planes == 2, bf16, 0, 0, 0, absmax_cell("in"));
      }
    }
    packed_t &a_pack = *a_pack_p;
    long long const pixels = (long long)cp.N * cp.OH * cp.OW;
    int const p_rows_n = cp.swapped ? cp.OC : (sp.halo ? (int)sp.m_rows : (int)pixels), q_rows_n = cp.swapped ? (int)pixels : cp.OC;
    int const p_tiles = ceil_div(p_rows_n, b200::IGEMM_BM), q_tiles = ceil_div(q_rows_n, BN);
    // CTA pairs (tcgen05 cta_group::2, igemm2.cuh) whenever the layer has >= 2 row tiles and needs no split-K; else one CTA per tile,
    // optionally with TMA-multicast clusters (use_clusters, off by default: measured slower than plain tiles on AlexNet)
    bool const two_cta = !sp.use && rtc.use_2cta && !cp.swapped && cp.splits == 1 && p_tiles >= 2;
    cluster_choice_t const cl = choose_cluster(p_tiles, q_tiles, BN, planes, rtc.use_clusters && !sp.use && !two_cta && !cp.swapped && cp.splits == 1);
    CUtensorMap act_hi, act_lo, w_hi, w_lo;
    bool const pairs = sp.use || two_cta;
    uint32_t const act_box = cp.swapped ? (pairs ? BN / 2 : BN) : b200::IGEMM_BM / cl.cn, w_box = cp.swapped ? b200::IGEMM_BM : (pairs ? BN / 2 : BN / cl.cm);
    if (sp.halo) {
      act_hi = make_tiled_map(a_pack.hi->p, bf16, cp.Cpad, (uint64_t)cp.N * sp.Hp * sp.Wp, cp.Cpad, sp.a_box_rows);
      act_lo = planes == 2 ? make_tiled_map(a_pack.lo->p, bf16, cp.Cpad, (uint64_t)cp.N * sp.Hp * sp.Wp, cp.Cpad, sp.a_box_rows) : act_hi;
    } else if (cp.im2col) {
      act_hi = make_im2col_map(a_pack.hi->p, bf16, cp, act_box);
      act_lo = planes == 2 ? make_im2col_map(a_pack.lo->p, bf16, cp, act_box) : act_hi;
    } else {
      uint64_t const kext = cp.full_kernel ? (uint64_t)cp.a_row_stride : (uint64_t)cp.Cpad;
      act_hi = make_tiled_map(a_pack.hi->p, bf16, kext, cp.a_rows, cp.a_row_stride, act_box);
      act_lo = planes == 2 ? make_tiled_map(a_pack.lo->p, bf16, kext, cp.a_rows, cp.a_row_stride, act_box) : act_hi;
    }
    uint64_t const w_rows = (uint64_t)(cp.w_row_stride / 64) * oc_pad;
    if (sp.halo) {  // [k-block][out chan (padded)][64]: a stage's ku consecutive k-blocks are one box
      w_hi = make_tiled_map3(f.w_pack.hi->p, bf16, oc_pad, (uint64_t)(cp.w_row_stride / 64), w_box, sp.ku);
      w_lo = planes == 2 ? make_tiled_map3(f.w_pack.lo->p, bf16, oc_pad, (uint64_t)(cp.w_row_stride / 64), w_box, sp.ku) : w_hi;
    } else {
      w_hi = make_tiled_map(f.w_pack.hi->p, bf16, 64, w_rows, 64, w_box);
      w_lo = planes == 2 ? make_tiled_map(f.w_pack.lo->p, bf16, 64, w_rows, 64, w_box) : w_hi;
    }

    b200::IgemmParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.p_rows = p_rows_n;
    prm.q_rows = q_rows_n;
    prm.kblks_total = cp.kblks_total;
    prm.kblks_per_split = cp.kblks_per_split;
    prm.chunk_kblks = std::max(1, planes == 2 ? rtc.acc_chunk_kblks : rtc.acc_chunk_kblks_16);
    prm.p_im2col = cp.im2col ? 1 : 0;
    prm.cblks = std::max(cp.cblks, 1); prm.kw = cp.KW; prm.ow = cp.OW; prm.ohw = cp.OH * cp.OW;
    prm.sx = cp.sx; prm.sy = cp.sy; prm.px = cp.px; prm.py = cp.py;
    if (cp.rowmerge) { prm.kw = 1; prm.sx = 1; prm.px = 0; }  // taps = filter rows; the w coordinate is the output column itself
    prm.swapped = cp.swapped ? 1 : 0;
    prm.out_chans = cp.OC; prm.out_hw = cp.OH * cp.OW;
    prm.relu = cp.relu; prm.has_bias = bias ? 1 : 0; prm.bias = bias;
    prm.p_scale = static_cast<float *>((cp.swapped ? f.w_pack : a_pack).scale2->p);
    prm.q_scale = static_cast<float *>((cp.swapped ? a_pack : f.w_pack).scale2->p);
    prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 128, BN);
    prm.cm = cl.cm; prm.cn = cl.cn;
    prm.kb_mod = cp.kb_mod; prm.ksteps_last = cp.ksteps_last;
    prm.p_kb_rows = cp.swapped ? (int)oc_pad : 0; prm.q_kb_rows = cp.swapped ? 0 : (int)oc_pad;
    prm.debug = rtc.debug_flags;
    prm.out_absmax = absmax_cell("out");
    long long const out_elems = (long long)cp.N * cp.OC * cp.OH * cp.OW;
    // Concat by offset: with an "out_concat" var + by-value "out_ocix" the epilogue writes this convolution's channels
    // straight into the Concat output at channel offset ocix (the per-image stride is the wider var's channel count); `out` itself is then
    // not written (split-K layers: the reduce kernel writes the strided slice).
    float *out_base = fptr(vout);
    var_info_t *vcat = nullptr;
    bool const splitk_pass = !sp.use && cp.splits > 1;  // two-pass split-K of the one-CTA kernel (the round-2 kernel reduces in place: stream-K)
    if (has_arg("out_concat")) {
      vcat = &var("out_concat");
      check_nchw(vcat->dims, "out_concat");
      int const ocix = (int)scalar("out_ocix");
      if ((int)vcat->dims.dsz("img") != cp.N || (int)vcat->dims.dsz("y") != cp.OH || (int)vcat->dims.dsz("x") != cp.OW || ocix + cp.OC > (int)vcat->dims.dsz("chan")) {
        rt_err("conv: out does not fit into out_concat at out_ocix");
      }
      out_base = fptr(*vcat) + (long long)ocix * cp.OH * cp.OW;
      if (!splitk_pass) { prm.out_chans = (int)vcat->dims.dsz("chan"); }  // split-K partials keep this layer's own geometry; the reduce kernel strides
    }
    // "out_pack": also write the NHWC 16-bit plane(s) the consuming convolutions read (shared act_packs entry of the destination var), so
    // they find it fresh and skip their pack kernel. Needs 16-byte aligned runs: channel offset and channel count multiples of 8. By-value
    // "out_pack_py" / "out_pack_px" select the shared-padding layout of a halo-mode consumer (0 / absent = plain pixel-major).
    packed_t *out_pk = nullptr;
    var_info_t &vdst = vcat ? *vcat : vout;
    int o16_py = 0, o16_px = 0;
    bool o16_padded = false;
    prm.res = res;
    prm.res_absmax = res ? absmax_cell("res") : nullptr;
    // (beside a residual the fp16 planes' bound needs max|res|: without that cell only bf16 planes are written; consumers then pack as usual)
    if (has_arg("out_pack") && scalar("out_pack") != 0 && !cp.swapped && !splitk_pass && (cp.OC % 8) == 0 && (bf16 || !vcat) && (bf16 || !res || prm.res_absmax)) {
      int const ocix = vcat ? (int)scalar("out_ocix") : 0;
      int const cdst = (int)vdst.dims.dsz("chan"), cdst_pad = (int)round_up(cdst, 8);
      if (sp.use && has_arg("out_pack_py")) { o16_py = (int)scalar("out_pack_py"); o16_px = (int)scalar("out_pack_px"); o16_padded = true; }  // only the round-2 kernel writes padded planes
      if ((ocix % 8) == 0) {
        out_pk = &im.act_packs[{vdst.buf->p, o16_padded ? b200_impl_t::pad_tag(o16_py, o16_px) : 0u}];
        uint64_t const bytes = (uint64_t)cp.N * (cp.OH + o16_py) * (cp.OW + o16_px) * cdst_pad * 2;
        long long const hw = (long long)cp.OH * cp.OW;
        uint64_t const lkey = o16_padded ? pack_layout_key({cp.N, cdst, hw, cdst_pad, cdst_pad, (long long)(cp.OH + o16_py) * (cp.OW + o16_px) * cdst_pad, cp.OW, (long long)(cp.OW + o16_px) * cdst_pad,
                                                            ((long long)o16_py * (cp.OW + o16_px) + o16_px) * cdst_pad, planes == 2, bf16, 0, 0, 0})
                                         : pack_layout_key({cp.N, cdst, hw, cdst_pad, cdst_pad, hw * cdst_pad, std::max<long long>(hw, 1), 0, 0, planes == 2, bf16, 0, 0, 0});
        if (!out_pk->hi || out_pk->hi->bytes < bytes || (planes == 2 && !out_pk->lo)) {
          out_pk->hi = std::make_shared<dev_buf_t>(bytes);
          CU_CHK(cudaMemsetAsync(out_pk->hi->p, 0, bytes, st));
          if (planes == 2) { out_pk->lo = std::make_shared<dev_buf_t>(bytes); CU_CHK(cudaMemsetAsync(out_pk->lo->p, 0, bytes, st)); }
          out_pk->scale2 = std::make_shared<dev_buf_t>(8);
          out_pk->absmax_bits = std::make_shared<dev_buf_t>(16);
          CU_CHK(cudaMemsetAsync(out_pk->absmax_bits->p, 0, 16, st));
          static float const ones[2] = {1.0f, 1.0f};
          CU_CHK(cudaMemcpyAsync(out_pk->scale2->p, ones, 8, cudaMemcpyHostToDevice, st));
        } else if (out_pk->layout_key != lkey) {  // same storage, other geometry: padding positions must be zero again
          CU_CHK(cudaMemsetAsync(out_pk->hi->p, 0, out_pk->hi->bytes, st));
          if (out_pk->lo) { CU_CHK(cudaMemsetAsync(out_pk->lo->p, 0, out_pk->lo->bytes, st)); }
        }
        out_pk->layout_key = lkey;
        prm.out16 = static_cast<uint16_t *>(out_pk->hi->p) + ocix;
        prm.out16_pitch = cdst_pad;
        if (!bf16) {  // fp16 planes: scale from the output bound (IgemmParams::w_l1max); the filter L1 norm is computed once per weight version
          if (!f.w_l1max) { f.w_l1max = std::make_shared<dev_buf_t>(4); }
          if (f.w_l1max_gen != *vf.gen) {
            CU_CHK(cudaMemsetAsync(f.w_l1max->p, 0, 4, st));
            B200_CARVEOUT_ONCE(b200::filts_l1max_kernel);
            launch_k(b200::filts_l1max_kernel, dim3(cp.OC), dim3(256), 0, fptr(vf), (long long)(vf.dims.dims_prod() / cp.OC), static_cast<unsigned int *>(f.w_l1max->p));
            launched();
            f.w_l1max_gen = *vf.gen;
          }
          prm.w_l1max = static_cast<float *>(f.w_l1max->p);
          prm.out16_lo = planes == 2 ? static_cast<uint16_t *>(out_pk->lo->p) + ocix : nullptr;
          prm.out16_scale2 = static_cast<float *>(out_pk->scale2->p);
          prm.n_bias = cp.OC;
          prm.in_absmax = absmax_cell("in");
        }
      }
    }
    if (splitk_pass) {
      uint64_t const need = (uint64_t)cp.splits * out_elems * 4;
      if (!f.splitk_ws || f.splitk_ws->bytes < need) { f.splitk_ws = std::make_shared<dev_buf_t>(need); }
      prm.out = static_cast<float *>(f.splitk_ws->p);
      prm.split_stride = out_elems;
      if (!vcat && rtc.fuse_splitk_reduce) {  // the last split CTA of every tile reduces it in the kernel (igemm.cuh): no splitk_reduce_kernel launch
        uint64_t const n_tix = (uint64_t)round_up(p_tiles, cl.cm) * round_up(q_tiles, cl.cn) * 8;  // {arrived, reduced} per tile
        if (!f.splitk_tickets || f.splitk_tickets->bytes < n_tix) {
          f.splitk_tickets = std::make_shared<dev_buf_t>(n_tix);
          CU_CHK(cudaMemsetAsync(f.splitk_tickets->p, 0, n_tix, st));  // the counters are re-armed by their last CTA from here on
        }
        prm.tile_tickets = static_cast<unsigned int *>(f.splitk_tickets->p);
        prm.out_final = out_base;
      }
    } else {
      prm.out = out_base;
      prm.split_stride = 0;
    }
    dim3 grid((unsigned)round_up(p_tiles, two_cta ? 2 : cl.cm), (unsigned)round_up(q_tiles, cl.cn), cp.splits);
    long long *ts_dev = nullptr;
    int ts_clusters = 0;
    mark_kernel_begin();
    if (sp.use) {
      b200::Sk4Params sk;
      memset(&sk, 0, sizeof(sk));
      prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 256, BN);
      prm.m_pair_tiles = sp.m_pair_tiles; prm.q_tiles = sp.q_tiles;
      prm.kblks_per_split = prm.kblks_total;
      prm.cm = prm.cn = 1;
      sk.p_mode = sp.halo ? 2 : (cp.im2col ? 1 : 0);
      sk.taps = sp.halo ? cp.KH * cp.KW : 1;
      sk.Wp = sp.Wp; sk.HpWp = sp.Hp * sp.Wp; sk.OH = cp.OH; sk.OW = cp.OW;
      sk.halo_rows = sp.halo_rows; sk.a_loads = sp.a_loads; sk.a_box_rows = sp.a_box_rows;
      sk.a_stages = sp.a_stages; sk.b_stages = sp.b_stages; sk.ku = sp.ku;
      sk.sk = sp.sk ? 1 : 0; sk.n_tiles = sp.n_tiles; sk.ukb = sp.ukb;
      sk.out_w = cp.OW;
      if (o16_padded) { sk.o16_Hp = cp.OH + o16_py; sk.o16_Wp = cp.OW + o16_px; sk.o16_py = o16_py; sk.o16_px = o16_px; }
      if (sp.sk) {
        uint64_t const ws_bytes = (uint64_t)im.num_sms * 128 * 128 * 4;  // one [BN <= 128][128] fp32 slot per CTA
        if (!im.sk_ws) {
          im.sk_ws = std::make_shared<dev_buf_t>(ws_bytes);
          im.sk_flags = std::make_shared<dev_buf_t>((uint64_t)im.num_sms * 4);
          CU_CHK(cudaMemsetAsync(im.sk_flags->p, 0, (uint64_t)im.num_sms * 4, st));  // flags are reset by their consumer from here on
        }
        sk.sk_ws = static_cast<float *>(im.sk_ws->p);
        sk.sk_flags = static_cast<unsigned int *>(im.sk_flags->p);
      }
      if (rtc.debug_flags & 16) { ts_clusters = sp.n_pairs; CU_CHK(cudaMalloc(&ts_dev, (size_t)sp.n_pairs * 16 * sizeof(long long))); CU_CHK(cudaMemsetAsync(ts_dev, 0, (size_t)sp.n_pairs * 16 * sizeof(long long), st)); prm.ts = ts_dev; }
      sk.g = prm;
      if (cp.swapped) { launch_sk4(BN, planes, sp, w_hi, w_lo, act_hi, act_lo, sk); }
      else { launch_sk4(BN, planes, sp, act_hi, act_lo, w_hi, w_lo, sk); }
    }
    else if (two_cta) {  // persistent: one cluster per SM pair (or per tile, if fewer), each walking its share of the tiles
      prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 256, BN);
      prm.m_pair_tiles = ceil_div(p_tiles, 2); prm.q_tiles = q_tiles;
      int const n_clusters = std::min(prm.m_pair_tiles * prm.q_tiles, im.num_sms / 2);
      if (rtc.debug_flags & 16) { ts_clusters = n_clusters; CU_CHK(cudaMalloc(&ts_dev, (size_t)n_clusters * 16 * sizeof(long long))); CU_CHK(cudaMemsetAsync(ts_dev, 0, (size_t)n_clusters * 16 * sizeof(long long), st)); prm.ts = ts_dev; }
      launch_igemm2(BN, planes, dim3(2 * n_clusters, 1, 1), act_hi, act_lo, w_hi, w_lo, prm);
    }
    else if (cp.swapped) { launch_igemm(BN, planes, grid, w_hi, w_lo, act_hi, act_lo, prm); }
    else { launch_igemm(BN, planes, grid, act_hi, act_lo, w_hi, w_lo, prm); }
    mark_kernel_end();
    if (ts_dev) { print_role_stamps(ts_dev, ts_clusters, BN, planes, cp.kblks_total); }
    if (splitk_pass && !prm.tile_tickets) {
      B200_CARVEOUT_ONCE(b200::splitk_reduce_kernel); launch_k(b200::splitk_reduce_kernel, dim3(ceil_div(out_elems, 256)), dim3(256), 0, static_cast<float *>(f.splitk_ws->p), out_base, bias, out_elems, cp.splits, cp.OC, cp.OH * cp.OW, cp.relu, prm.out_absmax,
               vcat ? (long long)vcat->dims.dsz("chan") * cp.OH * cp.OW : 0ll);
      launched();
    }
    im.bump(vdst);
    if (out_pk) { out_pk->src_gen = *vdst.gen; out_pk->src_ptr = vdst.buf->p; }  // the plane is current for this write generation (layout_key set above): consumers skip their pack
  }
  static constexpr uint32_t IGEMM_BM_host() { return b200::IGEMM_BM; }

  // A chain of inner-product-shaped convolutions as one persistent kernel (fcchain.cuh). Arguments: "in" (+ its abs-max cell), and per layer
  // i = 0.. : "filts<i>", "biases<i>", "out<i>" (+ "out<i>_absmax_cells" / "out<i>_absmax_ix"). The layers' plans and filter packs are those of
  // the conv functions named by the chain function's "layers" parameter, so each node is bit-identical to running them one by one.
  template <int kPlanes>
  void launch_fc_chain_t(int grid, b200::FcChainMaps const &maps, b200::FcChainParams const &prm) {
    using Cfg = b200::IgemmCfg<b200::FC_BN, kPlanes>;
    static uint64_t attr_set = 0;
    if (first_use_on_device(attr_set, rtc.device)) {
      CU_CHK(cudaFuncSetAttribute(b200::fc_chain_kernel<kPlanes>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(b200::IGEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rtc.use_pdl ? 1 : 0;
    CU_CHK(cudaLaunchKernelEx(&cfg, b200::fc_chain_kernel<kPlanes>, maps, prm));
    launched();
  }
  void run_fc_chain() {
    bool const bf16 = (rtc.prec == B200_PREC_BF16);
    int const planes = (rtc.prec == B200_PREC_FP32_SPLIT) ? 2 : 1;
    int const nl = (int)f.chain_funcs.size();
    b200::FcChainMaps maps;
    b200::FcChainParams prm;
    memset(&maps, 0, sizeof(maps));
    memset(&prm, 0, sizeof(prm));
    prm.n_layers = nl;
    prm.bf16 = bf16 ? 1 : 0;
    prm.chunk_kblks = std::max(1, planes == 2 ? rtc.acc_chunk_kblks : rtc.acc_chunk_kblks_16);
    prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 128, b200::FC_BN);
    prm.l2_ahead = rtc.fc_l2_ahead; prm.l2_next = rtc.fc_l2_next;
    if (!f.chain_sync) {
      f.chain_sync = std::make_shared<dev_buf_t>(b200::FC_SYNC_WORDS * 4);
      CU_CHK(cudaMemsetAsync(f.chain_sync->p, 0, b200::FC_SYNC_WORDS * 4, st));  // re-armed by the kernel's last CTA from here on
    }
    prm.sync = static_cast<unsigned int *>(f.chain_sync->p);
    var_info_t *vprev = &var("in");
    uint64_t ws_need = 0;
    int grid = 1;
    vector<var_info_t *> outs;
    for (int i = 0; i < nl; ++i) {
      func_t &lf = im.funcs.at(f.chain_funcs[i]);
      conv_plan_t const &cp = lf.cp;
      string const si = str(i);
      var_info_t &vf = var("filts" + si), &vout = var("out" + si);
      if (!(vprev->dims == lf.op.get_dims("in")) || !(vf.dims == lf.op.get_dims("filts")) || !(vout.dims == lf.op.get_dims("out"))) {
        rt_err("fc_chain call '" + rfc.rtc_func_name + "': layer " + si + " var dims differ from the dims '" + f.chain_funcs[i] + "' was compiled for");
      }
      long long const K = (long long)cp.KH * cp.KW * cp.Cpad;
      if ((cp.OC % 4) != 0 || cp.splits > 16) { unsup_err("fc_chain: layer " + si + " needs out chans % 4 == 0 and at most 16 splits"); }
      if (i > 0 && (cp.C != (int)outs.back()->dims.dsz("chan") || cp.C != cp.Cpad || cp.KH * cp.KW != 1 || (K % 64) != 0)) { unsup_err("fc_chain: layer " + si + " does not read the previous layer's output as a 64-multiple K"); }
      b200::FcLayer &L = prm.L[i];
      long long const oc_pad = round_up(cp.OC, 128);
      // filters: as run_conv packs them (once per weight version)
      pack(lf.w_pack, vf, cp.OC, cp.C, cp.KH * cp.KW, cp.Cpad, cp.w_tap_stride, cp.w_row_stride, oc_pad * cp.w_row_stride, planes == 2, bf16, 0, 0, 0, nullptr, 0, oc_pad, 0);
      uint64_t const w_rows = (uint64_t)(cp.w_row_stride / 64) * oc_pad;
      maps.w_hi[i] = make_tiled_map(lf.w_pack.hi->p, bf16, 64, w_rows, 64, b200::IGEMM_BM);
      maps.w_lo[i] = planes == 2 ? make_tiled_map(lf.w_pack.lo->p, bf16, 64, w_rows, 64, b200::IGEMM_BM) : maps.w_hi[i];
      if (i == 0) {  // the first layer's activation planes: the node's shared NHWC planes (written by its producer, or packed here)
        long long const act_elems = (long long)cp.N * cp.H * cp.W * cp.Cpad;
        packed_t &a_pack = im.act_packs[{vprev->buf->p, 0u}];
        pack(a_pack, *vprev, cp.N, cp.C, cp.H * cp.W, cp.Cpad, cp.Cpad, (long long)cp.H * cp.W * cp.Cpad, act_elems, planes == 2, bf16, 0, 0, 0, absmax_cell("in"));
        uint64_t const kext = cp.full_kernel ? (uint64_t)cp.a_row_stride : (uint64_t)cp.Cpad;
        maps.a_hi[0] = make_tiled_map(a_pack.hi->p, bf16, kext, cp.a_rows, cp.a_row_stride, b200::FC_BN);
        maps.a_lo[0] = planes == 2 ? make_tiled_map(a_pack.lo->p, bf16, kext, cp.a_rows, cp.a_row_stride, b200::FC_BN) : maps.a_hi[0];
        prm.a0_scale2 = static_cast<float *>(a_pack.scale2->p);
        prm.batch = cp.N;
      } else {  // planes between the layers belong to the chain: [32][K], rows beyond the batch stay zero
        packed_t &pk = f.chain_planes[i];
        uint64_t const bytes = (uint64_t)b200::FC_BN * K * 2;
        if (!pk.hi || pk.hi->bytes < bytes || (planes == 2 && !pk.lo)) {
          pk.hi = std::make_shared<dev_buf_t>(bytes);
          CU_CHK(cudaMemsetAsync(pk.hi->p, 0, bytes, st));
          if (planes == 2) { pk.lo = std::make_shared<dev_buf_t>(bytes); CU_CHK(cudaMemsetAsync(pk.lo->p, 0, bytes, st)); }
          pk.scale2 = std::make_shared<dev_buf_t>(8);
        }
        maps.a_hi[i] = make_tiled_map(pk.hi->p, bf16, (uint64_t)K, b200::FC_BN, (uint64_t)K, b200::FC_BN);
        maps.a_lo[i] = planes == 2 ? make_tiled_map(pk.lo->p, bf16, (uint64_t)K, b200::FC_BN, (uint64_t)K, b200::FC_BN) : maps.a_hi[i];
        b200::FcLayer &P = prm.L[i - 1];
        P.nxt_hi = static_cast<uint16_t *>(pk.hi->p);
        P.nxt_lo = planes == 2 ? static_cast<uint16_t *>(pk.lo->p) : nullptr;
        P.nxt_scale2 = static_cast<float *>(pk.scale2->p);
        if (cp.N != prm.batch) { rt_err("fc_chain: batch changes along the chain"); }
      }
      L.n_out = cp.OC; L.oc_pad = (int)oc_pad;
      L.tiles = ceil_div(cp.OC, b200::IGEMM_BM); L.splits = cp.splits; L.kblks_total = cp.kblks_total; L.kblks_per_split = cp.kblks_per_split;
      L.kb_mod = cp.kb_mod; L.ksteps_last = cp.ksteps_last;
      L.relu = cp.relu;
      L.has_bias = has_arg("biases" + si) ? 1 : 0;
      if (L.has_bias) { var_info_t &vb = var("biases" + si); if ((int)vb.dims.dims_prod() != cp.OC) { rt_err("fc_chain: biases size mismatch"); } L.bias = fptr(vb); }
      L.out = fptr(vout);
      L.w_scale2 = static_cast<float *>(lf.w_pack.scale2->p);
      unsigned int *cell = absmax_cell("out" + si);
      L.out_absmax = cell ? cell : prm.sync + 16 + i;
      ws_need = std::max<uint64_t>(ws_need, (uint64_t)cp.splits * cp.N * cp.OC * 4);
      grid = std::max(grid, L.tiles * L.splits);
      outs.push_back(&vout);
      vprev = &vout;
    }
    if (grid > im.num_sms) { unsup_err("fc_chain: more units than SMs"); }
    if (im.tail_gather_on && rfc.arg_map.at("out" + str(nl - 1)).is_var() && rfc.arg_map.at("out" + str(nl - 1)).get_var() == im.tail_gather_vn) {
      b200_gather_desc_t const &d = im.tail_gather;
      if (d.bytes_per_rank != (uint64_t)prm.batch * prm.L[nl - 1].n_out * 4) { rt_err("fc_chain: the gather buffer's bytes per rank differ from the size of '" + im.tail_gather_vn + "'"); }
      for (int r = 0; r < d.world; ++r) { prm.g.peer_base[r] = d.peer_base[r]; }
      prm.g.local_base = d.local_base; prm.g.bytes_per_rank = d.bytes_per_rank; prm.g.flag_bytes = d.flag_bytes; prm.g.rank = d.rank; prm.g.world = d.world;
    }
    if (!f.splitk_ws || f.splitk_ws->bytes < ws_need) { f.splitk_ws = std::make_shared<dev_buf_t>(ws_need); }
    prm.ws = static_cast<float *>(f.splitk_ws->p);
    long long *ts_dev = nullptr;
    if (rtc.debug_flags & 16) { CU_CHK(cudaMalloc(&ts_dev, (size_t)grid * 32 * sizeof(long long))); CU_CHK(cudaMemsetAsync(ts_dev, 0, (size_t)grid * 32 * sizeof(long long), st)); prm.ts = ts_dev; }
    mark_kernel_begin();
    if (planes == 2) { launch_fc_chain_t<2>(grid, maps, prm); } else { launch_fc_chain_t<1>(grid, maps, prm); }
    mark_kernel_end();
    if (ts_dev) {  // per event: min / median / max over the CTAs, in us since the first CTA started
      std::vector<long long> ts((size_t)grid * 32);
      CU_CHK(cudaStreamSynchronize(st));
      CU_CHK(cudaMemcpy(ts.data(), ts_dev, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      cudaFree(ts_dev);
      long long t0 = LLONG_MAX;
      for (int c = 0; c < grid; ++c) { if (ts[(size_t)c * 32 + 31]) { t0 = std::min(t0, ts[(size_t)c * 32 + 31]); } }
      static char const *names[8] = {"w_issued", "planes_ready", "drained", "barA", "reduced", "barB", "packed", ""};
      string line = "fc_chain stamps '" + rfc.rtc_func_name + "' grid=" + str(grid) + " (us since start, min/median/max over CTAs):";
      for (int l = 0; l < nl; ++l) {
        line += "\n  layer " + str(l) + ":";
        for (int k = 0; k < 7; ++k) {
          std::vector<long long> v;
          for (int c = 0; c < grid; ++c) { long long const t = ts[(size_t)c * 32 + 8 * l + k]; if (t) { v.push_back(t - t0); } }
          if (v.empty()) { continue; }
          std::sort(v.begin(), v.end());
          char buf[96];
          snprintf(buf, sizeof(buf), " %s=%.1f/%.1f/%.1f", names[k], v.front() * 1e-3, v[v.size() / 2] * 1e-3, v.back() * 1e-3);
          line += buf;
        }
      }
      fprintf(stderr, "%s\n", line.c_str());
    }
    for (auto *v : outs) { im.bump(*v); }
  }

  // c[M,N] = a[K,M]^T b[K,N]  (test/rtc/sgemm.cucl:1-3). P = a^T rows (M), Q = b^T rows (N), both packed K-major.
  void run_sgemm() {
    var_info_t &va = var("a"), &vb = var("b"), &vc = var("c");
    if (va.dims.size() != 2 || vb.dims.size() != 2 || vc.dims.size() != 2) { unsup_err("sgemm expects 2-d a (K:M), b (K:N), c (M:N)"); }
    int const K = va.dims.dsz("K"), M = va.dims.dsz("M"), N = vb.dims.dsz("N");
    if ((int)vb.dims.dsz("K") != K || (int)vc.dims.dsz("M") != M || (int)vc.dims.dsz("N") != N) { rt_err("sgemm: inconsistent dims"); }
    bool const bf16 = (rtc.prec == B200_PREC_BF16);
    int const planes = (rtc.prec == B200_PREC_FP32_SPLIT) ? 2 : 1;
    long long const Kpad = round_up(K, 64);
    pack(f.a_pack, va, 1, K, M, (int)Kpad, Kpad, 0, (long long)M * Kpad, planes == 2, bf16);
    pack(f.w_pack, vb, 1, K, N, (int)Kpad, Kpad, 0, (long long)N * Kpad, planes == 2, bf16);
    int const BN = N > 64 ? 128 : (N > 32 ? 64 : 32);
    int const p_tiles = ceil_div(M, b200::IGEMM_BM), q_tiles = ceil_div(N, BN);
    bool const two_cta = rtc.use_2cta && p_tiles >= 2;
    cluster_choice_t const cl = choose_cluster(p_tiles, q_tiles, BN, planes, rtc.use_clusters && !two_cta);
    uint32_t const a_box = b200::IGEMM_BM / cl.cn, b_box = two_cta ? BN / 2 : BN / cl.cm;
    CUtensorMap a_hi = make_tiled_map(f.a_pack.hi->p, bf16, Kpad, M, Kpad, a_box);
    CUtensorMap a_lo = planes == 2 ? make_tiled_map(f.a_pack.lo->p, bf16, Kpad, M, Kpad, a_box) : a_hi;
    CUtensorMap b_hi = make_tiled_map(f.w_pack.hi->p, bf16, Kpad, N, Kpad, b_box);
    CUtensorMap b_lo = planes == 2 ? make_tiled_map(f.w_pack.lo->p, bf16, Kpad, N, Kpad, b_box) : b_hi;
    b200::IgemmParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.p_rows = M; prm.q_rows = N;
    prm.kblks_total = prm.kblks_per_split = (int)(Kpad / 64);
    prm.chunk_kblks = std::max(1, planes == 2 ? rtc.acc_chunk_kblks : rtc.acc_chunk_kblks_16);
    prm.cblks = 1; prm.kw = 1; prm.ow = 1; prm.ohw = 1; prm.sx = prm.sy = 1;
    prm.out_chans = N; prm.out_hw = 1;
    prm.out = fptr(vc);
    prm.p_scale = static_cast<float *>(f.a_pack.scale2->p);
    prm.q_scale = static_cast<float *>(f.w_pack.scale2->p);
    prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 128, BN);
    prm.cm = cl.cm; prm.cn = cl.cn;
    prm.kb_mod = (int)(Kpad / 64); prm.ksteps_last = ceil_div(K - (Kpad / 64 - 1) * 64, 16);
    prm.debug = rtc.debug_flags;
    dim3 grid((unsigned)round_up(p_tiles, two_cta ? 2 : cl.cm), (unsigned)round_up(q_tiles, cl.cn), 1);
    mark_kernel_begin();
    if (two_cta) {
      prm.idesc = b200::make_idesc_f16(bf16 ? 1u : 0u, 256, BN);
      prm.m_pair_tiles = ceil_div(p_tiles, 2); prm.q_tiles = q_tiles;
      int const n_clusters = std::min(prm.m_pair_tiles * prm.q_tiles, im.num_sms / 2);
      launch_igemm2(BN, planes, dim3(2 * n_clusters, 1, 1), a_hi, a_lo, b_hi, b_lo, prm);
    }
    else { launch_igemm(BN, planes, grid, a_hi, a_lo, b_hi, b_lo, prm); }
    mark_kernel_end();
    im.bump(vc);
  }

  // the consumer's NHWC planes as a second output of a pooling kernel: allocate / re-zero them in the layout the consumer reads (by-value
  // "out_pack_py" / "out_pack_px": the shared-padding layout of a halo-mode convolution, igemm4.cuh) and fill in the kernel's PoolPlanes argument
  packed_t *pool_out_planes(b200::PoolPlanes &pp, var_info_t &vout, int C, int OH, int OW, int npl, bool bf16, unsigned int *in_cell) {
    int o16_py = 0, o16_px = 0;
    bool o16_padded = false;
    int const cpad = (int)round_up(C, 8);
    long long const nimg = (long long)vout.dims.dsz("img");
    if (has_arg("out_pack_py")) { o16_py = (int)scalar("out_pack_py"); o16_px = (int)scalar("out_pack_px"); o16_padded = true; }
    packed_t *out_pk = &im.act_packs[{vout.buf->p, o16_padded ? b200_impl_t::pad_tag(o16_py, o16_px) : 0u}];
    uint64_t const bytes = (uint64_t)nimg * (OH + o16_py) * (OW + o16_px) * cpad * 2;
    long long const ohw2 = (long long)OH * OW;
    uint64_t const lkey = o16_padded ? pack_layout_key({nimg, C, ohw2, cpad, cpad, (long long)(OH + o16_py) * (OW + o16_px) * cpad, OW, (long long)(OW + o16_px) * cpad,
                                                        ((long long)o16_py * (OW + o16_px) + o16_px) * cpad, npl == 2, bf16, 0, 0, 0})
                                     : pack_layout_key({nimg, C, ohw2, cpad, cpad, ohw2 * cpad, std::max<long long>(ohw2, 1), 0, 0, npl == 2, bf16, 0, 0, 0});
    if (!out_pk->hi || out_pk->hi->bytes < bytes || (npl == 2 && !out_pk->lo)) {
      out_pk->hi = std::make_shared<dev_buf_t>(bytes);
      CU_CHK(cudaMemsetAsync(out_pk->hi->p, 0, bytes, st));
      if (npl == 2) { out_pk->lo = std::make_shared<dev_buf_t>(bytes); CU_CHK(cudaMemsetAsync(out_pk->lo->p, 0, bytes, st)); }
      out_pk->scale2 = std::make_shared<dev_buf_t>(8);
      out_pk->absmax_bits = std::make_shared<dev_buf_t>(16);
      CU_CHK(cudaMemsetAsync(out_pk->absmax_bits->p, 0, 16, st));
      if (bf16) { static float const ones[2] = {1.0f, 1.0f}; CU_CHK(cudaMemcpyAsync(out_pk->scale2->p, ones, 8, cudaMemcpyHostToDevice, st)); }
    } else if (out_pk->layout_key != lkey) {  // same storage, other geometry: padding positions must be zero again
      CU_CHK(cudaMemsetAsync(out_pk->hi->p, 0, out_pk->hi->bytes, st));
      if (out_pk->lo) { CU_CHK(cudaMemsetAsync(out_pk->lo->p, 0, out_pk->lo->bytes, st)); }
    }
    out_pk->layout_key = lkey;
    pp.hi = static_cast<uint16_t *>(out_pk->hi->p);
    pp.lo = npl == 2 ? static_cast<uint16_t *>(out_pk->lo->p) : nullptr;
    pp.scale2 = static_cast<float *>(out_pk->scale2->p);
    pp.in_absmax = in_cell;
    pp.C = C; pp.cpad = cpad; pp.bf16 = bf16 ? 1 : 0;
    pp.OW = OW;
    if (o16_padded) { pp.dHp = OH + o16_py; pp.dWp = OW + o16_px; pp.dpy = o16_py; pp.dpx = o16_px; }
    return out_pk;
  }

  void run_pool() {
    var_info_t &vin = var("in"), &vout = var("out");
    check_nchw(vin.dims, "in"); check_nchw(vout.dims, "out");
    int const H = vin.dims.dsz("y"), W = vin.dims.dsz("x"), OH = vout.dims.dsz("y"), OW = vout.dims.dsz("x");
    int KH, KW, sy = 1, sx = 1, py = 0, px = 0;
    if (f.op.has("kern_sz")) {
      KH = f.op.yx("kern_sz", "y", 1); KW = f.op.yx("kern_sz", "x", 1);
      sy = f.op.yx("stride", "y", 1); sx = f.op.yx("stride", "x", 1); py = f.op.yx("in_pad", "y", 0); px = f.op.yx("in_pad", "x", 0);
      int const eoh = (H + 2 * py < KH) ? 1 : ceil_div(H + 2 * py - KH, sy) + 1, eow = (W + 2 * px < KW) ? 1 : ceil_div(W + 2 * px - KW, sx) + 1;
      if (eoh != OH || eow != OW) { rt_err("pool: out dims " + vout.dims.pretty() + " do not match the Caffe ceil rule (" + str(eoh) + "x" + str(eow) + ")"); }
    } else {  // global pooling (src/cnn_op.cc:39-45)
      KH = H; KW = W;
      if (OH != 1 || OW != 1) { rt_err("global pooling must produce 1x1 output"); }
    }
    if (vin.dims.dsz("img") != vout.dims.dsz("img") || vin.dims.dsz("chan") != vout.dims.dsz("chan")) { rt_err("pool: img/chan mismatch"); }
    if (scalar("emit_out_in_yx", true, 0) != 0) { unsup_err("pool: emit_out_in_yx (training only) is out of scope for be=b200"); }
    long long const n_out = vout.dims.dims_prod();
    long long const planes = (long long)vin.dims.dsz("img") * vin.dims.dsz("chan");
    int const avg = (int)scalar("avg_pool", true, 0);
    // LRN in front of the pool, fused (by-value "lrn_local_size" + "lrn_alpha" / "lrn_beta" / "lrn_k"; `in` is then the LRN's INPUT):
    // lrn_maxpool_kernel. The whole-net driver asks for it only for a 3x3 / 2 max pool without padding behind a local_size-5 LRN.
    if (has_arg("lrn_local_size")) {
      int const ls = (int)scalar("lrn_local_size");
      if (ls != 5 || avg || KH != 3 || KW != 3 || sy != 2 || sx != 2 || py != 0 || px != 0) { unsup_err("pool: the fused LRN form handles local_size 5 in front of a 3x3 / 2 max pool without padding"); }
      int const C = (int)vin.dims.dsz("chan"), N = (int)vin.dims.dsz("img");
      bool const bf16 = (rtc.prec == B200_PREC_BF16);
      int const npl = (rtc.prec == B200_PREC_FP32_SPLIT) ? 2 : 1;
      float const alpha = (float)scalar("lrn_alpha", true, 1.0), beta = (float)scalar("lrn_beta", true, 0.75), kk = (float)scalar("lrn_k", true, 1.0);
      unsigned int *in_cell = bf16 ? nullptr : absmax_cell("in");
      bool const want_planes = has_arg("out_pack") && scalar("out_pack") != 0 && (C % 8) == 0 && (bf16 || (in_cell && kk >= 1.0f));  // (k >= 1: the LRN factor is <= 1, max|in| bounds the output)
      b200::PoolPlanes pp;
      memset(&pp, 0, sizeof(pp));
      packed_t *out_pk = nullptr;
      if (want_planes) { out_pk = pool_out_planes(pp, vout, C, OH, OW, npl, bf16, in_cell); }
      int const row_groups = ceil_div(OH, b200::kLpRows), chunks = ceil_div(C, b200::kLpCC);
      size_t const smem = ((size_t)b200::kLpCC * (2 * b200::kLpRows + 1) * W + (size_t)b200::kLpCC * b200::kLpRows * OW) * 4;
      if (smem > 200 * 1024 || N > 65535 || chunks > 65535) { unsup_err("pool: map too wide for the fused LRN form"); }
      // (128 threads at 64 registers, up to six CTAs per SM: measured best of {512x2, 512x1, 256x4, 256x2, 128x4, 128x6, 128x8, 64x8, 64x12}, r02)
      static uint64_t attr_ = 0;
      if (first_use_on_device(attr_, rtc.device)) {
        CU_CHK(cudaFuncSetAttribute(b200::lrn_maxpool_kernel<128, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); prefer_max_smem(b200::lrn_maxpool_kernel<128, 6>);
      }
      launch_k(b200::lrn_maxpool_kernel<128, 6>, dim3(row_groups, chunks, N), dim3(128), smem, fptr(vin), fptr(vout), C, H, W, OH, OW, alpha / 5.0f, -beta, kk, absmax_cell("out"), pp);
      launched();
      im.bump(vout);
      if (out_pk) { out_pk->src_gen = *vout.gen; out_pk->src_ptr = vout.buf->p; }
      return;
    }
    // planes that fit shared memory are staged there (pool_plane_kernel): the 3x3 / 2x2 max pools of AlexNet / NiN / GoogLeNet / ResNet incl.
    // their 112x112 first pools (one 49 KB plane per CTA), GoogLeNet's 5x5 stride-3 average pools, 7x7 (global) average pools
    bool const plane_ks = KH == KW && sy == sx && ((KH == 3 && sy == 2) || (KH == 3 && sy == 1) || (KH == 2 && sy == 2) || (KH == 5 && sy == 3) || (KH == 7 && sy == 1));
    if (plane_ks && (long long)H * W <= 16384 && planes < (1ll << 31)) {
      int const C = (int)vin.dims.dsz("chan");
      bool const bf16 = (rtc.prec == B200_PREC_BF16);
      int const npl = (rtc.prec == B200_PREC_FP32_SPLIT) ? 2 : 1;
      long long const hw = (long long)H * W, ohw = (long long)OH * OW;
      unsigned int *in_cell = bf16 ? nullptr : absmax_cell("in");
      bool const want_planes = has_arg("out_pack") && scalar("out_pack") != 0 && (C % 8) == 0 && (bf16 || in_cell);
      // Persistent double-buffered kernel (pool_plane_pipe_kernel): groups of ~24 KB fetched by bulk copies, which need 16-byte multiples: the
      // group size is a multiple of q planes. With "out_pack" a group is a multiple of 8 channels of ONE image (16-byte NHWC runs).
      int pipe_ppc = 0, ppc = 0;
      size_t pipe_smem = 0;
      bool planes_ok = false;
      {
        int const q = (hw % 4 == 0) ? 1 : (hw % 2 == 0) ? 2 : 4;
        long long const target = std::max<long long>(1, (24 * 1024) / (hw * 4));
        auto smem_of = [&](long long ppc_, bool pl) { return (size_t)(2 * round_up(ppc_ * hw, 32) * 4 + (pl ? ppc_ * ohw * 4 : 0)); };
        auto bulk_ok = [&](long long c) {
          long long const tail = planes % c;
          return (reinterpret_cast<uintptr_t>(fptr(vin)) & 15) == 0 && ((c * hw) % 4) == 0 && ((tail * hw) % 4) == 0 && c * hw * 4 < (1ll << 20);
        };
        // classic form (one group per CTA, loads by all threads): planes per CTA a multiple of 4 (16-byte aligned runs), tile <= 48 KB, and
        // at least ~4 CTAs per SM left to fill the chip
        int cl_ppc = (int)std::max<long long>(4, std::min<long long>((48 * 1024) / (hw * 4) / 4 * 4, round_up(ceil_div(planes, 4 * im.num_sms), 4)));
        if ((long long)cl_ppc * hw * 4 > 96 * 1024) { cl_ppc = 1; }
        int cl_ppc8 = 0;
        if (want_planes) {
          int p8 = (int)round_up(cl_ppc, 8);
          while (p8 > 8 && ((C % p8) != 0 || (long long)p8 * (hw + ohw) * 4 > 100 * 1024)) { p8 -= 8; }
          if ((C % p8) == 0 && (long long)p8 * (hw + ohw) * 4 <= 100 * 1024) { cl_ppc8 = p8; }
        }
        long long pipe8 = 0, pipe1 = std::max<long long>(q, target / q * q);
        if (want_planes) {
          for (long long c8 = std::max<long long>(8, target / 8 * 8); c8 >= 8; c8 -= 8) { if ((C % c8) == 0 && smem_of(c8, true) <= 100 * 1024 && bulk_ok(c8)) { pipe8 = c8; break; } }
        }
        if (smem_of(pipe1, false) > 100 * 1024 || !bulk_ok(pipe1)) { pipe1 = 0; }
        // preference: writing the consumer's planes (saves a pack kernel) first, the pipelined form second
        if (pipe8) { pipe_ppc = ppc = (int)pipe8; pipe_smem = smem_of(pipe8, true); planes_ok = true; }
        else if (cl_ppc8) { ppc = cl_ppc8; planes_ok = true; }
        else if (pipe1) { pipe_ppc = ppc = (int)pipe1; pipe_smem = smem_of(pipe1, false); }
        else { ppc = cl_ppc; }
      }
      b200::PoolPlanes pp;
      memset(&pp, 0, sizeof(pp));
      packed_t *out_pk = nullptr;
      if (planes_ok) { out_pk = pool_out_planes(pp, vout, C, OH, OW, npl, bf16, in_cell); }
      size_t const smem = pipe_ppc ? pipe_smem : (size_t)ppc * (H * W + (pp.hi ? OH * OW : 0)) * 4;
      unsigned int *cell = absmax_cell("out");
      long long const n_groups = ceil_div(planes, ppc);
      // persistent grid: as many CTAs as fit the chip at this shared-memory footprint, each walking its share of the groups
      unsigned const pipe_grid = (unsigned)std::min<long long>(n_groups, (long long)im.num_sms * std::max<long long>(1, std::min<long long>(8, (224 * 1024) / (long long)(smem + 1024))));
#define B200_POOL_PLANE(K_, S_) do { \
        static uint64_t attr_ = 0; \
        if (first_use_on_device(attr_, rtc.device)) { \
          CU_CHK(cudaFuncSetAttribute(b200::pool_plane_kernel<K_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); prefer_max_smem(b200::pool_plane_kernel<K_, S_>); \
          CU_CHK(cudaFuncSetAttribute(b200::pool_plane_pipe_kernel<K_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); prefer_max_smem(b200::pool_plane_pipe_kernel<K_, S_>); \
          } \
        if (pipe_ppc) { launch_k(b200::pool_plane_pipe_kernel<K_, S_>, dim3(pipe_grid), dim3(256), smem, fptr(vin), fptr(vout), H, W, OH, OW, py, px, avg, cell, ppc, planes, pp); } \
        else { launch_k(b200::pool_plane_kernel<K_, S_>, dim3((unsigned)n_groups), dim3(256), smem, fptr(vin), fptr(vout), H, W, OH, OW, py, px, avg, cell, ppc, planes, pp); } } while (0)
      if (KH == 3 && sy == 2) { B200_POOL_PLANE(3, 2); } else if (KH == 3) { B200_POOL_PLANE(3, 1); } else if (KH == 2) { B200_POOL_PLANE(2, 2); }
      else if (KH == 5) { B200_POOL_PLANE(5, 3); } else { B200_POOL_PLANE(7, 1); }
#undef B200_POOL_PLANE
      launched();
      im.bump(vout);
      if (out_pk) { out_pk->src_gen = *vout.gen; out_pk->src_ptr = vout.buf->p; }  // the planes are current (layout_key set above): the consuming convolution skips its pack
      return;
    }
    if (KH == KW && sy == sx && ((KH == 3 && sy == 2) || (KH == 3 && sy == 1) || (KH == 2 && sy == 2)) && planes <= 65535) {
      dim3 const grid(ceil_div((long long)OH * OW, 256), (unsigned)planes, 1);
      unsigned int *cell = absmax_cell("out");
      if (KH == 3 && sy == 2) { B200_CARVEOUT_ONCE(b200::pool_kernel_fixed<3, 2>); launch_k(b200::pool_kernel_fixed<3, 2>, dim3(grid), dim3(256), 0, fptr(vin), fptr(vout), H, W, OH, OW, py, px, avg, cell); }
      else if (KH == 3) { B200_CARVEOUT_ONCE(b200::pool_kernel_fixed<3, 1>); launch_k(b200::pool_kernel_fixed<3, 1>, dim3(grid), dim3(256), 0, fptr(vin), fptr(vout), H, W, OH, OW, py, px, avg, cell); }
      else { B200_CARVEOUT_ONCE(b200::pool_kernel_fixed<2, 2>); launch_k(b200::pool_kernel_fixed<2, 2>, dim3(grid), dim3(256), 0, fptr(vin), fptr(vout), H, W, OH, OW, py, px, avg, cell); }
      launched();
      im.bump(vout);
      return;
    }
    B200_CARVEOUT_ONCE(b200::pool_kernel); launch_k(b200::pool_kernel, dim3(ceil_div(n_out, 256)), dim3(256), 0, fptr(vin), fptr(vout), n_out, H, W, OH, OW, KH, KW, sy, sx, py, px, (int)scalar("avg_pool", true, 0), absmax_cell("out"));
    launched();
    im.bump(vout);
  }

  void run_lrn() {
    var_info_t &vin = var("in"), &vout = var("out");
    check_nchw(vin.dims, "in");
    if (!(vin.dims == vout.dims)) { rt_err("lrn: in/out dims differ"); }
    if (scalar("emit_out_scale_base", true, 0) != 0) { unsup_err("lrn: emit_out_scale_base (training only) is out of scope for be=b200"); }
    int const C = vin.dims.dsz("chan"), HW = vin.dims.dsz("y") * vin.dims.dsz("x");
    long long const n_pels = (long long)vin.dims.dsz("img") * HW;
    int const ls = (int)scalar("local_size", true, 5);
    float const alpha = (float)scalar("alpha", true, 1.0), beta = (float)scalar("beta", true, 0.75), k = (float)scalar("k", true, 1.0);
    int const blocks = ceil_div(n_pels, 128);
    constexpr int kChunk = 16;
    dim3 const grid(blocks, ceil_div(C, kChunk));
    if (ls == 5) { B200_CARVEOUT_ONCE(b200::lrn_kernel<5, kChunk>); launch_k(b200::lrn_kernel<5, kChunk>, dim3(grid), dim3(128), 0, fptr(vin), fptr(vout), n_pels, C, HW, alpha, beta, k, absmax_cell("out")); }
    else if (ls == 3) { B200_CARVEOUT_ONCE(b200::lrn_kernel<3, kChunk>); launch_k(b200::lrn_kernel<3, kChunk>, dim3(grid), dim3(128), 0, fptr(vin), fptr(vout), n_pels, C, HW, alpha, beta, k, absmax_cell("out")); }
    else if (ls <= 32) { B200_CARVEOUT_ONCE(b200::lrn_kernel_generic); launch_k(b200::lrn_kernel_generic, dim3(blocks), dim3(128), 0, fptr(vin), fptr(vout), n_pels, C, HW, ls, alpha, beta, k, absmax_cell("out")); }
    else { unsup_err("lrn: local_size > 32"); }
    launched();
    im.bump(vout);
  }

  void run_relu() {
    var_info_t &v = var("inout");
    long long const n = v.dims.dims_prod();
    B200_CARVEOUT_ONCE(b200::relu_kernel); launch_k(b200::relu_kernel, dim3(ceil_div(ceil_div(n, 4), 256)), dim3(256), 0, fptr(v), n);
    launched();
    im.bump(v);
  }

  void run_softmax() {
    var_info_t &vin = var("in"), &vout = var("prob");
    check_nchw(vin.dims, "in");
    if (!(vin.dims == vout.dims)) { rt_err("softmax: in/prob dims differ"); }
    int const C = vin.dims.dsz("chan"), HW = vin.dims.dsz("y") * vin.dims.dsz("x");
    long long const n_pels = (long long)vin.dims.dsz("img") * HW;
    B200_CARVEOUT_ONCE(b200::softmax_kernel); launch_k(b200::softmax_kernel, dim3(ceil_div(n_pels * 32, 256)), dim3(256), 0, fptr(vin), fptr(vout), n_pels, C, HW);
    launched();
    im.bump(vout);
  }

  void run_copy() {  // Concat piece: out[:, ocix:ocix+C] = in  (src/rtc_fwd.cc:267-280)
    var_info_t &vin = var("in"), &vout = var("out");
    check_nchw(vin.dims, "in"); check_nchw(vout.dims, "out");
    int const ocix = (int)scalar("ocix");
    int const reverse = (int)scalar("reverse", true, 0);  // 1: in = out[:, ocix:ocix+C] (read a Concat input back out of the Concat output)
    int const C = vin.dims.dsz("chan"), HW = vin.dims.dsz("y") * vin.dims.dsz("x"), OC = vout.dims.dsz("chan"), n_img = vin.dims.dsz("img");
    if (vout.dims.dsz("y") != vin.dims.dsz("y") || vout.dims.dsz("x") != vin.dims.dsz("x") || (int)vout.dims.dsz("img") != n_img || ocix + C > OC) { rt_err("copy: in does not fit into out at ocix"); }
    long long const per_img = (long long)C * HW, out_img_stride = (long long)OC * HW, out_off = (long long)ocix * HW;
    int const vec4 = ((per_img % 4) == 0 && (out_img_stride % 4) == 0 && (out_off % 4) == 0) ? 1 : 0;
    long long const work = vec4 ? per_img * n_img / 4 : per_img * n_img;
    B200_CARVEOUT_ONCE(b200::concat_copy_kernel); launch_k(b200::concat_copy_kernel, dim3(ceil_div(work, 256)), dim3(256), 0, fptr(vin), fptr(vout), per_img, out_img_stride, out_off, n_img, vec4, reverse ? nullptr : absmax_cell("out"), reverse);
    launched();
    im.bump(reverse ? vin : vout);
  }

  void run_reduce() {
    var_info_t &vout = var("out");
    int const ins_num = (int)scalar("ins_num");
    if (ins_num < 1 || ins_num > 8) { unsup_err("reduce: 1..8 inputs supported"); }
    b200::ReduceArgs a;
    a.ins_num = ins_num;
    for (int i = 0; i < ins_num; ++i) {
      var_info_t &vi = var("ins_" + str(i));
      if (!(vi.dims == vout.dims)) { rt_err("reduce: input dims differ from out"); }
      a.ins[i] = fptr(vi);
    }
    long long const n = vout.dims.dims_prod();
    int vec4 = (reinterpret_cast<uintptr_t>(fptr(vout)) & 15) == 0 ? 1 : 0;
    for (int i = 0; i < ins_num; ++i) { if (reinterpret_cast<uintptr_t>(a.ins[i]) & 15) { vec4 = 0; } }
    int const blocks = (int)std::min<long long>(ceil_div(ceil_div(n, vec4 ? 4 : 1), 256), (long long)im.num_sms * 16);
    B200_CARVEOUT_ONCE(b200::reduce_sum_kernel); launch_k(b200::reduce_sum_kernel, dim3(std::max(blocks, 1)), dim3(256), 0, a, fptr(vout), n, (int)scalar("relu", true, 0), absmax_cell("out"), vec4);
    launched();
    im.bump(vout);
  }

  // parameter-only: fold BatchNorm (use_global_stats) / Scale into a convolution's filters and biases (pointwise.cuh: bn_fold_kernel)
  void run_bn_fold() {
    var_info_t &vf = var("filts"), &vof = var("out_filts"), &vob = var("out_biases");
    if (!(vf.dims == vof.dims)) { rt_err("bn_fold: out_filts dims differ from filts"); }
    int const OC = vf.dims.dsz("out_chan");
    if ((int)vob.dims.dims_prod() != OC) { rt_err("bn_fold: out_biases size mismatch"); }
    auto opt = [&](char const *an, uint64_t want) -> float const * {
      if (!has_arg(an)) { return nullptr; }
      var_info_t &v = var(an);
      if (v.dims.dims_prod() != want) { rt_err(string("bn_fold: '") + an + "' has " + str(v.dims.dims_prod()) + " elements, expected " + str(want)); }
      return fptr(v);
    };
    float const *biases = opt("biases", OC), *mean = opt("mean", OC), *varp = opt("var", OC), *sf = opt("sf", 1), *gamma = opt("gamma", OC), *beta = opt("beta", OC);
    if ((mean != nullptr) != (varp != nullptr) || (mean != nullptr) != (sf != nullptr)) { rt_err("bn_fold: mean, var and sf come together"); }
    if ((gamma != nullptr) != (beta != nullptr)) { rt_err("bn_fold: gamma and beta come together"); }
    long long const per_oc = (long long)(vf.dims.dims_prod() / OC);
    B200_CARVEOUT_ONCE(b200::bn_fold_kernel);
    launch_k(b200::bn_fold_kernel, dim3(ceil_div(per_oc, 256), OC), dim3(256), 0, fptr(vf), biases, mean, varp, sf, gamma, beta, (float)scalar("eps", true, 1e-5), fptr(vof), fptr(vob), per_oc);
    launched();
    im.bump(vof);
    im.bump(vob);
  }

  void run_gen_data() {  // test/rtc/gen_data_*.cucl; the generator's name carries the op type and the argument
    var_info_t &v = var(f.gen_arg);
    string const &fn = f.op.get_func_name();
    uint32_t const mode = (uint32_t)scalar("mode");
    float const vi = (float)scalar("vi", true, 0.0);
    long long const n = v.dims.dims_prod();
    if (n >= (1ll << 32)) { unsup_err("gen_data: tensors >= 2^32 elements (the reference's flat index is uint32_t, src/boda_base.H:608)"); }
    int kind; uint32_t inner = 1, inner2 = 1, salt;
    if (fn == "gen_data_Convolution_in" || fn == "gen_data_Convolution_filts") {
      kind = 0; inner = v.dims.dsz("x"); inner2 = v.dims.dsz("y"); salt = (f.gen_arg == "in") ? 234234567u : 8753985u;
    } else if (fn == "gen_data_Convolution_biases") { kind = 1; salt = 39475612u; }
    else if (fn == "gen_data_sgemm_a") { kind = 2; inner = v.dims.dsz("M"); inner2 = v.dims.dsz("K"); salt = 12738732u; }
    else if (fn == "gen_data_sgemm_b") { kind = 3; inner = v.dims.dsz("N"); inner2 = v.dims.dsz("K"); salt = 12738732u; }
    else { unsup_err("be=b200: unknown generator '" + fn + "'"); }
    B200_CARVEOUT_ONCE(b200::gen_data_kernel); launch_k(b200::gen_data_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, fptr(v), (uint32_t)n, kind, inner, inner2, mode, vi, salt);
    launched();
    im.bump(v);
  }
};

}  // namespace

uint32_t b200_compute_t::run(rtc_func_call_t const &rfc) {
  assert_st(impl->inited);
  if (plan_only) { rt_err("run: this is a plan-only instance (no device); there is no CPU fallback"); }
  auto fi = impl->funcs.find(rfc.rtc_func_name);
  if (fi == impl->funcs.end()) { rt_err("run: function '" + rfc.rtc_func_name + "' was not compiled"); }
  for (auto const &kv : rfc.arg_map) { if (!kv.second.is_valid()) { rt_err("run: invalid argument '" + kv.first + "'"); } }
  call_ev_t ev;
  if (impl->timing) {
    CU_CHK(cudaEventCreate(&ev.b));
    CU_CHK(cudaEventCreate(&ev.e));
    CU_CHK(cudaEventRecord(ev.b, impl->stream));
  }
  run_ctx_t ctx{*this, *impl, fi->second, rfc, impl->stream, &ev};
  switch (fi->second.kind) {
    case FK_CONV: ctx.run_conv(); break;
    case FK_SGEMM: ctx.run_sgemm(); break;
    case FK_POOL: ctx.run_pool(); break;
    case FK_LRN: ctx.run_lrn(); break;
    case FK_RELU: ctx.run_relu(); break;
    case FK_SOFTMAX: ctx.run_softmax(); break;
    case FK_COPY: ctx.run_copy(); break;
    case FK_REDUCE: ctx.run_reduce(); break;
    case FK_GEN_DATA: ctx.run_gen_data(); break;
    case FK_BN_FOLD: ctx.run_bn_fold(); break;
    case FK_FC_CHAIN: ctx.run_fc_chain(); break;
  }
  if (impl->timing) { CU_CHK(cudaEventRecord(ev.e, impl->stream)); }
  impl->calls.push_back(ev);
  return static_cast<uint32_t>(impl->calls.size() - 1);
}

}  // namespace boda
