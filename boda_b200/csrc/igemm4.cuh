// igemm4.cuh -- the round-2 contraction kernel: persistent CTA pairs (tcgen05 cta_group::2) with
//   * a HALO operand mode for stride-1 KH x KW convolutions (the 3x3 / 5x5 layers that dominate AlexNet-ng / GoogLeNet / NiN / ResNet):
//     one shared-memory tile of 128 + (KH-1)*Wp + (KW-1) activation rows per 64-channel block feeds EVERY filter tap through UMMA descriptors
//     that start (ky*Wp + kx) rows into it (the layout and the legality argument are in igemm3.cuh), so the activation bytes cross the
//     L2 -> SM path once per channel block instead of once per tap. Measured on the im2col pair kernel (profiles/diag_r02a_conv_roles.txt):
//     the MMA warp spends 22 % (bf16) of its loop waiting for `full` barriers and both precisions move the same ~35 B/cycle/SM -- the kernel is
//     bound by the operand feed, not by the tensor pipe; the halo mode cuts the bytes per MMA cycle from 96 to ~40 (bf16, BN = 128);
//   * a deep ring of SMALL filter stages (8 KB per k-block and CTA): 16+ k-blocks in flight cover the ~2.5k-cycle TMA round trip that 4 fat
//     stages (8 k-blocks = 2048 MMA cycles) of the im2col kernel could not;
//   * STREAM-K work distribution: the (tile, k-block) space is cut into one contiguous, equally long range per CTA pair, so 75 tiles on 74 pairs
//     (AlexNet conv3-5 in the halo layout) or 49 tiles on 74 pairs (14x14 layers) cost 1.0 rounds instead of 2 or 0.66 of a round. A pair that
//     starts inside a tile writes its partial accumulators to a workspace slot and raises a flag; the pair that owns the tile's FIRST k-block
//     (it reaches that tile last in its own range) adds the partials in pair order and runs the epilogue -- deterministic, no atomics on data,
//     no second kernel (this also replaces split-K + splitk_reduce_kernel for the inner-product layers);
//   * the per-k-block operand modes of igemm2.cuh (2-d tiled, im2col TMA) behind the same loop, so strided / 1x1 / inner-product-shaped
//     layers and sgemm get the stream-K scheduler too.
// Replaces the same CUCL functions as igemm.cuh (conv / tconv / k1conv / ipconv / sgemm*, test/rtc/*.cucl + src/cnn_codegen.cc).
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + (leader CTA) MMA issuer, warps 2..9 = epilogue (two warpgroups,
// each thread one row x half the tile's columns).
#pragma once
#include "igemm.cuh"

namespace b200 {

constexpr int SK4_MAX_A_STAGES = 4, SK4_MAX_B_STAGES = 12;
constexpr int SK4_BAR_BYTES = 3072;  // barriers in the first 512 B, reduction scratch at +512, stage records at +1024, two staged bias vectors (segment parity) at +2048
constexpr int SK4_KU_MAX = 4;        // at most 4 k-blocks (single-plane) / 2 k-blocks x 2 planes per stage: 4 operand slots

// What the MMA warp needs to know about one k-block of a stage, written by the producer into shared memory next to the stage's barrier.
// The tensor pipe queues only ~2 MMAs: every instruction the issuing thread executes between two tcgen05.mma is tensor-pipe idle time, and with
// all operand bookkeeping (tap -> descriptor offset, ring slots, K tails) on that thread the stage loop measured 500-700 cycles WITHOUT any MMA
// (profiles/diag_r02*: "noMMA" runs). So the producer, which walks the same (channel block, tap) sequence anyway and has slack, precomputes it.
struct Sk4Rec { uint32_t p_lo; uint32_t flags; };  // P-operand descriptor low word (hi plane); flags below
constexpr uint32_t SK4_F_NK = 7u, SK4_F_VALID = 8u, SK4_F_ADONE = 16u, SK4_F_AWAIT = 32u, SK4_F_APAR = 64u, SK4_F_ASLOT_SHIFT = 8u;

struct Sk4Params {
  IgemmParams g;     // extents, im2col addressing, epilogue, scales, layout-transform-elimination outputs: exactly as in igemm.cuh
  int p_mode;        // P operand: 0 = 2-d tiled {k, row}, 1 = im2col TMA (one activation tile per k-block), 2 = halo (one tile per channel block)
  int taps;          // halo: KH*KW (k-blocks are ordered channel block major, tap minor); 1 otherwise
  int Wp, HpWp, OH, OW;                 // halo: padded row pitch, virtual pixels per image, valid output rows / columns
  int halo_rows, a_loads, a_box_rows;   // halo: activation rows per stage = a_loads boxes of a_box_rows
  int a_stages, b_stages;               // ring depths (modes 0/1: one ring, the P tiles ride on the Q stages' barriers)
  int ku;            // k-blocks per stage-unit: ku * planes <= 4 operand slots per stage (halo: 4 / planes, per-k-block operand modes: 2 / planes)
  int sk;            // 1 = stream-K ranges, 0 = whole tiles, pair p takes tiles p, p + #pairs, ...
  int n_tiles, ukb;  // tiles; stage-units (kKb k-blocks) per tile
  float *sk_ws;      // stream-K partials: [CTA slot = pair*2 + rank][BN][128] fp32, raw accumulator units
  unsigned int *sk_flags;  // one per CTA slot: raised by the contributor, reset by the finisher
  int out_w;         // output width, for the padded destination planes
  int o16_Hp, o16_Wp, o16_py, o16_px;  // destination planes in the shared-padding layout of a halo-mode consumer (o16_Wp = 0: plain pixel-major)
};

template <int BN, int kPlanes>
struct Sk4Cfg {
  static constexpr int kBQ = BN / 2;                  // filter rows per CTA
  static constexpr uint32_t kBSlot = kBQ * 128;       // one k-block of one plane of the Q operand, per CTA
  static constexpr uint32_t kPSlot = IGEMM_BM * 128;  // modes 0/1: one k-block of one plane of the P operand
  static constexpr uint32_t kBufCols = tmem_buf_cols(BN);
  static constexpr uint32_t kColsNeeded = (kPlanes == 2 ? 4 : 2) * kBufCols;
  static constexpr uint32_t kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128 : kColsNeeded <= 256 ? 256 : 512;
  static_assert(kColsNeeded <= 512, "TMEM has 512 columns");
};

// contributor -> finisher hand-off of stream-K partials (gpu scope)
__device__ __forceinline__ void sk4_flag_raise(unsigned int *f) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(1u) : "memory"); }
__device__ __forceinline__ unsigned int sk4_flag_peek(unsigned int const *f) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  return v;
}
__device__ __forceinline__ void sk4_flag_wait_and_reset(unsigned int *f) {
  uint32_t spins = 0;
  while (sk4_flag_peek(f) == 0u) {
    __nanosleep(40);
    if (++spins > (1u << 23)) { printf("b200: stream-K partial never arrived (block %d flag %p)\n", blockIdx.x, (void *)f); __trap(); }
  }
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(f), "r"(0u) : "memory");  // ready for the next launch (one consumer per flag)
}

// where GEMM row `prow` of a pixel-major launch lives in the output: image, pixel, (y, x)
struct Sk4RowGeom { bool valid; int img, pix, oy, ox; };
__device__ __forceinline__ Sk4RowGeom sk4_row_geom(Sk4Params const &prm, int prow, bool halo) {
  Sk4RowGeom r;
  if (halo) {  // rows are virtual pixels of the shared-padding layout: the ones in the padding columns / rows are computed and dropped
    r.img = prow / prm.HpWp;
    int const rem = prow - r.img * prm.HpWp;
    r.oy = rem / prm.Wp; r.ox = rem - r.oy * prm.Wp;
    r.valid = prow < prm.g.p_rows && r.oy < prm.OH && r.ox < prm.OW;
    r.pix = r.oy * prm.OW + r.ox;
  } else {
    r.img = prow / prm.g.out_hw; r.pix = prow - r.img * prm.g.out_hw;
    r.valid = prow < prm.g.p_rows;
    r.oy = r.pix / prm.out_w; r.ox = r.pix - r.oy * prm.out_w;
  }
  return r;
}

// the work of one pair: stream-K = a contiguous range of stage-units, else whole tiles with stride #pairs. Warp-uniform. 32-bit arithmetic
// on purpose (the host guarantees units * pairs < 2^31): a 64-bit integer division is ~100 instructions on the single thread that walks this.
struct Sk4Work {
  uint32_t u, u_end;  // stream-K
  int tile_next, n_tiles, stride, ukb;
  bool sk;
  static __device__ __forceinline__ uint32_t range_begin(uint32_t U, int pair, int n_pairs) { return (U * static_cast<uint32_t>(pair)) / static_cast<uint32_t>(n_pairs); }
  __device__ __forceinline__ Sk4Work(Sk4Params const &prm, int pair, int n_pairs) {
    sk = prm.sk != 0; n_tiles = prm.n_tiles; stride = n_pairs; ukb = prm.ukb; tile_next = pair;
    uint32_t const U = static_cast<uint32_t>(prm.n_tiles) * static_cast<uint32_t>(prm.ukb);
    u = range_begin(U, pair, n_pairs); u_end = range_begin(U, pair + 1, n_pairs);
  }
  __device__ __forceinline__ bool next(int &tile, int &u0, int &u1) {
    if (sk) {
      if (u >= u_end) { return false; }
      tile = static_cast<int>(u / static_cast<uint32_t>(ukb));
      u0 = static_cast<int>(u) - tile * ukb;
      int const left = static_cast<int>(u_end - u);
      u1 = (left < ukb - u0) ? u0 + left : ukb;
      u += static_cast<uint32_t>(u1 - u0);
      return true;
    }
    if (tile_next >= n_tiles) { return false; }
    tile = tile_next; tile_next += stride; u0 = 0; u1 = ukb;
    return true;
  }
};

// kEpi selects which epilogue code is COMPILED IN (the instruction cache is the reason: the straight-line epilogues of all variants together are
// ~190 KB of SASS; ncu attributed 35 % of the warp stalls of the all-in-one kernel to `no_inst`, 4.5x the round-1 kernel -- profiles/):
//   SK4_EPI_LEAN = fp32 NCHW node only; SK4_EPI_FULL = + residual input + the consumers' 16-bit planes; SK4_EPI_SWAPPED = weights as the 128-row
//   operand (inner-product-shaped layers: row = out chan, columns = pixels).
constexpr int SK4_EPI_LEAN = 0, SK4_EPI_FULL = 1, SK4_EPI_SWAPPED = 2;

constexpr int SK4_THREADS = 320;  // warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue (two warpgroups)

// acc[0..N) += this warp's 32 lanes x N fp32 columns of TMEM starting at taddr (N a multiple of 16)
template <int N>
__device__ __forceinline__ void sk4_drain(float (&acc)[N], uint32_t taddr) {
#pragma unroll
  for (int j0 = 0; j0 + 32 <= N; j0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(taddr + j0, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
  }
  if (N % 32) {
    uint32_t r[16];
    tmem_ld_32x16(taddr + (N / 32) * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) { acc[(N / 32) * 32 + j] += __uint_as_float(r[j]); }
  }
}

template <int BN, int kPlanes, int kEpi>
__global__ void __launch_bounds__(SK4_THREADS, 1)
igemm_sk4_kernel(const __grid_constant__ CUtensorMap p_hi_map, const __grid_constant__ CUtensorMap p_lo_map,
                 const __grid_constant__ CUtensorMap q_hi_map, const __grid_constant__ CUtensorMap q_lo_map, const Sk4Params prm) {
  using Cfg = Sk4Cfg<BN, kPlanes>;
  constexpr uint32_t kBSlot = Cfg::kBSlot, kPSlot = Cfg::kPSlot, kBufCols = Cfg::kBufCols;
  IgemmParams const &g = prm.g;
  bool const halo = (prm.p_mode == 2);
  int const a_stages = prm.a_stages, b_stages = prm.b_stages;
  int const ku = prm.ku;                                         // k-blocks per stage
  uint32_t const slots = static_cast<uint32_t>(ku) * kPlanes;    // operand slots per stage (<= 4)
  uint32_t const b_stage = slots * kBSlot;
  // operand slots of a stage: [hi plane of k-block 0 .. ku-1][lo plane of k-block 0 .. ku-1]
  uint32_t const a_plane = halo ? static_cast<uint32_t>(prm.halo_rows) * 128u : static_cast<uint32_t>(ku) * kPSlot;  // distance hi plane -> lo plane of a P tile
  uint32_t const b_plane = static_cast<uint32_t>(ku) * kBSlot;                                                     // same for a Q tile
  uint32_t const a_stage = halo ? kPlanes * a_plane : slots * kPSlot;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_ring = smem;
  uint8_t *b_ring = a_ring + a_stages * a_stage;
  uint8_t *bar_mem = b_ring + b_stages * b_stage;
  uint64_t *a_full = reinterpret_cast<uint64_t *>(bar_mem);  // leader's copies collect both CTAs' bytes
  uint64_t *a_empty = a_full + SK4_MAX_A_STAGES;
  uint64_t *b_full = a_empty + SK4_MAX_A_STAGES;
  uint64_t *b_empty = b_full + SK4_MAX_B_STAGES;
  uint64_t *tmem_full_bar = b_empty + SK4_MAX_B_STAGES;  // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;          // [2], leader's copies collect both CTAs' epilogue warps
  uint64_t *x_empty_bar = tmem_empty_bar + 2;            // [2], cross-term accumulators (fp32-parity mode)
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(x_empty_bar + 2);
  Sk4Rec *recs = reinterpret_cast<Sk4Rec *>(bar_mem + 1024);  // [b_stages][SK4_KU_MAX]
  float *bias_s = reinterpret_cast<float *>(bar_mem + 2048);  // [2][BN]

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  uint32_t const cta_rank = blockIdx.x & 1u;  // = %cluster_ctarank for (2,1,1) clusters
  bool const leader = (cta_rank == 0);
  int const n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  int const q_tiles = g.q_tiles;
  int const nkb = g.kblks_total;
  int const chunk_u = max(1, g.chunk_kblks / ku);  // stage-units per accumulation chunk
  int const taps = prm.taps, cblks = g.cblks;

  if (warp_id == 0) {  // barrier setup spread over the lanes (one thread doing ~40 mbarrier.init in a row is a microsecond of prologue)
    if (lane == 0) {
      tma_prefetch_desc(&p_hi_map);
      tma_prefetch_desc(&q_hi_map);
      if (kPlanes == 2) { tma_prefetch_desc(&p_lo_map); tma_prefetch_desc(&q_lo_map); }
    }
    if (lane < a_stages) { mbar_init(&a_full[lane], 2); mbar_init(&a_empty[lane], 1); }
    if (lane < b_stages) { mbar_init(&b_full[lane], 2); mbar_init(&b_empty[lane], 1); }
    if (lane < 2) { mbar_init(&tmem_full_bar[lane], 1); mbar_init(&tmem_empty_bar[lane], 16); mbar_init(&x_empty_bar[lane], 16); }  // 8 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp_id == 1) { tmem_alloc_2sm<Cfg::kTmemCols>(tmem_ptr_smem); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  uint32_t const tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer (both CTAs; whole warp walks the loop with warp-uniform state, one elected lane issues) ==========
    long long const t_begin = g.ts ? clock64() : 0;
    long long w_empty = 0, t_issue = 0;
    uint32_t const a_base = smem_u32(a_ring);
    int const ksteps_full = IGEMM_BK / IGEMM_UMMA_K;
    Sk4Work work(prm, pair, n_pairs);
    int tile, u0, u1;
    int sb = 0; uint32_t b_par = 1;  // next B stage, parity to wait for on its empty barrier (fresh barriers pass parity 1)
    // halo: activation tiles are issued and consumed in one global order. Issue side: next slot / parity of its empty barrier. Consume side (what
    // the records tell the MMA warp): slot / parity of the full barrier of the channel block being multiplied.
    int ia = 0; uint32_t ia_par = 1;
    int ca = 0; uint32_t ca_par = 0;
    while (!(g.debug & 1) && work.next(tile, u0, u1)) {
      int const mt = tile / q_tiles, nt = tile - mt * q_tiles;
      int const m0 = (mt * 2 + static_cast<int>(cta_rank)) * IGEMM_BM;
      int const q_row0 = nt * BN + static_cast<int>(cta_rank) * Cfg::kBQ;
      int const kb0 = u0 * ku, kb1 = min(u1 * ku, nkb);
      // ---- per-segment operand addressing state (integer divisions here only: once per segment, never per k-block) ----
      int img = 0, h_base = 0, w_base = 0, cb = 0, kx = 0, ky = 0, kb_in_grp = 0;  // per-k-block operand modes
      int c = 0, t = 0, kxx = 0, tap_row = 0, c_next = 0, c_last = 0;             // halo: channel block / tap of the next k-block; next A tile to request; last one of the segment
      bool a_first = true;                                                        // halo: the next k-block is the first one of its channel block in this segment
      if (prm.p_mode == 1) {
        img = m0 / g.ohw;
        int const rem = m0 - img * g.ohw;
        int const oy = rem / g.ow, ox = rem - oy * g.ow;
        h_base = oy * g.sy - g.py;
        w_base = ox * g.sx - g.px;
        int const tap0 = kb0 / cblks;
        cb = kb0 - tap0 * cblks; ky = tap0 / g.kw; kx = tap0 - ky * g.kw;
      } else if (halo) {
        c = kb0 / taps; t = kb0 - c * taps;
        int const kyy = t / g.kw; kxx = t - kyy * g.kw; tap_row = kyy * prm.Wp + kxx;
        c_next = c; c_last = (kb1 - 1) / taps;
      }
      if (!halo) { kb_in_grp = kb0 % g.kb_mod; }
      auto issue_a = [&](int cc) {  // halo tile of channel block cc: rows [m0, m0 + halo_rows) of the padded activation matrix
        if (elect_one_sync()) {
          if (leader) { mbar_expect_tx(&a_full[ia], 2u * a_stage); } else { mbar_arrive_remote(&a_full[ia], 0); }
          uint8_t *dst = a_ring + ia * a_stage;
          for (int l = 0; l < prm.a_loads; ++l) {
            uint8_t *d = dst + l * prm.a_box_rows * 128;
            int const row = m0 + l * prm.a_box_rows;
            tma_load_2d_2sm(d, &p_hi_map, &a_full[ia], cc * IGEMM_BK, row);
            if (kPlanes == 2) { tma_load_2d_2sm(d + a_plane, &p_lo_map, &a_full[ia], cc * IGEMM_BK, row); }
          }
        }
        __syncwarp();
        if (++ia == a_stages) { ia = 0; ia_par ^= 1; }
      };
      if (halo) {
        // ---- halo mode: per stage ONE filter load per plane (a 3-d box of ku consecutive k-blocks: the filters are packed channel block
        // major, tap minor for this mode) + the records; activation tiles as their ring frees up. This thread's instruction count per stage IS
        // the producer's speed (one warp, dependent chains: ~5 cycles an instruction), so everything is kept incremental.
        uint32_t p_cur = sw128_desc_lo(a_base + static_cast<uint32_t>(ca) * a_stage + static_cast<uint32_t>(tap_row) * 128u);  // descriptor low word of (block c, tap t)
        uint32_t const row_wrap = static_cast<uint32_t>(prm.Wp - (g.kw - 1)) * 8u;  // last tap of a filter row -> first tap of the next: (Wp - (kw - 1)) rows of 128 B, in 16-byte units
        int kbi = c * taps + t;  // index of the next k-block in the packed filters
        for (int kb = kb0; kb < kb1; kb += ku) {
          int const nh = min(ku, kb1 - kb);
          {
            int const c_need = min(c + (t + nh - 1) / taps, c_last);
            while (c_next <= c_last) {
              if (c_next <= c_need) { mbar_wait(&a_empty[ia], ia_par); }
              else if (!mbar_test_wait(&a_empty[ia], ia_par)) { break; }  // (test_wait: try_wait would suspend the producer for its hardware time limit)
              issue_a(c_next);
              ++c_next;
            }
          }
          if (g.ts) { long long const t0 = clock64(); mbar_wait(&b_empty[sb], b_par); w_empty += clock64() - t0; } else { mbar_wait(&b_empty[sb], b_par); }
          long long const t_i0 = g.ts ? clock64() : 0;
          bool const issue = elect_one_sync();
          if (issue) {
            uint8_t *bst = b_ring + sb * b_stage;
            tma_load_3d_2sm(bst, &q_hi_map, &b_full[sb], 0, q_row0, kbi);  // slots 0..ku-1 (single plane) / hi planes
            if (kPlanes == 2) { tma_load_3d_2sm(bst + b_plane, &q_lo_map, &b_full[sb], 0, q_row0, kbi); }
          }
          Sk4Rec *rec = recs + sb * SK4_KU_MAX;
#pragma unroll
          for (int j = 0; j < SK4_KU_MAX; ++j) {
            if (j < nh) {
              bool const done = (t == taps - 1) || (kb + j == kb1 - 1);  // last tap of the block, or the segment ends inside it: the tile is released
              uint32_t flags = SK4_F_VALID | static_cast<uint32_t>((c == cblks - 1) ? g.ksteps_last : ksteps_full) | (static_cast<uint32_t>(ca) << SK4_F_ASLOT_SHIFT);
              if (a_first) { flags |= SK4_F_AWAIT | (ca_par ? SK4_F_APAR : 0u); }
              if (done) { flags |= SK4_F_ADONE; }
              if (issue) { rec[j].p_lo = p_cur; rec[j].flags = flags; }
              a_first = done;
              if (done) { if (++ca == a_stages) { ca = 0; ca_par ^= 1; } }
              if (++t == taps) { t = 0; kxx = 0; ++c; p_cur = sw128_desc_lo(a_base + static_cast<uint32_t>(ca) * a_stage); }
              else if (done) { p_cur = sw128_desc_lo(a_base + static_cast<uint32_t>(ca) * a_stage); }  // (segment end inside a block: the next segment re-initialises anyway)
              else if (++kxx == g.kw) { kxx = 0; p_cur += row_wrap; } else { p_cur += 8u; }
            } else if (issue) { rec[j].flags = 0u; }
          }
          kbi += ku;
          // the records are ordinary shared-memory stores of the issuing lane: its arrive on the full barrier (release) publishes them to the
          // MMA warp's wait (acquire). The box always carries ku k-blocks per plane (a short last stage just multiplies fewer of them).
          if (issue) {
            if (leader) { mbar_expect_tx(&b_full[sb], 2u * slots * kBSlot); } else { mbar_arrive_remote(&b_full[sb], 0); }
          }
          __syncwarp();
          if (g.ts) { t_issue += clock64() - t_i0; }
          if (++sb == b_stages) { sb = 0; b_par ^= 1; }
        }
      } else {
        // ---- per-k-block operand modes (2-d tiled / im2col TMA): P and Q tile of every k-block of the stage ----
        for (int kb = kb0; kb < kb1; kb += ku) {
          int const nh = min(ku, kb1 - kb);
          if (g.ts) { long long const t0 = clock64(); mbar_wait(&b_empty[sb], b_par); w_empty += clock64() - t0; } else { mbar_wait(&b_empty[sb], b_par); }
          long long const t_i0 = g.ts ? clock64() : 0;
          bool const issue = elect_one_sync();
          uint8_t *bst = b_ring + sb * b_stage;
          uint8_t *pst = a_ring + sb * a_stage;
          Sk4Rec *rec = recs + sb * SK4_KU_MAX;
#pragma unroll
          for (int j = 0; j < SK4_KU_MAX; ++j) {  // the records first: this lane's arrive below (release) publishes them
            if (j < nh) {
              int nk = ksteps_full;
              if (++kb_in_grp == g.kb_mod) { kb_in_grp = 0; nk = g.ksteps_last; }
              if (issue) {
                rec[j].p_lo = sw128_desc_lo(a_base + static_cast<uint32_t>(sb) * a_stage + static_cast<uint32_t>(j) * kPSlot);
                rec[j].flags = SK4_F_VALID | static_cast<uint32_t>(nk);
              }
            } else if (issue) { rec[j].flags = 0u; }
          }
          if (issue) {  // expect_tx, then the loads, as in igemm2.cuh
            if (leader) { mbar_expect_tx(&b_full[sb], 2u * static_cast<uint32_t>(nh) * kPlanes * (kBSlot + kPSlot)); } else { mbar_arrive_remote(&b_full[sb], 0); }
          }
#pragma unroll
          for (int j = 0; j < SK4_KU_MAX; ++j) {
            if (j < nh) {
              int const kbj = kb + j;
              if (issue) {
                uint8_t *p_hi = pst + static_cast<uint32_t>(j) * kPSlot, *p_lw = p_hi + a_plane;
                if (prm.p_mode == 1) {  // the P tile first: it is the slower request (128 gathered pixel rows)
                  tma_load_im2col_4d_2sm(p_hi, &p_hi_map, &b_full[sb], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky);
                  if (kPlanes == 2) { tma_load_im2col_4d_2sm(p_lw, &p_lo_map, &b_full[sb], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky); }
                } else {
                  int const pc0 = g.p_kb_rows ? 0 : kbj * IGEMM_BK, pc1 = m0 + kbj * g.p_kb_rows;
                  tma_load_2d_2sm(p_hi, &p_hi_map, &b_full[sb], pc0, pc1);
                  if (kPlanes == 2) { tma_load_2d_2sm(p_lw, &p_lo_map, &b_full[sb], pc0, pc1); }
                }
                uint8_t *q_hi = bst + static_cast<uint32_t>(j) * kBSlot, *q_lw = q_hi + b_plane;
                int const qc0 = g.q_kb_rows ? 0 : kbj * IGEMM_BK, qc1 = q_row0 + kbj * g.q_kb_rows;
                tma_load_2d_2sm(q_hi, &q_hi_map, &b_full[sb], qc0, qc1);
                if (kPlanes == 2) { tma_load_2d_2sm(q_lw, &q_lo_map, &b_full[sb], qc0, qc1); }
              }
              if (prm.p_mode == 1) { if (++cb == cblks) { cb = 0; if (++kx == g.kw) { kx = 0; ++ky; } } }
            }
          }
          __syncwarp();
          if (g.ts) { t_issue += clock64() - t_i0; }
          if (++sb == b_stages) { sb = 0; b_par ^= 1; }
        }
      }
    }
    if (g.ts && leader && lane == 0) { long long *ts = g.ts + pair * 16; ts[0] = clock64() - t_begin; ts[1] = w_empty; ts[13] = t_issue; }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (leader CTA only, one elected thread for the pair) =====================
    // Per stage: wait for the stage, read its records, issue. No operand arithmetic on this thread (see Sk4Rec).
    if (leader) {
      long long const t_begin = g.ts ? clock64() : 0;
      long long w_full = 0, w_tmem = 0, w_afull = 0, t_first_full = 0;
      uint32_t const idesc = g.idesc;  // M = 256 (the pair), N = BN
      uint32_t const b_base = smem_u32(b_ring), recs_base = smem_u32(recs);
      uint32_t const p_lo_delta = a_plane >> 4;  // descriptor low word: start address in 16-byte units
      bool const no_tma = (g.debug & 1) != 0, no_mma = (g.debug & 2) != 0;
      Sk4Work work(prm, pair, n_pairs);
      int tile, u0, u1;
      int sb = 0; uint32_t b_par = 0;
      int gc = 0, si = 0, n_stage = 0;  // accumulation chunks, segments and stages done so far
      while (work.next(tile, u0, u1)) {
        uint32_t const tmem_x = tmem_base + (2 + (si & 1)) * kBufCols;
        if (kPlanes == 2) { mbar_wait(&x_empty_bar[si & 1], ((si >> 1) & 1) ^ 1); }  // the epilogue has read the cross terms of segment si - 2
        uint32_t x_acc = 0u;  // the segment's first MMA overwrites the cross-term accumulator
        for (int u = u0; u < u1;) {
          int const buf = gc & 1;
          if (g.ts) { long long const t0 = clock64(); mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); w_tmem += clock64() - t0; } else { mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); }
          uint32_t const tmem_d = tmem_base + buf * kBufCols;
          int const u_end = min(u + chunk_u, u1);
          uint32_t main_acc = 0u;  // the chunk's first MMA overwrites the main accumulator
          for (; u < u_end; ++u, ++n_stage) {
            if (no_tma) {}  // experiments: MMA on whatever the shared memory holds
            else if (g.ts) { long long const t0 = clock64(); mbar_wait(&b_full[sb], b_par); long long const t1 = clock64(); if (n_stage == 0) { t_first_full = t1 - t_begin; } else { w_full += t1 - t0; } }
            else { mbar_wait(&b_full[sb], b_par); }
            uint32_t rp[SK4_KU_MAX], rf[SK4_KU_MAX];
            {
              uint32_t const ra = recs_base + static_cast<uint32_t>(sb) * (SK4_KU_MAX * 8u);
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rp[0]), "=r"(rf[0]), "=r"(rp[1]), "=r"(rf[1]) : "r"(ra));
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(rp[2]), "=r"(rf[2]), "=r"(rp[3]), "=r"(rf[3]) : "r"(ra + 16u));
            }
            if (no_tma) {  // (no producer, no records: a full-K stage on stage-aligned addresses)
#pragma unroll
              for (int j = 0; j < SK4_KU_MAX; ++j) { rp[j] = sw128_desc_lo(smem_u32(a_ring)); rf[j] = (j < ku) ? (SK4_F_VALID | 4u) : 0u; }
            }
#pragma unroll
            for (int j = 0; j < SK4_KU_MAX; ++j) {  // halo: the first k-block of a channel block waits for its activation tile
              if (rf[j] & SK4_F_AWAIT) {
                uint64_t *bar = &a_full[(rf[j] >> SK4_F_ASLOT_SHIFT) & 3u];
                uint32_t const par = (rf[j] & SK4_F_APAR) ? 1u : 0u;
                if (g.ts) { long long const t0 = clock64(); mbar_wait(bar, par); w_afull += clock64() - t0; } else { mbar_wait(bar, par); }
              }
            }
            tc_fence_after();
            uint32_t const bst = b_base + static_cast<uint32_t>(sb) * b_stage;
            bool const chunk_end = (u + 1 == u_end);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < SK4_KU_MAX; ++j) {
                if (rf[j] & SK4_F_VALID) {
                  uint32_t const q_addr = bst + static_cast<uint32_t>(j) * kBSlot;
                  issue_kblock<kPlanes, true>(tmem_d, tmem_x, rp[j], rp[j] + p_lo_delta, sw128_desc_lo(q_addr), sw128_desc_lo(q_addr + b_plane), idesc, main_acc, x_acc,
                                              no_mma ? 0 : static_cast<int>(rf[j] & SK4_F_NK));
                  main_acc = 1u; x_acc = 1u;
                  if (rf[j] & SK4_F_ADONE) { umma_commit_2sm(&a_empty[(rf[j] >> SK4_F_ASLOT_SHIFT) & 3u], 0x3); }
                }
              }
              umma_commit_2sm(&b_empty[sb], 0x3);  // release the stage in both CTAs
              if (chunk_end) { umma_commit_2sm(&tmem_full_bar[buf], 0x3); }  // wake both CTAs' epilogues
            }
            __syncwarp();
            main_acc = 1u; x_acc = 1u;
            if (++sb == b_stages) { sb = 0; b_par ^= 1; }
          }
          ++gc;
        }
        ++si;
      }
      if (g.ts && lane == 0) { long long *ts = g.ts + pair * 16; ts[2] = clock64() - t_begin; ts[3] = w_full; ts[4] = w_tmem; ts[5] = t_first_full; ts[6] = n_stage; ts[7] = w_afull; }
    }
  } else {
    // ===================== epilogue: 8 warps = 2 warpgroups (each CTA: its own 128 rows; each thread one row x HALF the tile's columns) ==========
    // Two warpgroups because (1) the straight-line epilogue code is per column: half the columns per thread halves the instructions every tile
    // streams through the 32 KB instruction cache it shares with the producer / MMA loops (ncu on the first version of this kernel: 35 % of
    // all warp stalls were `no_instruction`, 4.5x the round-1 kernel, profiles/) and (2) twice the warps keep twice the stores in flight.
    constexpr int HN = BN / 2;                 // columns per thread
    int const ewarp = warp_id - 2;
    int const q = warp_id & 3;                 // TMEM lane quarter this warp may read
    int const half = ewarp >> 2;               // which half of the tile's columns
    int const row = q * 32 + lane;
    int const etid = half * 128 + row;         // 0..255 among the epilogue threads
    int const c0 = half * HN;                  // first column of this thread inside the tile
    float const inv = g.p_scale[1] * g.q_scale[1], inv_recip = g.p_scale[0] * g.q_scale[0];
    float const floor_v = g.relu ? 0.0f : -INFINITY;
    float amax = 0.0f;
    float s_out = 1.0f;
    if (kEpi == SK4_EPI_FULL && g.out16 && g.w_l1max) {  // scale of the consumer's fp16 planes from the output bound (identical in every CTA); CTA 0 publishes it
      float *red = reinterpret_cast<float *>(bar_mem + 512);
      float bm = 0.0f;
      if (g.has_bias) { for (int j = etid; j < g.n_bias; j += 256) { bm = fmaxf(bm, fabsf(__ldg(g.bias + j))); } }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o)); }
      if (lane == 0) { red[ewarp] = bm; }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      bm = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
      float in_bound = 16384.0f * g.p_scale[1];  // max|in| < 2^14 / s_in
      if (g.in_absmax) { float const t = __uint_as_float(*g.in_absmax); if (t > 0.0f) { in_bound = fminf(in_bound, t); } }
      float const res_bound = g.res_absmax ? __uint_as_float(*g.res_absmax) : 0.0f;
      s_out = scale_from_absmax_bits(__float_as_uint((__ldg(g.w_l1max) * in_bound + bm + res_bound) * 1.01f));  // same bound as igemm_out_scale (igemm.cuh)
      if (blockIdx.x == 0 && etid == 0) { g.out16_scale2[0] = s_out; g.out16_scale2[1] = 1.0f / s_out; }
    }
    uint32_t const t_begin = g.ts ? static_cast<uint32_t>(clock()) : 0u;  // 32-bit cycle counters here
    uint32_t w_acc = 0, t_drain = 0, t_store = 0;
    uint32_t const U = static_cast<uint32_t>(prm.n_tiles) * static_cast<uint32_t>(prm.ukb);
    Sk4Work work(prm, pair, n_pairs);
    int tile, u0, u1;
    int gc = 0, si = 0;
    unsigned int const my_slot = static_cast<unsigned int>(pair) * 2u + cta_rank;
    while (work.next(tile, u0, u1)) {
      int const mt = tile / q_tiles, nt = tile - mt * q_tiles;
      int const m0 = (mt * 2 + static_cast<int>(cta_rank)) * IGEMM_BM, n0 = nt * BN;
      bool const is_head = (u0 == 0), is_whole = is_head && (u1 == prm.ukb);
      int const nchunks = (u1 - u0 + chunk_u - 1) / chunk_u;
      int const prow = m0 + row;
      float *bias_t = bias_s + (si & 1) * BN;
      if (is_head) {  // this tile's bias, staged in the buffer of the segment's parity (the other one may still be read by a slower warp's stores)
        for (int j = etid; j < BN; j += 256) { bias_t[j] = (g.has_bias && kEpi != SK4_EPI_SWAPPED && (n0 + j) < g.q_rows) ? __ldg(g.bias + n0 + j) : 0.0f; }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // epilogue warps only
      }
      float acc[HN];
      {
        float const *res_row = nullptr;
        if (kEpi == SK4_EPI_FULL && g.res && is_head) {  // (host: a residual input only on pixel-major launches)
          Sk4RowGeom const rg = sk4_row_geom(prm, prow, halo);
          if (rg.valid) { res_row = g.res + (static_cast<long long>(rg.img) * g.out_chans + n0 + c0) * g.out_hw + rg.pix; }
        }
        igemm_acc_init<HN>(acc, res_row, g.out_hw, g.q_rows - n0 - c0, inv_recip);
      }
      for (int c = 0; c < nchunks; ++c, ++gc) {
        int const buf = gc & 1;
        uint32_t t0 = 0;
        if (g.ts) { t0 = static_cast<uint32_t>(clock()); mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); uint32_t const t1 = static_cast<uint32_t>(clock()); w_acc += t1 - t0; t0 = t1; } else { mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); }
        tc_fence_after();
        sk4_drain<HN>(acc, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBufCols + c0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) { mbar_arrive(&tmem_empty_bar[buf]); } else { mbar_arrive_remote(&tmem_empty_bar[buf], 0); } }
        if (g.ts) { t_drain += static_cast<uint32_t>(clock()) - t0; }
      }
      uint32_t const t_st0 = g.ts ? static_cast<uint32_t>(clock()) : 0u;
      if (kPlanes == 2) {  // the last chunk's commit also covered every cross-term MMA of this segment
        tc_fence_after();
        sk4_drain<HN>(acc, tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (2 + (si & 1)) * kBufCols + c0);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) { mbar_arrive(&x_empty_bar[si & 1]); } else { mbar_arrive_remote(&x_empty_bar[si & 1], 0); } }
      }
      ++si;
      if (!is_head) {
        // ---- stream-K contributor: raw partial accumulators -> workspace slot of this CTA, then raise its flag ----
        float *w = prm.sk_ws + (static_cast<size_t>(my_slot) * BN + c0) * 128 + row;
#pragma unroll
        for (int j = 0; j < HN; ++j) { __stcg(w + j * 128, acc[j]); }
        __threadfence();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (etid == 0) { sk4_flag_raise(prm.sk_flags + my_slot); }
        if (g.ts) { t_store += static_cast<uint32_t>(clock()) - t_st0; }
        continue;
      }
      if (!is_whole) {
        // ---- stream-K finisher: this pair owns the tile's first k-block; the pairs after it hold the rest, each in its own slot ----
        uint32_t const tile_end = static_cast<uint32_t>(tile + 1) * static_cast<uint32_t>(prm.ukb);
        for (int pp = pair + 1; pp < n_pairs && Sk4Work::range_begin(U, pp, n_pairs) < tile_end; ++pp) {
          unsigned int const slot = static_cast<unsigned int>(pp) * 2u + cta_rank;
          if (etid == 0) { sk4_flag_wait_and_reset(prm.sk_flags + slot); }
          asm volatile("bar.sync 1, 256;" ::: "memory");
          float const *w = prm.sk_ws + (static_cast<size_t>(slot) * BN + c0) * 128 + row;
#pragma unroll
          for (int j0 = 0; j0 < HN; j0 += 16) {  // 16 loads in flight per thread, 8 warps: an L2 round trip per 16 columns
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = __ldcg(w + (j0 + j) * 128); }
#pragma unroll
            for (int j = 0; j < 16; ++j) { acc[j0 + j] += v[j]; }
          }
        }
      }
      // ---- final epilogue of the tile ----
      if (kEpi != SK4_EPI_SWAPPED) {
        Sk4RowGeom const rg = sk4_row_geom(prm, prow, halo);  // where this thread's row lives in the output (computed here, not held across the main loop)
        if (rg.valid && !(g.debug & 4)) {
          float *o = g.out + (static_cast<long long>(rg.img) * g.out_chans + n0 + c0) * g.out_hw + rg.pix;
          amax = fmaxf(amax, igemm_store_row<HN>(acc, inv, bias_t + c0, floor_v, o, g.out_hw, g.q_rows - n0 - c0));
          if (kEpi == SK4_EPI_FULL && g.out16) {
            long long const orow = prm.o16_Wp ? (static_cast<long long>(rg.img) * prm.o16_Hp + rg.oy + prm.o16_py) * prm.o16_Wp + rg.ox + prm.o16_px : static_cast<long long>(rg.img) * g.out_hw + rg.pix;
            long long const o16 = orow * g.out16_pitch + n0 + c0;
            if (g.w_l1max) { igemm_store_row_split16<HN>(acc, inv, bias_t + c0, floor_v, s_out, g.out16 + o16, g.out16_lo ? g.out16_lo + o16 : nullptr, g.q_rows - n0 - c0); }
            else { igemm_store_row_bf16<HN>(acc, inv, bias_t + c0, floor_v, g.out16 + o16, g.q_rows - n0 - c0); }
          }
        }
      } else if (prow < g.p_rows && !(g.debug & 4)) {  // row = out chan, columns = pixels (inner-product-shaped layers)
        float const b = g.has_bias ? __ldg(g.bias + prow) : 0.0f;
        int const p0 = n0 + c0;  // first pixel of this thread's columns
        int const img0 = p0 / g.out_hw;
        int px = p0 - img0 * g.out_hw;
        long long off = (static_cast<long long>(img0) * g.out_chans + prow) * g.out_hw + px;
        long long const img_step = static_cast<long long>(g.out_chans) * g.out_hw - (g.out_hw - 1);  // last pixel of an image -> first of the next
#pragma unroll
        for (int j = 0; j < HN; ++j) {
          if (p0 + j < g.q_rows) {
            float const v = fmaxf(fmaf(acc[j], inv, b), floor_v);
            amax = fmaxf(amax, fabsf(v));
            g.out[off] = v;
          }
          if (++px == g.out_hw) { px = 0; off += img_step; } else { off += 1; }
        }
      }
      if (g.ts) { t_store += static_cast<uint32_t>(clock()) - t_st0; }
    }
    if (g.ts && leader && warp_id == 2 && lane == 0) { long long *ts = g.ts + pair * 16; ts[8] = static_cast<uint32_t>(clock()) - t_begin; ts[9] = w_acc; ts[10] = t_drain; ts[11] = t_store; ts[12] = si; }
    if (g.out_absmax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
      if (lane == 0 && amax > 0.0f) { atomicMax(g.out_absmax, __float_as_uint(amax)); }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be arriving on the leader's barriers / the MMA may still read the peer's shared memory
  if (warp_id == 1) { tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base); }
}

}  // namespace b200
