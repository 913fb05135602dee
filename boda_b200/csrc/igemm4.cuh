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
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + (leader CTA) MMA issuer, warps 2..5 = epilogue.
#pragma once
#include "igemm.cuh"

namespace b200 {

constexpr int SK4_MAX_A_STAGES = 4, SK4_MAX_B_STAGES = 12;
constexpr int SK4_BAR_BYTES = 2048;  // barriers in the first 512 B, reduction scratch at +768, two staged bias vectors (segment parity) at +1024

struct Sk4Params {
  IgemmParams g;     // extents, im2col addressing, epilogue, scales, layout-transform-elimination outputs: exactly as in igemm.cuh
  int p_mode;        // P operand: 0 = 2-d tiled {k, row}, 1 = im2col TMA (one activation tile per k-block), 2 = halo (one tile per channel block)
  int taps;          // halo: KH*KW (k-blocks are ordered channel block major, tap minor); 1 otherwise
  int Wp, HpWp, OH, OW;                 // halo: padded row pitch, virtual pixels per image, valid output rows / columns
  int halo_rows, a_loads, a_box_rows;   // halo: activation rows per stage = a_loads boxes of a_box_rows
  int a_stages, b_stages;               // ring depths (modes 0/1: one ring, the P tiles ride on the Q stages' barriers)
  int sk;            // 1 = stream-K ranges, 0 = whole tiles, pair p takes tiles p, p + #pairs, ...
  int n_tiles, ukb;  // tiles; stage-units (kKb k-blocks) per tile
  float *sk_ws;      // stream-K partials: [CTA slot = pair*2 + rank][BN][128] fp32, raw accumulator units
  unsigned int *sk_flags;  // one per CTA slot: raised by the contributor, reset by the finisher
  int out_w;         // output width, for the padded destination planes
  int o16_Hp, o16_Wp, o16_py, o16_px;  // destination planes in the shared-padding layout of a halo-mode consumer (o16_Wp = 0: plain pixel-major)
};

template <int BN, int kPlanes>
struct Sk4Cfg {
  static constexpr int kKb = (kPlanes == 1) ? 2 : 1;  // k-blocks per stage: 8 MMAs per barrier round trip in the single-plane modes (see igemm2.cuh)
  static constexpr int kBQ = BN / 2;                  // filter rows per CTA
  static constexpr uint32_t kBSlot = kBQ * 128, kBStage = 2 * kBSlot;  // two slots: hi / lo planes of one k-block, or two k-blocks
  static constexpr uint32_t kPSlot = IGEMM_BM * 128, kPStage = 2 * kPSlot;  // modes 0/1: the P tiles of a stage
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

template <int BN, int kPlanes>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_sk4_kernel(const __grid_constant__ CUtensorMap p_hi_map, const __grid_constant__ CUtensorMap p_lo_map,
                 const __grid_constant__ CUtensorMap q_hi_map, const __grid_constant__ CUtensorMap q_lo_map, const Sk4Params prm) {
  using Cfg = Sk4Cfg<BN, kPlanes>;
  constexpr int kKb = Cfg::kKb;
  constexpr uint32_t kBSlot = Cfg::kBSlot, kBStage = Cfg::kBStage, kPSlot = Cfg::kPSlot, kBufCols = Cfg::kBufCols;
  IgemmParams const &g = prm.g;
  bool const halo = (prm.p_mode == 2);
  int const a_stages = prm.a_stages, b_stages = prm.b_stages;
  uint32_t const a_plane = halo ? static_cast<uint32_t>(prm.halo_rows) * 128u : kPSlot;
  uint32_t const a_stage = halo ? kPlanes * a_plane : Cfg::kPStage;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_ring = smem;
  uint8_t *b_ring = a_ring + a_stages * a_stage;
  uint8_t *bar_mem = b_ring + b_stages * kBStage;
  uint64_t *a_full = reinterpret_cast<uint64_t *>(bar_mem);  // leader's copies collect both CTAs' bytes
  uint64_t *a_empty = a_full + SK4_MAX_A_STAGES;
  uint64_t *b_full = a_empty + SK4_MAX_A_STAGES;
  uint64_t *b_empty = b_full + SK4_MAX_B_STAGES;
  uint64_t *tmem_full_bar = b_empty + SK4_MAX_B_STAGES;  // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;          // [2], leader's copies collect both CTAs' epilogue warps
  uint64_t *x_empty_bar = tmem_empty_bar + 2;            // [2], cross-term accumulators (fp32-parity mode)
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(x_empty_bar + 2);
  float *bias_s = reinterpret_cast<float *>(bar_mem + 1024);  // [2][BN]

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  uint32_t const cta_rank = cluster_ctarank();
  bool const leader = (cta_rank == 0);
  int const n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  int const q_tiles = g.q_tiles;
  int const nkb = g.kblks_total;
  int const chunk_u = max(1, g.chunk_kblks / kKb);  // stage-units per accumulation chunk
  int const taps = prm.taps, cblks = g.cblks;

  if (warp_id == 0) {  // barrier setup spread over the lanes (one thread doing ~40 mbarrier.init in a row is a microsecond of prologue)
    if (lane == 0) {
      tma_prefetch_desc(&p_hi_map);
      tma_prefetch_desc(&q_hi_map);
      if (kPlanes == 2) { tma_prefetch_desc(&p_lo_map); tma_prefetch_desc(&q_lo_map); }
    }
    if (lane < a_stages) { mbar_init(&a_full[lane], 2); mbar_init(&a_empty[lane], 1); }
    if (lane < b_stages) { mbar_init(&b_full[lane], 2); mbar_init(&b_empty[lane], 1); }
    if (lane < 2) { mbar_init(&tmem_full_bar[lane], 1); mbar_init(&tmem_empty_bar[lane], 8); mbar_init(&x_empty_bar[lane], 8); }  // 4 epilogue warps x 2 CTAs
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
    long long w_empty = 0;
    Sk4Work work(prm, pair, n_pairs);
    int tile, u0, u1;
    int sb = 0; uint32_t b_par = 1;  // next B stage, parity to wait for on its empty barrier (fresh barriers pass parity 1)
    int sa = 0; uint32_t a_par = 1;  // halo: next A stage
    while (!(g.debug & 1) && work.next(tile, u0, u1)) {
      int const mt = tile / q_tiles, nt = tile - mt * q_tiles;
      int const m0 = (mt * 2 + static_cast<int>(cta_rank)) * IGEMM_BM;
      int const q_row0 = nt * BN + static_cast<int>(cta_rank) * Cfg::kBQ;
      int const kb0 = u0 * kKb, kb1 = min(u1 * kKb, nkb);
      // ---- per-segment operand addressing state (integer divisions here only: once per segment, never per k-block) ----
      int img = 0, h_base = 0, w_base = 0, cb = 0, kx = 0, ky = 0;  // im2col
      int c = 0, t = 0, c_next = 0, c_last = 0;                     // halo: channel block / tap of the next k-block; next A tile to request; last one of the segment
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
        c_next = c; c_last = (kb1 - 1) / taps;
      }
      auto issue_a = [&](int cc) {  // halo tile of channel block cc: rows [m0, m0 + halo_rows) of the padded activation matrix
        if (elect_one_sync()) {
          if (leader) { mbar_expect_tx(&a_full[sa], 2u * a_stage); } else { mbar_arrive_remote(&a_full[sa], 0); }
          uint8_t *dst = a_ring + sa * a_stage;
          for (int l = 0; l < prm.a_loads; ++l) {
            uint8_t *d = dst + l * prm.a_box_rows * 128;
            int const row = m0 + l * prm.a_box_rows;
            tma_load_2d_2sm(d, &p_hi_map, &a_full[sa], cc * IGEMM_BK, row);
            if (kPlanes == 2) { tma_load_2d_2sm(d + a_plane, &p_lo_map, &a_full[sa], cc * IGEMM_BK, row); }
          }
        }
        __syncwarp();
        if (++sa == a_stages) { sa = 0; a_par ^= 1; }
      };
      for (int kb = kb0; kb < kb1; kb += kKb) {
        int const nh = min(kKb, kb1 - kb);
        if (halo) {
          // the tile of the channel block about to be multiplied must be on its way (blocking); later blocks of this segment are requested
          // as soon as a stage is free, without ever blocking the filter stream behind them
          int const c_need = (nh == 2 && t == taps - 1) ? min(c + 1, c_last) : c;
          while (c_next <= c_last) {
            if (c_next <= c_need) { mbar_wait(&a_empty[sa], a_par); }
            else if (!mbar_test_wait(&a_empty[sa], a_par)) { break; }  // (test_wait: try_wait would suspend the producer for its hardware time limit)
            issue_a(c_next);
            ++c_next;
          }
        }
        if (g.ts) { long long const t0 = clock64(); mbar_wait(&b_empty[sb], b_par); w_empty += clock64() - t0; } else { mbar_wait(&b_empty[sb], b_par); }
        bool const issue = elect_one_sync();
        uint32_t const b_bytes = (kPlanes == 2 ? 2u : static_cast<uint32_t>(nh)) * kBSlot, p_bytes = halo ? 0u : (kPlanes == 2 ? 2u : static_cast<uint32_t>(nh)) * kPSlot;
        if (issue) {
          if (leader) { mbar_expect_tx(&b_full[sb], 2u * (b_bytes + p_bytes)); } else { mbar_arrive_remote(&b_full[sb], 0); }
        }
        uint8_t *bst = b_ring + sb * kBStage;
        uint8_t *pst = a_ring + sb * a_stage;  // modes 0/1: the P slots of this stage
#pragma unroll
        for (int j = 0; j < kKb; ++j) {
          if (j < nh) {
            int const kbj = kb + j;
            if (issue) {
              uint8_t *q_hi = bst + (kPlanes == 2 ? 0 : j) * kBSlot, *q_lo = bst + kBSlot;
              int qc0, qc1;
              if (halo) { qc0 = 0; qc1 = (t * cblks + c) * g.q_kb_rows + q_row0; }  // packed filters: k-block (tap * cblks + block), rows = out chans
              else { qc0 = g.q_kb_rows ? 0 : kbj * IGEMM_BK; qc1 = q_row0 + kbj * g.q_kb_rows; }
              tma_load_2d_2sm(q_hi, &q_hi_map, &b_full[sb], qc0, qc1);
              if (kPlanes == 2) { tma_load_2d_2sm(q_lo, &q_lo_map, &b_full[sb], qc0, qc1); }
              if (!halo) {
                uint8_t *p_hi = pst + (kPlanes == 2 ? 0 : j) * kPSlot, *p_lo = pst + kPSlot;
                if (prm.p_mode == 1) {
                  tma_load_im2col_4d_2sm(p_hi, &p_hi_map, &b_full[sb], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky);
                  if (kPlanes == 2) { tma_load_im2col_4d_2sm(p_lo, &p_lo_map, &b_full[sb], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky); }
                } else {
                  int const pc0 = g.p_kb_rows ? 0 : kbj * IGEMM_BK, pc1 = m0 + kbj * g.p_kb_rows;
                  tma_load_2d_2sm(p_hi, &p_hi_map, &b_full[sb], pc0, pc1);
                  if (kPlanes == 2) { tma_load_2d_2sm(p_lo, &p_lo_map, &b_full[sb], pc0, pc1); }
                }
              }
            }
            if (halo) { if (++t == taps) { t = 0; ++c; } }
            else if (prm.p_mode == 1) { if (++cb == cblks) { cb = 0; if (++kx == g.kw) { kx = 0; ++ky; } } }
          }
        }
        __syncwarp();
        if (++sb == b_stages) { sb = 0; b_par ^= 1; }
      }
    }
    if (g.ts && leader && lane == 0) { long long *ts = g.ts + pair * 16; ts[0] = clock64() - t_begin; ts[1] = w_empty; }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (leader CTA only, one elected thread for the pair) =====================
    // The tensor pipe queues only ~2 MMAs, so every instruction the issuing thread executes between two tcgen05.mma is tensor-pipe idle time
    // (DESIGN section 4): all waits and all loop state of a stage are handled by the whole warp BEFORE the elected lane's block, which then holds
    // nothing but the stage's MMAs and commits.
    if (leader) {
      long long const t_begin = g.ts ? clock64() : 0;
      long long w_full = 0, w_tmem = 0, w_afull = 0, t_first_full = 0;
      uint32_t const idesc = g.idesc;  // M = 256 (the pair), N = BN
      int const ksteps_full = (g.debug & 2) ? 0 : IGEMM_BK / IGEMM_UMMA_K, ksteps_last = (g.debug & 2) ? 0 : g.ksteps_last;
      uint32_t const a_base = smem_u32(a_ring), b_base = smem_u32(b_ring);
      bool const no_tma = (g.debug & 1) != 0;
      Sk4Work work(prm, pair, n_pairs);
      int tile, u0, u1;
      int sb = 0; uint32_t b_par = 0;
      int sa = 0; uint32_t a_par = 0;
      uint32_t a_cur = a_base;          // halo: shared-memory address of the current activation tile (stage sa)
      int gc = 0, si = 0, n_stage = 0;  // accumulation chunks, segments and stages done so far
      while (work.next(tile, u0, u1)) {
        int const kb1 = min(u1 * kKb, nkb);
        uint32_t const tmem_x = tmem_base + (2 + (si & 1)) * kBufCols;
        if (kPlanes == 2) { mbar_wait(&x_empty_bar[si & 1], ((si >> 1) & 1) ^ 1); }  // the epilogue has read the cross terms of segment si - 2
        int c = 0, t = 0, kxx = 0, tap_row = 0, kb_in_grp = 0;
        bool a_ready = false;  // halo: the current channel block's tile has been waited for
        {
          int const kb0 = u0 * kKb;
          if (halo) { c = kb0 / taps; t = kb0 - c * taps; int const kyy = t / g.kw; kxx = t - kyy * g.kw; tap_row = kyy * prm.Wp + kxx; }
          else { kb_in_grp = kb0 % g.kb_mod; }
        }
        uint32_t x_acc = 0u;  // the segment's first MMA overwrites the cross-term accumulator
        for (int u = u0; u < u1;) {
          int const buf = gc & 1;
          if (g.ts) { long long const t0 = clock64(); mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); w_tmem += clock64() - t0; } else { mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); }
          uint32_t const tmem_d = tmem_base + buf * kBufCols;
          int const u_end = min(u + chunk_u, u1);
          uint32_t main_acc = 0u;  // the chunk's first MMA overwrites the main accumulator
          for (; u < u_end; ++u, ++n_stage) {
            int const kb = u * kKb;
            int const nh = min(kKb, kb1 - kb);
            // ---- per-k-block operands and bookkeeping of this stage, by the whole warp ----
            uint32_t p_addr[kKb];
            int nk[kKb];
            bool a_done[kKb];
            uint32_t a_rel[kKb];  // halo: the activation stage k-block j releases (index), when a_done[j]
#pragma unroll
            for (int j = 0; j < kKb; ++j) { p_addr[j] = 0; nk[j] = 0; a_done[j] = false; a_rel[j] = 0; }
            if (halo) {
#pragma unroll
              for (int j = 0; j < kKb; ++j) {
                if (j < nh) {
                  if (!a_ready) {  // first k-block of a channel block (or of the segment): its tile must have landed
                    if (no_tma) {} else if (g.ts) { long long const t0 = clock64(); mbar_wait(&a_full[sa], a_par); w_afull += clock64() - t0; } else { mbar_wait(&a_full[sa], a_par); }
                    a_ready = true;
                  }
                  p_addr[j] = a_cur + ((g.debug & 32) ? 0u : static_cast<uint32_t>(tap_row) * 128u);  // (debug bit 5: every tap reads tap 0 -- 1024-byte aligned descriptors, timing experiments)
                  nk[j] = (c == cblks - 1) ? ksteps_last : ksteps_full;
                  a_done[j] = (t == taps - 1) || (kb + j == kb1 - 1);  // last tap of the block, or the segment ends inside it: the tile is released
                  a_rel[j] = static_cast<uint32_t>(sa);
                  if (a_done[j]) { a_ready = false; if (++sa == a_stages) { sa = 0; a_par ^= 1; a_cur = a_base; } else { a_cur += a_stage; } }
                  if (++t == taps) { t = 0; kxx = 0; tap_row = 0; ++c; }
                  else if (++kxx == g.kw) { kxx = 0; tap_row += prm.Wp - (g.kw - 1); } else { ++tap_row; }
                }
              }
            } else {
              uint32_t const pst = a_base + static_cast<uint32_t>(sb) * Cfg::kPStage;
#pragma unroll
              for (int j = 0; j < kKb; ++j) {
                if (j < nh) {
                  p_addr[j] = pst + (kPlanes == 2 ? 0u : static_cast<uint32_t>(j)) * kPSlot;
                  nk[j] = ksteps_full;
                  if (++kb_in_grp == g.kb_mod) { kb_in_grp = 0; nk[j] = ksteps_last; }
                }
              }
            }
            if (no_tma) {}  // experiments: MMA on whatever the shared memory holds
            else if (g.ts) { long long const t0 = clock64(); mbar_wait(&b_full[sb], b_par); long long const t1 = clock64(); if (n_stage == 0) { t_first_full = t1 - t_begin; } else { w_full += t1 - t0; } }
            else { mbar_wait(&b_full[sb], b_par); }
            tc_fence_after();
            uint32_t const bst = b_base + static_cast<uint32_t>(sb) * kBStage;
            bool const chunk_end = (u + 1 == u_end);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < kKb; ++j) {
                if (j < nh) {
                  uint32_t const q_addr = bst + (kPlanes == 2 ? 0u : static_cast<uint32_t>(j)) * kBSlot;
                  issue_kblock<kPlanes, true>(tmem_d, tmem_x, sw128_desc_lo(p_addr[j]), sw128_desc_lo(p_addr[j] + a_plane), sw128_desc_lo(q_addr), sw128_desc_lo(q_addr + kBSlot), idesc,
                                              (j == 0) ? main_acc : 1u, (j == 0) ? x_acc : 1u, nk[j]);
                  if (a_done[j]) { umma_commit_2sm(&a_empty[a_rel[j]], 0x3); }
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
    // ===================== epilogue warps (each CTA: its own 128 rows) =====================
    int const q = warp_id & 3;
    int const row = q * 32 + lane;
    float const inv = g.p_scale[1] * g.q_scale[1], inv_recip = g.p_scale[0] * g.q_scale[0];
    float const floor_v = g.relu ? 0.0f : -INFINITY;
    float amax = 0.0f;
    float s_out = 1.0f;
    if (g.out16 && g.w_l1max) {  // scale of the consumer's fp16 planes from the output bound (identical in every CTA); CTA 0 publishes it
      s_out = igemm_out_scale(g, reinterpret_cast<float *>(bar_mem + 768), row);
      if (blockIdx.x == 0 && row == 0) { g.out16_scale2[0] = s_out; g.out16_scale2[1] = 1.0f / s_out; }
    }
    uint32_t const t_begin = g.ts ? static_cast<uint32_t>(clock()) : 0u;  // 32-bit cycle counters here: the epilogue threads sit at the register limit
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
        for (int j = row; j < BN; j += 128) { bias_t[j] = (g.has_bias && !g.swapped && (n0 + j) < g.q_rows) ? __ldg(g.bias + n0 + j) : 0.0f; }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      }
      float acc[BN];
      {
        float const *res_row = nullptr;
        if (g.res && is_head) {  // (host: a residual input only on pixel-major launches)
          Sk4RowGeom const rg = sk4_row_geom(prm, prow, halo);
          if (rg.valid) { res_row = g.res + (static_cast<long long>(rg.img) * g.out_chans + n0) * g.out_hw + rg.pix; }
        }
        igemm_acc_init<BN>(acc, res_row, g.out_hw, g.q_rows - n0, inv_recip);
      }
      for (int c = 0; c < nchunks; ++c, ++gc) {
        int const buf = gc & 1;
        uint32_t t0 = 0;
        if (g.ts) { t0 = static_cast<uint32_t>(clock()); mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); uint32_t const t1 = static_cast<uint32_t>(clock()); w_acc += t1 - t0; t0 = t1; } else { mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); }
        tc_fence_after();
        uint32_t const taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kBufCols;
#pragma unroll
        for (int j0 = 0; j0 < BN; j0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + j0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) { mbar_arrive(&tmem_empty_bar[buf]); } else { mbar_arrive_remote(&tmem_empty_bar[buf], 0); } }
        if (g.ts) { t_drain += static_cast<uint32_t>(clock()) - t0; }
      }
      uint32_t const t_st0 = g.ts ? static_cast<uint32_t>(clock()) : 0u;
      if (kPlanes == 2) {  // the last chunk's commit also covered every cross-term MMA of this segment
        tc_fence_after();
        uint32_t const xaddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (2 + (si & 1)) * kBufCols;
#pragma unroll
        for (int j0 = 0; j0 < BN; j0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(xaddr + j0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (leader) { mbar_arrive(&x_empty_bar[si & 1]); } else { mbar_arrive_remote(&x_empty_bar[si & 1], 0); } }
      }
      ++si;
      if (!is_head) {
        // ---- stream-K contributor: raw partial accumulators -> workspace slot of this CTA, then raise its flag ----
        float *w = prm.sk_ws + (static_cast<size_t>(my_slot) * BN) * 128 + row;
#pragma unroll
        for (int j = 0; j < BN; ++j) { __stcg(w + j * 128, acc[j]); }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row == 0) { sk4_flag_raise(prm.sk_flags + my_slot); }
        if (g.ts) { t_store += static_cast<uint32_t>(clock()) - t_st0; }
        continue;
      }
      if (!is_whole) {
        // ---- stream-K finisher: this pair owns the tile's first k-block; the pairs after it hold the rest, each in its own slot ----
        uint32_t const tile_end = static_cast<uint32_t>(tile + 1) * static_cast<uint32_t>(prm.ukb);
        for (int pp = pair + 1; pp < n_pairs && Sk4Work::range_begin(U, pp, n_pairs) < tile_end; ++pp) {
          unsigned int const slot = static_cast<unsigned int>(pp) * 2u + cta_rank;
          if (row == 0) { sk4_flag_wait_and_reset(prm.sk_flags + slot); }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float const *w = prm.sk_ws + (static_cast<size_t>(slot) * BN) * 128 + row;
#pragma unroll
          for (int j0 = 0; j0 < BN; j0 += 32) {  // 32 loads in flight per thread (the TMEM drain's registers are free by now): an L2 round trip per 32 columns
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] = __ldcg(w + (j0 + j) * 128); }
#pragma unroll
            for (int j = 0; j < 32; ++j) { acc[j0 + j] += v[j]; }
          }
        }
      }
      // ---- final epilogue of the tile ----
      if (!g.swapped) {
        Sk4RowGeom const rg = sk4_row_geom(prm, prow, halo);  // where this thread's row lives in the output (computed here, not held across the main loop)
        bool const valid = rg.valid;
        int const img = rg.img, pix = rg.pix, oy = rg.oy, ox = rg.ox;
        if (valid && !(g.debug & 4)) {
          float *o = g.out + (static_cast<long long>(img) * g.out_chans + n0) * g.out_hw + pix;
          amax = fmaxf(amax, igemm_store_row<BN>(acc, inv, bias_t, floor_v, o, g.out_hw, g.q_rows - n0));
          if (g.out16) {
            long long const orow = prm.o16_Wp ? (static_cast<long long>(img) * prm.o16_Hp + oy + prm.o16_py) * prm.o16_Wp + ox + prm.o16_px : static_cast<long long>(img) * g.out_hw + pix;
            long long const o16 = orow * g.out16_pitch + n0;
            if (g.w_l1max) { igemm_store_row_split16<BN>(acc, inv, bias_t, floor_v, s_out, g.out16 + o16, g.out16_lo ? g.out16_lo + o16 : nullptr, g.q_rows - n0); }
            else { igemm_store_row_bf16<BN>(acc, inv, bias_t, floor_v, g.out16 + o16, g.q_rows - n0); }
          }
        }
      } else if (prow < g.p_rows && !(g.debug & 4)) {  // row = out chan, columns = pixels (inner-product-shaped layers)
        float const b = g.has_bias ? __ldg(g.bias + prow) : 0.0f;
        int const img0 = n0 / g.out_hw;
        int px = n0 - img0 * g.out_hw;
        long long off = (static_cast<long long>(img0) * g.out_chans + prow) * g.out_hw + px;
        long long const img_step = static_cast<long long>(g.out_chans) * g.out_hw - (g.out_hw - 1);  // last pixel of an image -> first of the next
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          if (n0 + j < g.q_rows) {
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
