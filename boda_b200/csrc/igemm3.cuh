// igemm3.cuh -- "tap-reuse" contraction kernel for stride-1 convolutions with a KH x KW > 1 window (the 3x3 / 5x5 layers that
// dominate AlexNet-ng / GoogLeNet / NiN; replaces the same CUCL functions as igemm.cuh: conv / tconv, test/rtc/{conv,tconv}.cucl).
//
// Why: the im2col kernels (igemm.cuh / igemm2.cuh) fetch the activation tile once PER FILTER TAP, so a 3x3 layer pulls every activation
// byte 9 times (5x5: 25 times) from L2 into shared memory, and the L2 -> SM path (not the tensor pipe) bounds the kernel (ncu:
// lts throughput ~75 % while the tensor pipe idles). Here the activations live in HBM as a zero-padded NHWC image with SHARED padding,
//     act[n][yp][xp][chan],  yp < Hp = H + pad_y,  xp < Wp = W + pad_x,  pixel (y,x) stored at (y + pad_y, x + pad_x),
// (the right padding of one image row is the left padding of the next, the bottom padding of one image the top padding of the next,
// rows past the end are zero-filled by TMA), and the GEMM rows are ALL "virtual pixels" v = (n*Hp + yp)*Wp + xp. For a stride-1
// convolution, output (n,oy,ox) is virtual pixel v = (n*Hp + oy)*Wp + ox and its tap (ky,kx) reads activation row v + ky*Wp + kx: a
// CONSTANT row offset per tap. So one shared-memory "halo" tile of 128 + (KH-1)*Wp + (KW-1) consecutive activation rows (x 64 channels)
// serves every tap of a 64-channel block: the UMMA shared-memory descriptor of tap (ky,kx) simply starts (ky*Wp + kx) rows (x 128 B) into
// the halo (a K-major SWIZZLE_128B descriptor may start at any 128-byte row of a TMA-written tile: the swizzle is a function of the
// shared-memory address bits; probed on hardware, tools/probe_umma_shift.cu). Virtual pixels with yp >= OH or xp >= OW are computed and
// dropped: (Hp*Wp)/(OH*OW) = 1.04 .. 1.16 x the MMA work for the layers above, against ~KH*KW x less activation traffic.
//
// Pipeline: two TMA rings -- A (one halo tile per 64-channel block, `a_stages` deep) and B (one BN x 64 filter tile per (tap, block),
// `b_stages` deep) -- one producer thread, one MMA-issuing thread, four epilogue warps; accumulator ping-pong in TMEM, periodic draining
// and the fp32-parity hi/lo split exactly as in igemm.cuh. k2 = true is the CTA-pair (cta_group::2) version: 256 virtual pixels per pair,
// each CTA holds its own halo and half of every filter tile.
#pragma once
#include "igemm.cuh"

namespace b200 {

struct TapsParams {
  int m_rows;      // virtual pixels that may hold an output: (N-1)*Hp*Wp + (OH-1)*Wp + OW
  int q_rows;      // out chans
  int cblks;       // 64-channel blocks
  int w_kb_rows;   // packed filters are k-block-major [tap*cblks + block][out_chan (padded to w_kb_rows)][64]
  int ksteps_last; // 16-wide k-steps of real data in the last channel block (1..4): zero-padded channels are not multiplied
  int taps, kw;    // KH*KW, KW
  int Wp, HpWp;    // padded row pitch, virtual pixels per image
  int OH, OW;
  int halo_rows;   // activation rows per A stage = a_loads * a_box_rows (multiple of 8)
  int a_loads, a_box_rows;
  int a_stages, b_stages;
  int chunk_kblks; // drain TMEM every this many (tap, block) steps
  int out_chans, out_hw;
  int relu, has_bias;
  float *out;
  float const *bias;
  float const *p_scale, *q_scale;
  unsigned int *out_absmax;
  uint32_t idesc;
  int debug;  // timing experiments only (results are garbage): bit 0 = no TMA loads, bit 1 = no MMA issue (barrier traffic only)
  long long *ts;  // experiments: CTA (0,0) writes clock64() stamps of its phases here (null = off)
};
#define TAPS_STAMP(i) do { if (prm.ts && blockIdx.x == 0 && blockIdx.y == 0) { prm.ts[i] = clock64(); } } while (0)

constexpr int TAPS_MAX_A_STAGES = 4, TAPS_MAX_B_STAGES = 8;
constexpr int TAPS_BAR_BYTES = 2048;  // barriers in the first 512 B, staged bias at +1024

template <int BN, int kPlanes, bool k2>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_taps_kernel(const __grid_constant__ CUtensorMap a_hi_map, const __grid_constant__ CUtensorMap a_lo_map,
                  const __grid_constant__ CUtensorMap w_hi_map, const __grid_constant__ CUtensorMap w_lo_map, const TapsParams prm) {
  constexpr int BQ = k2 ? BN / 2 : BN;  // filter rows held by this CTA
  constexpr uint32_t kBBytes = BQ * 128, kBStage = kPlanes * kBBytes;
  constexpr uint32_t kColsNeeded = (kPlanes == 2 ? 3 : 2) * tmem_buf_cols(BN);
  constexpr uint32_t kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128 : kColsNeeded <= 256 ? 256 : 512;
  uint32_t const a_plane = static_cast<uint32_t>(prm.halo_rows) * 128, a_stage = kPlanes * a_plane;
  int const a_stages = prm.a_stages, b_stages = prm.b_stages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_ring = smem;
  uint8_t *b_ring = a_ring + a_stages * a_stage;
  uint8_t *bar_mem = b_ring + b_stages * kBStage;
  uint64_t *a_full = reinterpret_cast<uint64_t *>(bar_mem);  // (k2: the leader's copies collect both CTAs' bytes)
  uint64_t *a_empty = a_full + TAPS_MAX_A_STAGES;
  uint64_t *b_full = a_empty + TAPS_MAX_A_STAGES;
  uint64_t *b_empty = b_full + TAPS_MAX_B_STAGES;
  uint64_t *tmem_full_bar = b_empty + TAPS_MAX_B_STAGES;   // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
  float *bias_s = reinterpret_cast<float *>(bar_mem + 1024);

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) { TAPS_STAMP(0); }
  uint32_t const cta_rank = k2 ? cluster_ctarank() : 0u;
  bool const leader = (cta_rank == 0);
  int const m0 = blockIdx.x * IGEMM_BM;  // this CTA's first virtual pixel
  int const n0 = blockIdx.y * BN;
  int const taps = prm.taps, cblks = prm.cblks;
  int const nkb = cblks * taps;
  int const chunk = prm.chunk_kblks;
  int const nchunks = (nkb + chunk - 1) / chunk;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&a_hi_map);
    tma_prefetch_desc(&w_hi_map);
    if (kPlanes == 2) { tma_prefetch_desc(&a_lo_map); tma_prefetch_desc(&w_lo_map); }
    uint32_t const n_prod = k2 ? 2 : 1;
    for (int i = 0; i < a_stages; ++i) { mbar_init(&a_full[i], n_prod); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < b_stages; ++i) { mbar_init(&b_full[i], n_prod); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], k2 ? 8 : 4); }
    fence_barrier_init();
  }
  if (warp_id == 1) { if (k2) { tmem_alloc_2sm<kTmemCols>(tmem_ptr_smem); } else { tmem_alloc<kTmemCols>(tmem_ptr_smem); } }
  tc_fence_before();
  if (k2) { cluster_sync_all(); } else { __syncthreads(); }
  tc_fence_after();
  pdl_wait();
  uint32_t const tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) { TAPS_STAMP(1); }

  if (warp_id == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues; in a pair: one warp per CTA) ==========
    // The loop state is warp-uniform on purpose: with `if (lane == 0)` around everything the compiler cannot prove that the operands of
    // UTMALDG / UTCHMMA (uniform-register instructions) are uniform and wraps every one of them in an ELECT / R2UR.BROADCAST retry loop --
    // ~100 SASS instructions per step on a single thread, more than the 256 cycles the MMAs of a step take.
    if (!(prm.debug & 1)) {
      // NOTE: no runtime integer divisions in these single-thread loops -- a dependent div/mod chain costs ~150 cycles on one thread and
      // six of them per step made the issue loop, not TMA or the tensor pipe, the bottleneck (measured: ~940 of 1560 cycles per step).
      int a_s = 0;            // stage the next halo tile goes to
      uint32_t a_par = 1;     // parity to wait for on its empty barrier (fresh barriers pass parity 1)
      auto issue_a = [&](int c) {  // halo tile of 64-channel block c: rows [m0, m0 + halo_rows) of the padded activation matrix
        int const s = a_s;
        mbar_wait(&a_empty[s], a_par);
        if (++a_s == a_stages) { a_s = 0; a_par ^= 1; }
        if (elect_one_sync()) {
          if (leader) { mbar_expect_tx(&a_full[s], (k2 ? 2u : 1u) * a_stage); } else { mbar_arrive_remote(&a_full[s], 0); }
          uint8_t *dst = a_ring + s * a_stage;
          for (int l = 0; l < prm.a_loads; ++l) {
            uint8_t *d = dst + l * prm.a_box_rows * 128;
            int const row = m0 + l * prm.a_box_rows;
            if (k2) {
              tma_load_2d_2sm(d, &a_hi_map, &a_full[s], c * IGEMM_BK, row);
              if (kPlanes == 2) { tma_load_2d_2sm(d + a_plane, &a_lo_map, &a_full[s], c * IGEMM_BK, row); }
            } else {
              tma_load_2d(d, &a_hi_map, &a_full[s], c * IGEMM_BK, row);
              if (kPlanes == 2) { tma_load_2d(d + a_plane, &a_lo_map, &a_full[s], c * IGEMM_BK, row); }
            }
          }
        }
        __syncwarp();
      };
      int const q_row0 = n0 + static_cast<int>(cta_rank) * BQ;
      // A tile of block c + a_stages - 1 is requested while block c runs, after enough filter tiles are in flight to keep the MMA busy
      int const t_pref = (a_stages == 1) ? 0 : min(b_stages, taps - 1);
      for (int c = 0; c < min(a_stages - 1, cblks); ++c) { issue_a(c); }
      int s = 0;
      uint32_t b_par = 1;
      for (int c = 0; c < cblks; ++c) {
        int wrow = c * prm.w_kb_rows + q_row0;  // packed filters: k-block (t * cblks + c) starts at row (t * cblks + c) * w_kb_rows
        for (int t = 0; t < taps; ++t, wrow += cblks * prm.w_kb_rows) {
          if (t == t_pref && c + a_stages - 1 < cblks) { issue_a(c + a_stages - 1); }
          mbar_wait(&b_empty[s], b_par);
          if (elect_one_sync()) {
            if (leader) { mbar_expect_tx(&b_full[s], (k2 ? 2u : 1u) * kBStage); } else { mbar_arrive_remote(&b_full[s], 0); }
            uint8_t *d = b_ring + s * kBStage;
            if (k2) {
              tma_load_2d_2sm(d, &w_hi_map, &b_full[s], 0, wrow);
              if (kPlanes == 2) { tma_load_2d_2sm(d + kBBytes, &w_lo_map, &b_full[s], 0, wrow); }
            } else {
              tma_load_2d(d, &w_hi_map, &b_full[s], 0, wrow);
              if (kPlanes == 2) { tma_load_2d(d + kBBytes, &w_lo_map, &b_full[s], 0, wrow); }
            }
          }
          __syncwarp();
          if (++s == b_stages) { s = 0; b_par ^= 1; }
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues; in a pair: the leader's warp) ============
    if (leader) {
      uint32_t const idesc = prm.idesc;
      uint32_t const tmem_x = tmem_base + 2 * tmem_buf_cols(BN);  // cross-term accumulator (hi*lo + lo*hi), drained once at the end
      uint32_t tmem_d = tmem_base;
      int t = 0, kx = 0, tap_row = 0, buf = 0;  // tap_row = ky*Wp + kx
      int sa = 0, sb = 0, in_chunk = 0, ci = 0;  // ring positions, steps into the current accumulation chunk, chunk index
      int cblk = 0;
      int const ksteps_last = (prm.debug & 2) ? 0 : prm.ksteps_last, ksteps_full = (prm.debug & 2) ? 0 : IGEMM_BK / IGEMM_UMMA_K;
      uint32_t a_par = 0, b_par = 0;
      bool first = true;
      for (int g = 0; g < nkb; ++g) {
        if (in_chunk == 0) {
          buf = ci & 1;
          mbar_wait(&tmem_empty_bar[buf], ((ci >> 1) & 1) ^ 1);
          tc_fence_after();
          tmem_d = tmem_base + buf * tmem_buf_cols(BN);
          first = true;
        }
        if (!(prm.debug & 1)) {
          if (t == 0) { mbar_wait(&a_full[sa], a_par); }
          mbar_wait(&b_full[sb], b_par);
        }
        tc_fence_after();
        if (g == 0 && lane == 0) { TAPS_STAMP(2); }
        uint32_t const a_addr = smem_u32(a_ring + sa * a_stage) + static_cast<uint32_t>(tap_row) * 128u;
        uint32_t const b_addr = smem_u32(b_ring + sb * kBStage);
        uint32_t const p_hi = sw128_desc_lo(a_addr), p_lo = sw128_desc_lo(a_addr + a_plane);
        uint32_t const q_hi = sw128_desc_lo(b_addr), q_lo = sw128_desc_lo(b_addr + kBBytes);
        bool const last_tap = (t == taps - 1);
        bool const chunk_end = (in_chunk + 1 == chunk || g == nkb - 1);
        int const nk = (cblk == cblks - 1) ? ksteps_last : ksteps_full;
        if (elect_one_sync()) {
        issue_kblock<kPlanes, k2>(tmem_d, tmem_x, p_hi, p_lo, q_hi, q_lo, idesc, first ? 0u : 1u, g == 0 ? 0u : 1u, nk);
        bool const ring_commits = !(prm.debug & 8);  // experiments (with bit 0): no stage-release commits
        if (k2) { if (ring_commits) { umma_commit_2sm(&b_empty[sb], 0x3); if (last_tap) { umma_commit_2sm(&a_empty[sa], 0x3); } } if (chunk_end) { umma_commit_2sm(&tmem_full_bar[buf], 0x3); } }
        else { if (ring_commits) { umma_commit(&b_empty[sb]); if (last_tap) { umma_commit(&a_empty[sa]); } } if (chunk_end) { umma_commit(&tmem_full_bar[buf]); } }
        }
        __syncwarp();
        first = false;
        if (chunk_end) { in_chunk = 0; ++ci; } else { ++in_chunk; }
        if (++sb == b_stages) { sb = 0; b_par ^= 1; }
        if (last_tap) { t = 0; kx = 0; tap_row = 0; ++cblk; if (++sa == a_stages) { sa = 0; a_par ^= 1; } }
        else { ++t; if (++kx == prm.kw) { kx = 0; tap_row += prm.Wp - (prm.kw - 1); } else { ++tap_row; } }
      }
      if (lane == 0) { TAPS_STAMP(3); }
    }
  } else {
    // ===================== epilogue warps (each CTA: its own 128 virtual pixels) =====================
    int const q = warp_id & 3;
    int const row = q * 32 + lane;
    for (int j = row; j < BN; j += 128) { bias_s[j] = (prm.has_bias && (n0 + j) < prm.q_rows) ? __ldg(prm.bias + n0 + j) : 0.0f; }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    float acc[BN];
#pragma unroll
    for (int j = 0; j < BN; ++j) { acc[j] = 0.0f; }
    for (int c = 0; c < nchunks; ++c) {
      int const buf = c & 1;
      mbar_wait(&tmem_full_bar[buf], (c >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 64 && c == 0) { TAPS_STAMP(4); }
      if (threadIdx.x == 64 && c == nchunks - 1) { TAPS_STAMP(5); }
      uint32_t const taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * tmem_buf_cols(BN);
#pragma unroll
      for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
      }
      if (kPlanes == 2 && c == nchunks - 1) {
        uint32_t const xaddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * tmem_buf_cols(BN);
#pragma unroll
        for (int j0 = 0; j0 < BN; j0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(xaddr + j0, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (leader) { mbar_arrive(&tmem_empty_bar[buf]); } else { mbar_arrive_remote(&tmem_empty_bar[buf], 0); } }
    }
    if (threadIdx.x == 64) { TAPS_STAMP(6); }
    // ---- write out: NCHW fp32; virtual pixels in the padding columns / rows are dropped ----
    float const inv = prm.p_scale[1] * prm.q_scale[1];
    float const floor_v = prm.relu ? 0.0f : -INFINITY;
    int const v = m0 + row;
    int const img = v / prm.HpWp, rem = v - img * prm.HpWp;
    int const yp = rem / prm.Wp, xp = rem - yp * prm.Wp;
    float amax = 0.0f;
    if (v < prm.m_rows && yp < prm.OH && xp < prm.OW) {
      float *o = prm.out + (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + yp * prm.OW + xp;
      amax = igemm_store_row<BN>(acc, inv, bias_s, floor_v, o, prm.out_hw, prm.q_rows - n0);
    }
    if (prm.out_absmax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
      if (lane == 0 && amax > 0.0f) { atomicMax(prm.out_absmax, __float_as_uint(amax)); }
    }
  }

  if (threadIdx.x == 64) { TAPS_STAMP(7); }
  tc_fence_before();
  if (k2) { cluster_sync_all(); } else { __syncthreads(); }
  if (threadIdx.x == 0) { TAPS_STAMP(8); }
  if (warp_id == 1) { if (k2) { tmem_dealloc_2sm<kTmemCols>(tmem_base); } else { tmem_dealloc<kTmemCols>(tmem_base); } }
}

}  // namespace b200
