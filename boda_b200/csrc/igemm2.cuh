// igemm2.cuh -- CTA-pair (tcgen05 cta_group::2) version of the contraction kernel in igemm.cuh.
//
// Why: with a 128 x 128 tile per SM the shared-memory port is the bottleneck, not the tensor pipe: every 128x128x16 fp16 MMA reads
// 4 KB of A + 4 KB of B from shared memory in 64 cycles (= the 128 B/clk port limit) while TMA is writing the next stage into the same
// memory, and the three passes of the fp32-parity mode re-read the hi planes. A CTA pair computes a 256 x BN tile with ONE
// tcgen05.mma.cta_group::2: each SM supplies its own 128 rows of P and only HALF of the Q tile (BN/2 rows), so per SM the
// shared-memory reads per MMA drop from 8 KB to 6 KB, the TMA fill per k-block from 64 KB to 48 KB (fp32-parity mode, BN = 128), and
// the L2->SM operand traffic for Q halves.
//
// Roles per CTA (192 threads): warp 0 = TMA producer (both CTAs; all transaction bytes are counted on the LEADER's full barrier),
// warp 1 = TMEM owner; in the leader CTA also the single-thread MMA issuer for the pair, warps 2..5 = epilogue (each CTA drains its own
// 128 TMEM lanes = its own 128 output rows). Accumulator ping-pong, cross-term accumulator and periodic draining are as in igemm.cuh.
#pragma once
#include "igemm.cuh"

namespace b200 {

template <int BN, int kPlanes>
struct Igemm2Cfg {
  static constexpr int kQRows = BN / 2;  // Q rows held by each CTA of the pair
  // A pipeline stage always has two P slots and two Q slots: hi + lo planes of ONE k-block in fp32-parity mode, or TWO consecutive k-blocks
  // in the single-plane 16-bit modes -- there a k-block is only 4 MMAs (256 cycles at BN = 128), less than the barrier wait + commit +
  // loop overhead of a step, so two k-blocks share one wait and one commit (8 MMAs back to back).
  static constexpr int kKbPerStage = (kPlanes == 1) ? 2 : 1;
  static constexpr int kKbBytes = kPlanes * (IGEMM_BM * 128 + kQRows * 128);  // bytes per k-block per CTA
  static constexpr int kStageBytes = 2 * (IGEMM_BM * 128 + kQRows * 128);
  static constexpr int kMaxSmem = 220 * 1024;
  static constexpr int kStagesRaw = (kMaxSmem - 3072) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kBarBytes = 2048;  // barriers in the first 512 B, two staged bias vectors (tile parity) at +1024
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kBarBytes;
  // TMEM: two ping-pong main accumulators (alternating per accumulation chunk) + in fp32-parity mode two cross-term accumulators
  // (alternating per TILE, so the MMA warp can start the next tile while the epilogue still reads this tile's cross terms)
  static constexpr uint32_t kBufCols = tmem_buf_cols(BN);
  static constexpr uint32_t kColsNeeded = (kPlanes == 2 ? 4 : 2) * kBufCols;
  static constexpr uint32_t kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128 : kColsNeeded <= 256 ? 256 : 512;
  static_assert(kColsNeeded <= 512, "TMEM has 512 columns");
};

// PERSISTENT: the grid is min(#tiles, #SM pairs) clusters; cluster c walks tiles c, c + #clusters, ... (N tiles fastest, so clusters running
// side by side share activation tiles in L2). Barrier setup, TMEM allocation and tensor-map prefetch are paid once per CTA instead of once
// per tile, and the epilogue of tile i (TMEM drain + global stores, ~3k cycles) overlaps the TMA / MMA main loop of tile i+1: the smem
// ring, the accumulator ping-pong and their mbarrier phases simply keep counting across tiles.
template <int BN, int kPlanes>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_umma_2cta_kernel(const __grid_constant__ CUtensorMap p_hi_map, const __grid_constant__ CUtensorMap p_lo_map,
                       const __grid_constant__ CUtensorMap q_hi_map, const __grid_constant__ CUtensorMap q_lo_map,
                       const IgemmParams prm) {
  using Cfg = Igemm2Cfg<BN, kPlanes>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kPBytes = IGEMM_BM * 128, kQBytes = Cfg::kQRows * 128;
  constexpr uint32_t kBufCols = Cfg::kBufCols;
  constexpr int kKb = Cfg::kKbPerStage;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *bar_mem = smem + kStages * Cfg::kStageBytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(bar_mem);   // used in the leader CTA only
  uint64_t *empty_bar = full_bar + kStages;                      // one per CTA, released by the leader's MMA commits
  uint64_t *tmem_full_bar = empty_bar + kStages;                 // [2], one per CTA
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;                  // [2], leader's copy collects both CTAs' epilogue warps
  uint64_t *x_empty_bar = tmem_empty_bar + 2;                    // [2], cross-term accumulators (fp32-parity mode), same counting
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(x_empty_bar + 2);
  float *bias_s = reinterpret_cast<float *>(bar_mem + 1024);     // [2][BN]

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  uint32_t const cta_rank = blockIdx.x & 1u;  // = %cluster_ctarank for (2,1,1) clusters; written this way the compiler can prove it warp-uniform and keeps the TMA operand
                                              // arithmetic of the producer in uniform registers (each R2UR on the single issuing thread's path costs issue latency)  // 0 = leader, 1 = peer (cluster dims are (2,1,1))
  bool const leader = (cta_rank == 0);
  int const n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  int const q_tiles = prm.q_tiles, n_tiles = prm.m_pair_tiles * q_tiles;
  int const nkb = prm.kblks_total;
  int const chunk = (prm.chunk_kblks + kKb - 1) / kKb * kKb;  // accumulation chunks end on stage boundaries
  int const nchunks = (nkb + chunk - 1) / chunk;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&p_hi_map);
    tma_prefetch_desc(&q_hi_map);
    if (kPlanes == 2) { tma_prefetch_desc(&p_lo_map); tma_prefetch_desc(&q_lo_map); }
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 2); mbar_init(&empty_bar[i], 1); }  // full: one arrive per CTA's producer
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 8); mbar_init(&x_empty_bar[i], 8); }  // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp_id == 1) { tmem_alloc_2sm<Cfg::kTmemCols>(tmem_ptr_smem); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  pdl_wait();
  uint32_t const tmem_base = *tmem_ptr_smem;

  if (warp_id == 0) {
    // ===================== TMA producer (both CTAs; whole warp walks the loop, one elected lane issues -- see igemm.cuh) ==========
    int it = 0;  // k-blocks issued so far, over all tiles: ring position and phase
    long long const t_begin = prm.ts ? clock64() : 0;
    long long w_empty = 0, t_issue = 0;
    for (int tile = cluster_id; tile < ((prm.debug & 1) ? 0 : n_tiles); tile += n_clusters) {  // (debug bit 0: no loads at all)
      int const mt = tile / q_tiles, nt = tile - mt * q_tiles;
      int const m0 = (mt * 2 + static_cast<int>(cta_rank)) * IGEMM_BM;  // this CTA's own 128 P rows (the pair covers 256 consecutive ones)
      int img = 0, h_base = 0, w_base = 0;
      if (prm.p_im2col) {
        img = m0 / prm.ohw;
        int const rem = m0 - img * prm.ohw;
        int const oy = rem / prm.ow, ox = rem - oy * prm.ow;
        h_base = oy * prm.sy - prm.py;
        w_base = ox * prm.sx - prm.px;
      }
      int const q_row0 = nt * BN + static_cast<int>(cta_rank) * Cfg::kQRows;  // this CTA's half of the Q tile
      int cb = 0, kx = 0, ky = 0;  // (tap, channel block) of k-block i, advanced without integer divisions (single-thread loop)
      for (int i = 0; i < nkb; i += kKb, ++it) {
        int const s = it % kStages;
        uint32_t const ph = (it / kStages) & 1;
        int const nh = min(kKb, nkb - i);  // k-blocks in this stage
        if (prm.ts) { long long const t0 = clock64(); mbar_wait(&empty_bar[s], ph ^ 1); w_empty += clock64() - t0; } else { mbar_wait(&empty_bar[s], ph ^ 1); }
        long long const t_i0 = prm.ts ? clock64() : 0;
        bool const issue = elect_one_sync();
        if (issue) {
          if (leader) { mbar_expect_tx(&full_bar[s], 2u * nh * Cfg::kKbBytes); }  // both CTAs' bytes land on the leader's barrier
          else { mbar_arrive_remote(&full_bar[s], 0); }
        }
        uint8_t *st = smem + s * Cfg::kStageBytes;
#pragma unroll
        for (int j = 0; j < kKb; ++j) {
          if (j < nh) {
            int const kb = i + j;
            if (issue) {
              uint8_t *p_hi = st + (kPlanes == 2 ? 0 : j) * kPBytes, *p_lo = st + kPBytes;              // slot j (16-bit modes) or hi / lo planes
              uint8_t *q_hi = st + 2 * kPBytes + (kPlanes == 2 ? 0 : j) * kQBytes, *q_lo = st + 2 * kPBytes + kQBytes;
              if (prm.p_im2col) {
                tma_load_im2col_4d_2sm(p_hi, &p_hi_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky);
                if (kPlanes == 2) { tma_load_im2col_4d_2sm(p_lo, &p_lo_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky); }
              } else {
                tma_load_2d_2sm(p_hi, &p_hi_map, &full_bar[s], kb * IGEMM_BK, m0);
                if (kPlanes == 2) { tma_load_2d_2sm(p_lo, &p_lo_map, &full_bar[s], kb * IGEMM_BK, m0); }
              }
              int const qc0 = prm.q_kb_rows ? 0 : kb * IGEMM_BK, qc1 = q_row0 + kb * prm.q_kb_rows;
              tma_load_2d_2sm(q_hi, &q_hi_map, &full_bar[s], qc0, qc1);
              if (kPlanes == 2) { tma_load_2d_2sm(q_lo, &q_lo_map, &full_bar[s], qc0, qc1); }
            }
            if (++cb == prm.cblks) { cb = 0; if (++kx == prm.kw) { kx = 0; ++ky; } }
          }
        }
        __syncwarp();
        if (prm.ts) { t_issue += clock64() - t_i0; }
      }
    }
    if (prm.ts && leader && lane == 0) { long long *t = prm.ts + cluster_id * 16; t[0] = clock64() - t_begin; t[1] = w_empty; t[13] = t_issue; }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (leader CTA only, one elected thread for the pair) =====================
    if (leader) {
      long long const t_begin = prm.ts ? clock64() : 0;
      long long w_full = 0, w_tmem = 0, t_first_full = 0;
      uint32_t const idesc = prm.idesc;  // M = 256 (the pair), N = BN
      int const kb_mod = prm.kb_mod, ksteps_last = prm.ksteps_last;
      int it = 0, gc = 0, ti = 0;  // k-blocks, accumulation chunks and tiles done so far
      for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++ti) {
        uint32_t const tmem_x = tmem_base + (2 + (ti & 1)) * kBufCols;
        if (kPlanes == 2) { mbar_wait(&x_empty_bar[ti & 1], ((ti >> 1) & 1) ^ 1); }  // the epilogue has read the cross terms of tile ti - 2
        int kb_in_grp = 0;
        int i = 0;
        for (int c = 0; c < nchunks; ++c, ++gc) {
          int const buf = gc & 1;
          if (prm.ts) { long long const t0 = clock64(); mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); w_tmem += clock64() - t0; } else { mbar_wait(&tmem_empty_bar[buf], ((gc >> 1) & 1) ^ 1); }
          tc_fence_after();
          uint32_t const tmem_d = tmem_base + buf * kBufCols;
          int const i_end = min(i + chunk, nkb);
          bool first = true;
          for (; i < i_end; i += kKb, ++it) {
            int const s = it % kStages;
            uint32_t const ph = (it / kStages) & 1;
            int const nh = min(kKb, nkb - i);
            if (prm.debug & 1) {}  // experiments: MMA on whatever the shared memory holds
            else if (prm.ts) { long long const t0 = clock64(); mbar_wait(&full_bar[s], ph); long long const t1 = clock64(); if (it == 0) { t_first_full = t1 - t_begin; } else { w_full += t1 - t0; } } else { mbar_wait(&full_bar[s], ph); }
            tc_fence_after();
            uint32_t const st = smem_u32(smem + s * Cfg::kStageBytes);
            int nk[kKb];
#pragma unroll
            for (int j = 0; j < kKb; ++j) { nk[j] = IGEMM_BK / IGEMM_UMMA_K; if (j < nh && ++kb_in_grp == kb_mod) { kb_in_grp = 0; nk[j] = ksteps_last; } if (prm.debug & 2) { nk[j] = 0; } }  // (debug bit 1: barrier traffic only)
            if (elect_one_sync()) {
              if (kPlanes == 2) {
                issue_kblock<2, true>(tmem_d, tmem_x, sw128_desc_lo(st), sw128_desc_lo(st + kPBytes), sw128_desc_lo(st + 2 * kPBytes),
                                      sw128_desc_lo(st + 2 * kPBytes + kQBytes), idesc, first ? 0u : 1u, i == 0 ? 0u : 1u, nk[0]);
              } else {
#pragma unroll
                for (int j = 0; j < kKb; ++j) {
                  if (j < nh) {
                    issue_kblock<1, true>(tmem_d, 0u, sw128_desc_lo(st + j * kPBytes), 0u, sw128_desc_lo(st + 2 * kPBytes + j * kQBytes), 0u, idesc,
                                          (first && j == 0) ? 0u : 1u, 1u, nk[j]);
                  }
                }
              }
              umma_commit_2sm(&empty_bar[s], 0x3);  // release the stage in both CTAs
              if (i + kKb >= i_end) { umma_commit_2sm(&tmem_full_bar[buf], 0x3); }  // wake both CTAs' epilogues
            }
            __syncwarp();
            first = false;
          }
        }
      }
      if (prm.ts && lane == 0) { long long *t = prm.ts + cluster_id * 16; t[2] = clock64() - t_begin; t[3] = w_full; t[4] = w_tmem; t[5] = t_first_full; t[6] = it; }
    }
  } else {
    // ===================== epilogue warps (each CTA: its own 128 rows) =====================
    int const q = warp_id & 3;
    int const row = q * 32 + lane;
    float const inv = prm.p_scale[1] * prm.q_scale[1], inv_recip = prm.p_scale[0] * prm.q_scale[0];
    float const floor_v = prm.relu ? 0.0f : -INFINITY;
    float amax = 0.0f;
    float s_out = 1.0f;
    if (prm.out16 && prm.w_l1max) {  // scale of the consumer's fp16 planes from the output bound (identical in every CTA); CTA 0 publishes it
      s_out = igemm_out_scale(prm, reinterpret_cast<float *>(bar_mem + 768), row);
      if (blockIdx.x == 0 && row == 0) { prm.out16_scale2[0] = s_out; prm.out16_scale2[1] = 1.0f / s_out; }
    }
    int gc = 0, ti = 0;
    long long const t_begin = prm.ts ? clock64() : 0;
    long long w_acc = 0, t_drain = 0, t_store = 0;
    for (int tile = cluster_id; tile < n_tiles; tile += n_clusters, ++ti) {
      int const mt = tile / q_tiles, nt = tile - mt * q_tiles;
      int const m0 = (mt * 2 + static_cast<int>(cta_rank)) * IGEMM_BM, n0 = nt * BN;
      // this tile's bias, staged in the buffer of its parity (the other one may still be read by a slower epilogue warp's stores)
      float *bias_t = bias_s + (ti & 1) * BN;
      for (int j = row; j < BN; j += 128) { bias_t[j] = (prm.has_bias && (n0 + j) < prm.q_rows) ? __ldg(prm.bias + n0 + j) : 0.0f; }
      asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
      float acc[BN];
      {
        float const *res_row = nullptr;
        if (prm.res && m0 + row < prm.p_rows) {
          int const img = (m0 + row) / prm.out_hw, pix = (m0 + row) - img * prm.out_hw;
          res_row = prm.res + (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + pix;
        }
        igemm_acc_init<BN>(acc, res_row, prm.out_hw, prm.q_rows - n0, inv_recip);
      }
      for (int c = 0; c < nchunks; ++c, ++gc) {
        int const buf = gc & 1;
        long long t0 = 0;
        if (prm.ts) { t0 = clock64(); mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); long long const t1 = clock64(); w_acc += t1 - t0; t0 = t1; } else { mbar_wait(&tmem_full_bar[buf], (gc >> 1) & 1); }
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
        if (prm.ts) { t_drain += clock64() - t0; }
      }
      long long const t_st0 = prm.ts ? clock64() : 0;
      if (kPlanes == 2) {  // the last chunk's commit also covered every cross-term MMA of this tile
        tc_fence_after();
        uint32_t const xaddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (2 + (ti & 1)) * kBufCols;
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
        if (lane == 0) { if (leader) { mbar_arrive(&x_empty_bar[ti & 1]); } else { mbar_arrive_remote(&x_empty_bar[ti & 1], 0); } }
      }
      int const prow = m0 + row;
      if (prow < prm.p_rows) {
        int const img = prow / prm.out_hw, pix = prow - img * prm.out_hw;
        float *o = prm.out + (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + pix;
        amax = fmaxf(amax, igemm_store_row<BN>(acc, inv, bias_t, floor_v, o, prm.out_hw, prm.q_rows - n0));
        if (prm.out16) {
          long long const o16 = static_cast<long long>(prow) * prm.out16_pitch + n0;
          if (prm.w_l1max) { igemm_store_row_split16<BN>(acc, inv, bias_t, floor_v, s_out, prm.out16 + o16, prm.out16_lo ? prm.out16_lo + o16 : nullptr, prm.q_rows - n0); }
          else { igemm_store_row_bf16<BN>(acc, inv, bias_t, floor_v, prm.out16 + o16, prm.q_rows - n0); }
        }
      }
      if (prm.ts) { t_store += clock64() - t_st0; }
    }
    if (prm.ts && leader && warp_id == 2 && lane == 0) { long long *t = prm.ts + cluster_id * 16; t[8] = clock64() - t_begin; t[9] = w_acc; t[10] = t_drain; t[11] = t_store; t[12] = ti; }
    if (prm.out_absmax) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
      if (lane == 0 && amax > 0.0f) { atomicMax(prm.out_absmax, __float_as_uint(amax)); }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still be arriving on the leader's barriers / the MMA may still read the peer's shared memory
  if (warp_id == 1) { tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base); }
}

}  // namespace b200
