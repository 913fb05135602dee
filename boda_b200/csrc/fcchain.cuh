// fcchain.cuh -- a chain of inner-product-shaped convolutions (AlexNet-ng fc6-conv -> fc7-conv -> fc8-conv; src/rtc_fwd.cc:486-520 issues
// them as three conv calls) as ONE persistent kernel.
//
// At batch <= 32 these layers are weight streams: 9216x4096 + 4096x4096 + 4096x1000 weights = 235 MB in the fp32-parity planes against
// 0.5 MB of activations, 36 us at the HBM rate. As separate launches (contraction, split-K reduce, activation pack per layer: 8 kernels)
// the stream stops eight times -- launch, prologue, pipeline fill, drain -- and the chain took ~100 us of the 374 us step (r02 launch list:
// fc6 3.7 TB/s, fc7 2.9 TB/s, fc8 1.1 TB/s while they run). Here every CTA walks the layers in turn and the weight stream never waits for
// the activations: the producer warp issues the next layer's weight tiles (which depend on nothing) into the free stages of its ring BEFORE
// it waits for the grid-wide barrier that says the next layer's activation planes are complete.
//
// Per layer l (the arithmetic of igemm_umma_kernel in swapped mode + splitk_reduce_kernel + pack_rows_split_kernel, in their order, so
// every node is BIT-IDENTICAL to the unchained path):
//   unit (tile, split) = CTA u: D[128 out chans x 32 images] over its k-blocks, TMEM drained every chunk_kblks k-blocks (ping-pong) with
//     round-to-nearest adds in the epilogue warps, + the cross-term accumulator; partial = acc * inv_scale -> ws[split][img][chan]
//   barrier A (all units of l written) -> every CTA reduces a slice of the outputs over the splits in split order (+ bias, ReLU), writes
//     the fp32 node out[img][chan] and folds max|out| into the node's abs-max cell
//   barrier B (all slices reduced) -> scale s from the abs-max cell; every CTA writes its slice of the next layer's fp16 hi / lo (or bf16)
//     planes [32][K]
//   barrier C (planes complete): the producer warps wait for it before the first activation tile of layer l+1.
// Barriers are counters in global memory; every CTA of the grid is resident (grid <= #SMs, one CTA per SM by shared memory), spins are
// bounded. The last CTA to leave re-arms the counters.
//
// Warp roles as igemm.cuh: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue / reduce / pack.
#pragma once
#include "igemm.cuh"

namespace b200 {

constexpr int FC_MAX_LAYERS = 4;
constexpr int FC_BN = 32;          // images per tile (UMMA N)
constexpr int FC_SYNC_WORDS = 32;  // [3 * l + {0, 1, 2}] barriers A, B, C of layer l; [15] exit counter; [16 + l] stand-in abs-max cells

struct FcChainMaps {
  CUtensorMap w_hi[FC_MAX_LAYERS], w_lo[FC_MAX_LAYERS];  // packed filters [k-block][out chan (padded)][64], box 64 x 128
  CUtensorMap a_hi[FC_MAX_LAYERS], a_lo[FC_MAX_LAYERS];  // activation planes [image][K], box 64 x 32
};

struct FcLayer {
  int n_out;         // out chans
  int oc_pad;        // rows per k-block of the packed filters
  int tiles, splits, kblks_total, kblks_per_split;
  int kb_mod, ksteps_last;
  int relu, has_bias;
  float *out;                 // fp32 node [batch][n_out]
  float const *bias;
  float const *w_scale2;      // {scale, 1 / scale} of the filter planes
  unsigned int *out_absmax;   // abs-max cell of `out` (zero at kernel start)
  uint16_t *nxt_hi, *nxt_lo;  // the planes of `out` the next layer reads, [FC_BN][n_out]; null for the last layer
  float *nxt_scale2;
};

// Multi-GPU: the last layer's logits go straight into every rank's gather buffer over NVLink (b200_shard.h: b200_gather_desc_t), from the CTAs
// that reduce them; the last CTA to leave publishes the step to every peer and waits (bounded) until the previous step of every rank has
// landed locally -- the compute and the collective are one kernel, and a CUDA graph can replay it (the step lives in device memory).
struct FcGather {
  unsigned char *peer_base[8];
  unsigned char *local_base;
  unsigned long long bytes_per_rank, flag_bytes;
  int rank, world;  // world == 0: off
};

struct FcChainParams {
  int n_layers, batch, chunk_kblks, bf16;
  uint32_t idesc;
  float const *a0_scale2;  // {scale, 1 / scale} of layer 0's activation planes (written by their producer)
  float *ws;               // split-K partials, max over layers of splits * batch * n_out floats
  unsigned int *sync;      // FC_SYNC_WORDS counters, zero at kernel start (re-armed by the last CTA)
  int l2_ahead, l2_next;   // L2 prefetch of filter tiles: layer 0, this many k-blocks ahead of the ring (0 = off); next layer's tiles at the end of a layer
  FcGather g;
  long long *ts;           // experiments (debug_flags bit 4): [CTA][32] globaltimer stamps (ns), see the FC_TS_* slots; null = off
  FcLayer L[FC_MAX_LAYERS];
};

// stamp slots per layer l (8 * l + ...): producer: 0 = first filter tiles issued, 1 = barrier C / pdl_wait passed; epilogue: 2 = accumulators drained,
// 3 = barrier A passed, 4 = slice reduced, 5 = barrier B passed, 6 = planes written. Slot 31 = kernel start.
__device__ __forceinline__ long long fc_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define FC_STAMP(slot) do { if (prm.ts) { prm.ts[static_cast<long long>(blockIdx.x) * 32 + (slot)] = fc_now(); } } while (0)

// one thread, after a CTA-level barrier: arrive on a grid-wide counter / wait until `target` CTAs have. The fence is cumulative: it also orders
// the stores of the CTA's other threads, which the CTA barrier made visible to this one, before the arrival.
__device__ __forceinline__ void fc_sync_arrive(unsigned int *ctr) {
  __threadfence();
  atomicAdd(ctr, 1u);
}
__device__ __forceinline__ void fc_sync_wait(unsigned int *ctr, unsigned int target, int what) {
  unsigned int v = 0, spins = 0;
  while (true) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) { break; }
    if (++spins > (1u << 22)) { printf("b200: fc chain barrier %d never completed (block %d: %u of %u)\n", what, blockIdx.x, v, target); __trap(); }
  }
}

template <int kPlanes>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
fc_chain_kernel(const __grid_constant__ FcChainMaps maps, const __grid_constant__ FcChainParams prm) {
  using Cfg = IgemmCfg<FC_BN, kPlanes>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kPBytes = IGEMM_BM * 128, kQBytes = FC_BN * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *bar_mem = smem + kStages * Cfg::kStageBytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(bar_mem);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
  float *bcast = reinterpret_cast<float *>(bar_mem + 512);  // epilogue warps: the next layer's activation scale

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  int const u = blockIdx.x;
  if (threadIdx.x == 0) { FC_STAMP(31); }
  if (warp_id == 0 && lane == 0) {
    for (int l = 0; l < prm.n_layers; ++l) {
      tma_prefetch_desc(&maps.w_hi[l]); tma_prefetch_desc(&maps.a_hi[l]);
      if (kPlanes == 2) { tma_prefetch_desc(&maps.w_lo[l]); tma_prefetch_desc(&maps.a_lo[l]); }
    }
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp_id == 1) { tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t const tmem_base = *tmem_ptr_smem;

  // this CTA's unit of layer l: k-blocks [kb_begin, kb_begin + nkb) of out-chan tile `tile` (nkb = 0: no unit)
  auto unit_of = [&](int l, int &tile, int &split, int &kb_begin) {
    FcLayer const &L = prm.L[l];
    if (u >= L.tiles * L.splits) { tile = 0; split = 0; kb_begin = 0; return 0; }
    tile = u % L.tiles;
    split = u / L.tiles;
    kb_begin = split * L.kblks_per_split;
    return max(0, min(kb_begin + L.kblks_per_split, L.kblks_total) - kb_begin);
  };

  if (warp_id == 0) {
    // ===================== TMA producer =====================
    int ig = 0;  // stages issued so far (ring position and parity run on across the layers)
    for (int l = 0; l < prm.n_layers; ++l) {
      int tile, split, kb_begin;
      int const nkb = unit_of(l, tile, split, kb_begin);
      FcLayer const &L = prm.L[l];
      int const m0 = tile * IGEMM_BM;
      // the first stages of a layer: filter tiles go out at once, the activation tiles after the planes are known to be complete
      int const pre = min(nkb, kStages);
      for (int i = 0; i < pre; ++i) {
        int const s = (ig + i) % kStages;
        mbar_wait(&empty_bar[s], (((ig + i) / kStages) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          uint8_t *st = smem + s * Cfg::kStageBytes;
          tma_load_2d(st, &maps.w_hi[l], &full_bar[s], 0, m0 + (kb_begin + i) * L.oc_pad);
          if (kPlanes == 2) { tma_load_2d(st + kPBytes, &maps.w_lo[l], &full_bar[s], 0, m0 + (kb_begin + i) * L.oc_pad); }
        }
        __syncwarp();
      }
      if (lane == 0) { FC_STAMP(8 * l + 0); }
      if (l == 0) { pdl_wait(); }  // layer 0's planes come from the previous kernel
      else {
        if (lane == 0) { fc_sync_wait(prm.sync + 3 * (l - 1) + 2, gridDim.x, 3 * (l - 1) + 2); }
        __syncwarp();
        asm volatile("fence.proxy.async.global;" ::: "memory");  // the planes were written with ordinary stores, the loads below go through the async proxy
      }
      if (lane == 0) { FC_STAMP(8 * l + 1); }
      for (int i = 0; i < pre; ++i) {
        int const s = (ig + i) % kStages;
        if (elect_one_sync()) {
          uint8_t *q = smem + s * Cfg::kStageBytes + kPlanes * kPBytes;
          tma_load_2d(q, &maps.a_hi[l], &full_bar[s], (kb_begin + i) * IGEMM_BK, 0);
          if (kPlanes == 2) { tma_load_2d(q + kQBytes, &maps.a_lo[l], &full_bar[s], (kb_begin + i) * IGEMM_BK, 0); }
        }
        __syncwarp();
      }
      // the rest of this layer's filter tiles start their way from HBM now, into L2 (layer 0; later layers were requested during the previous
      // layer's reduction, below): the ring holds 5 stages = 200 KB per SM, too little in flight to keep HBM busy (4.9 TB/s, r02 stamps)
      if (l == 0 && prm.l2_ahead) {
        for (int i = pre + lane; i < min(nkb, pre + prm.l2_ahead); i += 32) {
          tma_prefetch_l2_2d(&maps.w_hi[l], 0, m0 + (kb_begin + i) * L.oc_pad);
          if (kPlanes == 2) { tma_prefetch_l2_2d(&maps.w_lo[l], 0, m0 + (kb_begin + i) * L.oc_pad); }
        }
      }
      for (int i = pre; i < nkb; ++i) {
        int const s = (ig + i) % kStages;
        mbar_wait(&empty_bar[s], (((ig + i) / kStages) & 1) ^ 1);
        if (l == 0 && prm.l2_ahead && i + prm.l2_ahead < nkb && elect_one_sync()) {
          tma_prefetch_l2_2d(&maps.w_hi[l], 0, m0 + (kb_begin + i + prm.l2_ahead) * L.oc_pad);
          if (kPlanes == 2) { tma_prefetch_l2_2d(&maps.w_lo[l], 0, m0 + (kb_begin + i + prm.l2_ahead) * L.oc_pad); }
        }
        if (elect_one_sync()) {
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          uint8_t *st = smem + s * Cfg::kStageBytes;
          int const kb = kb_begin + i;
          tma_load_2d(st, &maps.w_hi[l], &full_bar[s], 0, m0 + kb * L.oc_pad);
          if (kPlanes == 2) { tma_load_2d(st + kPBytes, &maps.w_lo[l], &full_bar[s], 0, m0 + kb * L.oc_pad); }
          uint8_t *q = st + kPlanes * kPBytes;
          tma_load_2d(q, &maps.a_hi[l], &full_bar[s], kb * IGEMM_BK, 0);
          if (kPlanes == 2) { tma_load_2d(q + kQBytes, &maps.a_lo[l], &full_bar[s], kb * IGEMM_BK, 0); }
        }
        __syncwarp();
      }
      ig += nkb;
      // this layer's loads are all issued: while its accumulators drain and the grid reduces, HBM would idle -- request the next layer's
      // filter tiles (beyond the stages the ring will take directly) into L2
      if (prm.l2_next && l + 1 < prm.n_layers) {
        int t2, s2, kb2;
        int const nkb2 = unit_of(l + 1, t2, s2, kb2);
        FcLayer const &N = prm.L[l + 1];
        for (int i = min(nkb2, kStages) + lane; i < nkb2; i += 32) {
          tma_prefetch_l2_2d(&maps.w_hi[l + 1], 0, t2 * IGEMM_BM + (kb2 + i) * N.oc_pad);
          if (kPlanes == 2) { tma_prefetch_l2_2d(&maps.w_lo[l + 1], 0, t2 * IGEMM_BM + (kb2 + i) * N.oc_pad); }
        }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer =====================
    // (the cross-term accumulator of layer l is overwritten by layer l+1's first MMA: that one waits for activation tiles which exist only
    // after barrier C of layer l, i.e. after this CTA's epilogue warps have drained it)
    int ig = 0, cg = 0;  // stages / accumulator chunks consumed so far
    uint32_t const idesc = prm.idesc;
    int const chunk = prm.chunk_kblks;
    for (int l = 0; l < prm.n_layers; ++l) {
      int tile, split, kb_begin;
      int const nkb = unit_of(l, tile, split, kb_begin);
      FcLayer const &L = prm.L[l];
      int kb_in_grp = kb_begin % L.kb_mod;
      int const nchunks = (nkb + chunk - 1) / chunk;
      int i = 0;
      for (int c = 0; c < nchunks; ++c) {
        int const buf = (cg + c) & 1;
        mbar_wait(&tmem_empty_bar[buf], (((cg + c) >> 1) & 1) ^ 1);
        tc_fence_after();
        uint32_t const tmem_d = tmem_base + buf * tmem_buf_cols(FC_BN);
        uint32_t const tmem_x = tmem_base + 2 * tmem_buf_cols(FC_BN);
        int const i_end = min(i + chunk, nkb);
        bool first = true;
        for (; i < i_end; ++i) {
          int const s = (ig + i) % kStages;
          mbar_wait(&full_bar[s], ((ig + i) / kStages) & 1);
          tc_fence_after();
          uint32_t const st = smem_u32(smem + s * Cfg::kStageBytes);
          uint32_t const p_hi = sw128_desc_lo(st), p_lo = sw128_desc_lo(st + kPBytes);
          uint32_t const q_hi = sw128_desc_lo(st + kPlanes * kPBytes), q_lo = sw128_desc_lo(st + kPlanes * kPBytes + kQBytes);
          int nk = IGEMM_BK / IGEMM_UMMA_K;
          if (++kb_in_grp == L.kb_mod) { kb_in_grp = 0; nk = L.ksteps_last; }
          if (elect_one_sync()) {
            issue_kblock<kPlanes, false>(tmem_d, tmem_x, p_hi, p_lo, q_hi, q_lo, idesc, first ? 0u : 1u, i == 0 ? 0u : 1u, nk);
            umma_commit(&empty_bar[s]);
            if (i == i_end - 1) { umma_commit(&tmem_full_bar[buf]); }
          }
          __syncwarp();
          first = false;
        }
      }
      ig += nkb;
      cg += nchunks;
    }
  } else {
    // ===================== epilogue warps: drain, partials, reduce, planes =====================
    int const q = warp_id & 3;
    int const row = q * 32 + lane;
    int cg = 0;
    float a_inv = 0.0f;  // 1 / scale of the current layer's activation planes
    if (prm.g.world > 0 && u == 0 && row < prm.g.world) {
      // one step late: the PREVIOUS step's logits of every rank have landed in the local gather buffer before this forward ends -- polled here,
      // while these warps would only wait for the first accumulators (thread `row` watches rank `row`; bounded)
      unsigned int const prev = *reinterpret_cast<volatile unsigned int const *>(prm.g.local_base + 32 * 4);  // = this launch's step - 1
      unsigned int const *flag = reinterpret_cast<unsigned int const *>(prm.g.local_base) + row;
      unsigned int v = 0, spins = 0;
      while (prev > 0u) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (static_cast<int>(v - prev) >= 0) { break; }
        __nanosleep(100);
        if (++spins > (1u << 25)) { printf("b200: fc chain gather: rank %d never published step %u (have %u)\n", row, prev, v); __trap(); }
      }
    }
    int const gtid = u * 128 + row, gstride = static_cast<int>(gridDim.x) * 128;
    for (int l = 0; l < prm.n_layers; ++l) {
      int tile, split, kb_begin;
      int const nkb = unit_of(l, tile, split, kb_begin);
      FcLayer const &L = prm.L[l];
      int const nchunks = (nkb + prm.chunk_kblks - 1) / prm.chunk_kblks;
      int const total = prm.batch * L.n_out;
      float acc[FC_BN];
#pragma unroll
      for (int j = 0; j < FC_BN; ++j) { acc[j] = 0.0f; }
      for (int c = 0; c < nchunks; ++c) {
        int const buf = (cg + c) & 1;
        mbar_wait(&tmem_full_bar[buf], ((cg + c) >> 1) & 1);
        tc_fence_after();
        uint32_t const taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * tmem_buf_cols(FC_BN);
        {
          uint32_t r[32];
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j] += __uint_as_float(r[j]); }
        }
        if (kPlanes == 2 && c == nchunks - 1) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * tmem_buf_cols(FC_BN), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) { acc[j] += __uint_as_float(r[j]); }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&tmem_empty_bar[buf]); }
      }
      cg += nchunks;
      if (row == 0) { FC_STAMP(8 * l + 2); }
      if (nkb > 0) {
        // (the main loop above ran on tiles that exist only after pdl_wait / barrier C in the producer warp: the scale read below is ordered behind it)
        if (l == 0) { a_inv = __ldcg(prm.a0_scale2 + 1); }
        float const inv = L.w_scale2[1] * a_inv;
        int const ch = tile * IGEMM_BM + row;
        if (ch < L.n_out) {
          float *w = prm.ws + static_cast<long long>(split) * total + ch;
#pragma unroll
          for (int j = 0; j < FC_BN; ++j) { if (j < prm.batch) { w[static_cast<long long>(j) * L.n_out] = fmaf(acc[j], inv, 0.0f); } }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (row == 0) {
        if (nkb > 0) { fc_sync_arrive(prm.sync + 3 * l); }
        fc_sync_wait(prm.sync + 3 * l, static_cast<unsigned int>(L.tiles * L.splits), 3 * l);
        FC_STAMP(8 * l + 3);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- reduce this CTA's slice over the splits, in split order (splitk_reduce_kernel's arithmetic) ----
      float amax = 0.0f;
      float const floor_v = L.relu ? 0.0f : -INFINITY;
      bool const gather_here = prm.g.world > 0 && l + 1 == prm.n_layers;
      unsigned long long g_off = 0;
      if (gather_here) {  // slot [step parity][this rank] of a gather buffer; step = the device-side counter + 1 (same value in every CTA: only the last CTA to leave advances it)
        unsigned int const step = *reinterpret_cast<volatile unsigned int const *>(prm.g.local_base + 32 * 4) + 1u;
        g_off = prm.g.flag_bytes + (static_cast<unsigned long long>(step & 1u) * prm.g.world + prm.g.rank) * prm.g.bytes_per_rank;
      }
      // (four consecutive outputs per thread, the loads of all splits in flight together: with one output per iteration the 8 iterations of a
      // thread were 8 dependent L2 round trips, 6.5 us per layer, r02 stamps. Host: n_out % 4 == 0.)
      for (int i4 = gtid; i4 * 4 < total; i4 += gstride) {
        int const idx = i4 * 4;
        float4 p[16];
#pragma unroll
        for (int s = 0; s < 16; ++s) { if (s < L.splits) { p[s] = __ldcg(reinterpret_cast<float4 const *>(prm.ws + static_cast<long long>(s) * total + idx)); } }
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int s = 0; s < 16; ++s) { if (s < L.splits) { v.x += p[s].x; v.y += p[s].y; v.z += p[s].z; v.w += p[s].w; } }
        if (L.has_bias) {
          float4 const b4 = __ldg(reinterpret_cast<float4 const *>(L.bias + (idx % L.n_out)));
          v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
        }
        v.x = fmaxf(v.x, floor_v); v.y = fmaxf(v.y, floor_v); v.z = fmaxf(v.z, floor_v); v.w = fmaxf(v.w, floor_v);
        *reinterpret_cast<float4 *>(L.out + idx) = v;
        if (gather_here) {
#pragma unroll 1
          for (int pr = 0; pr < prm.g.world; ++pr) { *reinterpret_cast<float4 *>(prm.g.peer_base[pr] + g_off + static_cast<unsigned long long>(idx) * 4ull) = v; }
        }
        amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
      if (gather_here) {  // the CTA's stores into the peers' buffers are ordered (system scope, cumulative over the CTA barrier) before it counts itself out below
        // (one fence per CTA: a system-scope fence in every one of the 16k epilogue threads cost ~10 us per step, r02)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (row == 0) { __threadfence_system(); }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
      if (lane == 0 && amax > 0.0f) { atomicMax(L.out_absmax, __float_as_uint(amax)); }
      if (row == 0) { FC_STAMP(8 * l + 4); }
      if (l + 1 == prm.n_layers) { break; }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (row == 0) {
        fc_sync_arrive(prm.sync + 3 * l + 1);
        fc_sync_wait(prm.sync + 3 * l + 1, gridDim.x, 3 * l + 1);
        FC_STAMP(8 * l + 5);
        float const s = prm.bf16 ? 1.0f : scale_from_absmax_bits(*reinterpret_cast<volatile unsigned int *>(L.out_absmax));
        bcast[0] = s;
        if (u == 0) { L.nxt_scale2[0] = s; L.nxt_scale2[1] = 1.0f / s; }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // ---- the next layer's planes (pack_rows_split_kernel's arithmetic) ----
      float const s = bcast[0];
      a_inv = 1.0f / s;
      for (int i4 = gtid; i4 * 4 < total; i4 += gstride) {
        int const idx = i4 * 4;
        float4 const o4 = __ldcg(reinterpret_cast<float4 const *>(L.out + idx));
        float const v[4] = {o4.x * s, o4.y * s, o4.z * s, o4.w * s};
        uint16_t h[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (prm.bf16) {
            __nv_bfloat16 const hh = __float2bfloat16_rn(v[j]);
            __nv_bfloat16 const ll = __float2bfloat16_rn(v[j] - __bfloat162float(hh));
            h[j] = __bfloat16_as_ushort(hh); lo[j] = __bfloat16_as_ushort(ll);
          } else {
            __half const hh = __float2half_rn(v[j]);
            __half const ll = __float2half_rn(v[j] - __half2float(hh));
            h[j] = __half_as_ushort(hh); lo[j] = __half_as_ushort(ll);
          }
        }
        *reinterpret_cast<uint2 *>(L.nxt_hi + idx) = make_uint2(h[0] | (static_cast<uint32_t>(h[1]) << 16), h[2] | (static_cast<uint32_t>(h[3]) << 16));
        if (kPlanes == 2) { *reinterpret_cast<uint2 *>(L.nxt_lo + idx) = make_uint2(lo[0] | (static_cast<uint32_t>(lo[1]) << 16), lo[2] | (static_cast<uint32_t>(lo[3]) << 16)); }
      }
      asm volatile("fence.proxy.async.global;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (row == 0) { fc_sync_arrive(prm.sync + 3 * l + 2); FC_STAMP(8 * l + 6); }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_id == 1) { tmem_dealloc<Cfg::kTmemCols>(tmem_base); }
  if (threadIdx.x == 0) {  // the last CTA to leave re-arms every counter for the next launch
    __threadfence();
    if (atomicAdd(prm.sync + 15, 1u) == gridDim.x - 1) {
      for (int i = 0; i < FC_SYNC_WORDS; ++i) { prm.sync[i] = 0u; }
      __threadfence();
      if (prm.g.world > 0) {  // every CTA's logits are in every peer's buffer (each CTA fenced at system scope before it counted itself out): publish the step
        unsigned int *ctr = reinterpret_cast<unsigned int *>(prm.g.local_base + 32 * 4);
        unsigned int const step = *reinterpret_cast<volatile unsigned int *>(ctr) + 1u;
        __threadfence_system();
        for (int pr = 0; pr < prm.g.world; ++pr) {  // (one fence, then plain system-scope stores: a release per peer would repeat the fence)
          unsigned int *flag = reinterpret_cast<unsigned int *>(prm.g.peer_base[pr]) + prm.g.rank;
          asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
        }
        *ctr = step;
        __threadfence();
      }
    }
  }
}

}  // namespace b200
