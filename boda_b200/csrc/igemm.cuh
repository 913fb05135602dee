// igemm.cuh -- the tensor-core contraction kernel behind Boda's conv / k1conv / tconv / ipconv / sgemm functions
// (replaces test/rtc/{conv,tconv,k1conv,ipconv,sgemm}.cucl + src/cnn_codegen.cc's generated FMA loops).
//
//   D[P rows x Q rows] = sum_k  Pop[p, k] * Qop[q, k]            (both operands K-major, 16-bit, in "planes")
//
// * Pop is either a plain 2-d K-major matrix (sgemm, 1x1 convs, inner-product-shaped convs, swapped weights) read with
//   tiled TMA, or an NHWC activation tensor read with *im2col* TMA: the hardware walks 128 consecutive output pixels
//   (n,oy,ox) for one filter tap (ky,kx) and one 64-channel block, zero-filling padding taps.
// * Qop is always a plain 2-d K-major matrix (packed filters, or the activation matrix in swapped mode).
// * fp32-parity mode (kPlanes == 2): every fp32 operand x is carried as two fp16 planes, hi = fp16(s*x) and
//   lo = fp16(s*x - hi) (s a per-tensor power of two), and each k-step issues three tcgen05.mma:
//   hi*hi into the main TMEM accumulator, hi*lo + lo*hi into a separate cross-term accumulator (2^-11 the magnitude,
//   so the tensor core's truncating fp32 accumulation costs it nothing). That recovers ~22 bits of each product -- the
//   reference kernels are fp32 FFMA, and its compare tolerance (2e-4 .. 1e-3 mrd) cannot be met by one fp16/bf16/tf32 pass.
//   kPlanes == 1 is the single-pass fp16 / bf16 storage mode (BASELINE configs C3 / C4).
// * Tensor-core accumulation drift: TMEM accumulators are drained every `chunk_kblks` k-blocks into fp32 registers of
//   the epilogue warps (round-to-nearest adds) while the MMA warp continues into the second TMEM buffer.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_id % 4).
#pragma once
#include "umma.cuh"
#include "scale.cuh"
#include <cuda_fp16.h>
#include <cuda_bf16.h>

namespace b200 {

constexpr int IGEMM_BM = 128;      // rows of the P operand per CTA (UMMA M)
constexpr int IGEMM_BK = 64;       // K elements per k-block: 64 x 2 B = one 128-byte swizzle row
constexpr int IGEMM_UMMA_K = 16;   // K per tcgen05.mma for 16-bit operands
constexpr int IGEMM_THREADS = 192;

struct IgemmParams {
  // problem extents
  int p_rows;       // valid rows of P (output pixels; or out-chans when swapped)
  int q_rows;       // valid rows of Q
  int kblks_total;  // number of 64-wide k-blocks over the whole K
  int kblks_per_split;
  int chunk_kblks;  // drain TMEM into registers every this many k-blocks (>= kblks_per_split: single chunk)
  // P-operand addressing
  int p_im2col;     // 0: tiled 2-d {k, row}; 1: im2col
  int cblks;        // im2col: 64-channel blocks per filter tap
  int kw;           // im2col: filter width (taps are enumerated ky-major)
  int ow, ohw;      // im2col: output width, output pixels per image
  int sx, sy, px, py;
  // epilogue
  int swapped;      // 0: P rows = pixels, Q rows = channels; 1: P rows = channels, Q rows = pixels
  int out_chans;    // channels of the NCHW output
  int out_hw;       // pixels per image of the NCHW output
  int relu;
  int has_bias;
  long long split_stride;  // elements between split-K partial buffers (0 when writing final output)
  float *out;              // NCHW fp32 output, or split-K workspace
  float const *bias;       // per out-chan
  // residual join fused into the convolution (ResNet Eltwise SUM): out = max(floor, conv + bias + res), `res` an fp32
  // NCHW tensor with the output's dims, addressed like `out` (see igemm_acc_init). Non-swapped, non-split launches only (host-checked); null = off.
  float const *res;
  unsigned int const *res_absmax;  // max|res| published by its producer (bit pattern): the residual's term of the fp16 planes' output bound
  float const *p_scale;    // {scale, inv_scale} of the P tensor
  float const *q_scale;
  // split-K launches of the one-CTA kernel: with `tile_tickets` (one zeroed counter per output tile) the LAST of a tile's split CTAs to finish adds
  // the partial tiles in split order (+ bias, ReLU) and writes `out_final` -- the same arithmetic in the same order as splitk_reduce_kernel, without
  // the second launch. The counter is re-armed by that CTA. null = partials only (the reduce kernel follows).
  unsigned int *tile_tickets;
  float *out_final;
  long long *ts;  // experiments (debug_flags bit 4): per-cluster role stall counters of the CTA-pair kernel, [cluster][16] SM cycles (null = off)
  int debug;   // bit 0: skip TMA (MMA runs on whatever is in smem), bit 1: skip MMA issue (loads + barriers only), bit 2: skip the final global stores,
               // bit 3: skip the TMEM drains -- timing experiments only (results are garbage)
  int cm, cn;  // CTA-cluster shape: cm CTAs along P-tiles share (multicast) each Q tile, cn CTAs along Q-tiles share each P tile
  unsigned int *out_absmax;  // optional: publish max|out| (bit pattern) for the consumer's operand scaling
  uint32_t idesc;
  // K tail: k-blocks come in groups of kb_mod (the channel blocks of one filter tap; the whole K for 2-d operands); the last k-block of a
  // group holds only ksteps_last (1..4) 16-wide k-steps of real data -- the rest is zero padding in both operands and is not multiplied
  int kb_mod, ksteps_last;
  // k-block-major operand layout [k-block][rows][64] (packed filters): rows per k-block, 0 = plain row-major [rows][K]
  int p_kb_rows, q_kb_rows;
  int m_pair_tiles, q_tiles;  // persistent CTA-pair kernel (igemm2.cuh): 256-row tiles along P, BN-wide tiles along Q
  // bf16 storage mode, layout-transform elimination: besides the fp32 NCHW node the epilogue also writes the NHWC bf16
  // plane the consuming convolutions read ([pixel][out16_pitch] elements, already offset to this layer's first channel), so they skip
  // their activation pack. null = off.
  uint16_t *out16;
  int out16_pitch;
  // fp32-parity / fp16 modes: the same, with the consumer's hi (+ lo) fp16 planes. Their power-of-two scale must be known BEFORE this kernel
  // has seen its outputs, so it comes from an upper bound: |out| <= max_oc sum_k |w[oc,k]| * max|in| + max|bias|, with max|in| < 2^14 / s_in
  // from the input operand's own scale. A bound 2^5..2^7 above the true maximum costs nothing: fp16 is floating point, hi keeps 11 bits at any
  // magnitude and lo stays normal for every element within ~2^10 of the largest. Every CTA derives the same scale; CTA 0 publishes it.
  uint16_t *out16_lo;          // null in the single-plane modes
  float const *w_l1max;        // max over out chans of sum |w| (one float, computed once per weight version); null = out16 is a bf16 plane (no scale)
  float *out16_scale2;         // {scale, 1/scale} of the destination planes
  int n_bias;                  // number of biases (out chans) for the max|bias| term
  unsigned int const *in_absmax;  // optional: the TRUE max|in| its producer published (bit pattern); else 2^14 / s_in bounds it. Using the true
                                  // value keeps the overestimate from compounding along a chain of convolutions.
};

// TMEM columns reserved per accumulator buffer: BN rounded up to a power of two (BN = 96 accumulators sit at 128-column offsets)
__host__ __device__ constexpr uint32_t tmem_buf_cols(int bn) { return bn <= 32 ? 32u : bn <= 64 ? 64u : bn <= 128 ? 128u : 256u; }

template <int BN, int kPlanes>
struct IgemmCfg {
  static constexpr int kStageBytes = kPlanes * (IGEMM_BM * 128 + BN * 128);
  static constexpr int kMaxSmem = 220 * 1024;
  static constexpr int kStagesRaw = (kMaxSmem - 2048) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 1024 /*barriers*/;
  // TMEM columns: two ping-pong buffers for the main (hi*hi) accumulator + one for the cross-term (hi*lo + lo*hi)
  // accumulator in split mode; allocation must be a power of two >= 32
  static constexpr uint32_t kColsNeeded = (kPlanes == 2 ? 3 : 2) * tmem_buf_cols(BN);
  static constexpr uint32_t kTmemCols = kColsNeeded <= 32 ? 32 : kColsNeeded <= 64 ? 64 : kColsNeeded <= 128 ? 128 : kColsNeeded <= 256 ? 256 : 512;
};

// Final epilogue of one thread (= one P row) shared by the 1-CTA and 2-CTA kernels: out = max(floor, acc * inv + bias[ch]) for the BN
// channels (or pixels, when swapped) of this tile. `bias_s` is the tile's bias staged in shared memory (zeros when absent / split-K
// partial) and `floor` is 0 for ReLU, -inf otherwise, so the loop body has no data-dependent global load and no branch but the ragged-edge
// guard; addresses advance by a constant stride.
template <int BN>
__device__ __forceinline__ float igemm_store_row(float const (&acc)[BN], float inv, float const *bias_s, float floor_v, float *o, long long stride, int nvalid) {
  uint32_t const bias_sa = smem_u32(bias_s);
  if (nvalid >= BN) {
    // full tile (the common case): straight-line code -- bias read as 128-bit LDS with immediate offsets, four independent max chains, no
    // per-element branch (the guarded loop below serialises one LDS latency per element: measured 65 cycles/element, 8.4k cycles per tile)
    float am0 = 0.0f, am1 = 0.0f, am2 = 0.0f, am3 = 0.0f;
#pragma unroll
    for (int j = 0; j < BN; j += 4) {
      float b0, b1, b2, b3;
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_sa + 4 * j));
      float const v0 = fmaxf(fmaf(acc[j], inv, b0), floor_v), v1 = fmaxf(fmaf(acc[j + 1], inv, b1), floor_v);
      float const v2 = fmaxf(fmaf(acc[j + 2], inv, b2), floor_v), v3 = fmaxf(fmaf(acc[j + 3], inv, b3), floor_v);
      o[0] = v0; o[stride] = v1; o[2 * stride] = v2; o[3 * stride] = v3;
      o += 4 * stride;
      am0 = fmaxf(am0, fabsf(v0)); am1 = fmaxf(am1, fabsf(v1)); am2 = fmaxf(am2, fabsf(v2)); am3 = fmaxf(am3, fabsf(v3));
    }
    return fmaxf(fmaxf(am0, am1), fmaxf(am2, am3));
  }
  // ragged tile (out_chans not a multiple of BN): whole groups of four columns take the same straight-line path (nvalid is CTA-uniform, so
  // the branch does not diverge), the last partial group is guarded per element
  float amax = 0.0f;
#pragma unroll
  for (int j = 0; j < BN; j += 4) {
    if (j + 4 <= nvalid) {
      float b0, b1, b2, b3;
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b0), "=f"(b1), "=f"(b2), "=f"(b3) : "r"(bias_sa + 4 * j));
      float const v0 = fmaxf(fmaf(acc[j], inv, b0), floor_v), v1 = fmaxf(fmaf(acc[j + 1], inv, b1), floor_v);
      float const v2 = fmaxf(fmaf(acc[j + 2], inv, b2), floor_v), v3 = fmaxf(fmaf(acc[j + 3], inv, b3), floor_v);
      o[0] = v0; o[stride] = v1; o[2 * stride] = v2; o[3 * stride] = v3;
      amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v0), fabsf(v1))), fmaxf(fabsf(v2), fabsf(v3)));
    } else {
#pragma unroll
      for (int jj = j; jj < j + 4; ++jj) {
        if (jj < nvalid) {
          float b;
          asm("ld.shared.f32 %0, [%1];" : "=f"(b) : "r"(bias_sa + 4 * jj));  // LDS with an immediate offset (the generic pointer would compile to LD.E)
          float const v = fmaxf(fmaf(acc[jj], inv, b), floor_v);
          amax = fmaxf(amax, fabsf(v));
          o[(jj - j) * stride] = v;
        }
      }
    }
    o += 4 * stride;
  }
  return amax;
}

// Residual join (IgemmParams::res): the accumulators of a pixel row START at res / inv instead of zero, so the unchanged epilogue
// max(floor, acc * inv + bias) yields conv + bias + res. inv is a product of power-of-two operand scales, so the division is exact and the
// residual enters the fp32 sum like one more partial product; its BN loads are issued together at the top of the tile, before the first wait
// on the tensor pipe, and so overlap the main loop instead of serialising behind it. `r` = address of this row's first channel in `res`.
template <int BN>
__device__ __forceinline__ void igemm_acc_init(float (&acc)[BN], float const *r, long long stride, int nvalid, float inv_recip) {
  if (r == nullptr) {
#pragma unroll
    for (int j = 0; j < BN; ++j) { acc[j] = 0.0f; }
    return;
  }
  if (nvalid >= BN) {
#pragma unroll
    for (int j = 0; j < BN; ++j) { acc[j] = __ldg(r + j * stride) * inv_recip; }
  } else {
#pragma unroll
    for (int j = 0; j < BN; ++j) { acc[j] = (j < nvalid) ? __ldg(r + j * stride) * inv_recip : 0.0f; }
  }
}

// Second output of a pixel row in bf16 storage mode: the same values (scale, bias, floor) rounded to bf16 and written as 16-byte runs into
// the consumer's NHWC plane. `dst` is 16-byte aligned (channel offsets and pitches are multiples of 8 elements).
template <int BN>
__device__ __forceinline__ void igemm_store_row_bf16(float const (&acc)[BN], float inv, float const *bias_s, float floor_v, uint16_t *dst, int nvalid) {
  uint32_t const bias_sa = smem_u32(bias_s);
#pragma unroll
  for (int j = 0; j < BN; j += 8) {
    if (j < nvalid) {  // nvalid is a multiple of 8 on this path (host-checked)
      float b[8];
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]) : "r"(bias_sa + 4 * j));
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7]) : "r"(bias_sa + 4 * j + 16));
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float const v0 = fmaxf(fmaf(acc[j + 2 * k], inv, b[2 * k]), floor_v), v1 = fmaxf(fmaf(acc[j + 2 * k + 1], inv, b[2 * k + 1]), floor_v);
        __nv_bfloat162 const h = __floats2bfloat162_rn(v0, v1);
        w[k] = *reinterpret_cast<uint32_t const *>(&h);
      }
      *reinterpret_cast<uint4 *>(dst + j) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// fp16 hi (+ lo) planes of a pixel row, scaled by the bound-derived power of two `s_out` (see IgemmParams::w_l1max)
template <int BN>
__device__ __forceinline__ void igemm_store_row_split16(float const (&acc)[BN], float inv, float const *bias_s, float floor_v, float s_out, uint16_t *dst_hi,
                                                        uint16_t *dst_lo, int nvalid) {
  uint32_t const bias_sa = smem_u32(bias_s);
#pragma unroll
  for (int j = 0; j < BN; j += 8) {
    if (j < nvalid) {
      float b[8];
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[0]), "=f"(b[1]), "=f"(b[2]), "=f"(b[3]) : "r"(bias_sa + 4 * j));
      asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7]) : "r"(bias_sa + 4 * j + 16));
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float const t0 = fmaxf(fmaf(acc[j + 2 * k], inv, b[2 * k]), floor_v) * s_out, t1 = fmaxf(fmaf(acc[j + 2 * k + 1], inv, b[2 * k + 1]), floor_v) * s_out;
        __half2 const h = __floats2half2_rn(t0, t1);
        float2 const hf = __half22float2(h);
        __half2 const l = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
        wh[k] = *reinterpret_cast<uint32_t const *>(&h);
        wl[k] = *reinterpret_cast<uint32_t const *>(&l);
      }
      *reinterpret_cast<uint4 *>(dst_hi + j) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      if (dst_lo) { *reinterpret_cast<uint4 *>(dst_lo + j) = make_uint4(wl[0], wl[1], wl[2], wl[3]); }
    }
  }
}

// The destination planes' scale from the output bound, computed by the 128 epilogue threads of a CTA (named barrier 1); `red` = 4 floats of smem
__device__ __forceinline__ float igemm_out_scale(IgemmParams const &prm, float *red, int row /*0..127*/) {
  float bm = 0.0f;
  if (prm.has_bias) { for (int j = row; j < prm.n_bias; j += 128) { bm = fmaxf(bm, fabsf(__ldg(prm.bias + j))); } }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o)); }
  if ((row & 31) == 0) { red[row >> 5] = bm; }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  bm = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  float in_bound = 16384.0f * prm.p_scale[1];  // max|in| < 2^14 / s_in
  if (prm.in_absmax) { float const t = __uint_as_float(*prm.in_absmax); if (t > 0.0f) { in_bound = fminf(in_bound, t); } }
  float const res_bound = prm.res_absmax ? __uint_as_float(*prm.res_absmax) : 0.0f;  // host: fp16 planes beside a residual only with this cell
  float const bound = (__ldg(prm.w_l1max) * in_bound + bm + res_bound) * 1.01f;
  return scale_from_absmax_bits(__float_as_uint(bound));
}

template <int BN, int kPlanes>
__global__ void __launch_bounds__(IGEMM_THREADS, 1)
igemm_umma_kernel(const __grid_constant__ CUtensorMap p_hi_map, const __grid_constant__ CUtensorMap p_lo_map,
                  const __grid_constant__ CUtensorMap q_hi_map, const __grid_constant__ CUtensorMap q_lo_map,
                  const IgemmParams prm) {
  using Cfg = IgemmCfg<BN, kPlanes>;
  constexpr int kStages = Cfg::kStages;
  constexpr uint32_t kPBytes = IGEMM_BM * 128, kQBytes = BN * 128;

  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *bar_mem = smem + kStages * Cfg::kStageBytes;
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(bar_mem);
  uint64_t *empty_bar = full_bar + kStages;
  uint64_t *tmem_full_bar = empty_bar + kStages;   // [2]
  uint64_t *tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);
  float *bias_s = reinterpret_cast<float *>(bar_mem + 512);  // BN floats (the barrier block is 1 KB)

  int const warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();  // the next kernel of the forward pass may start launching; it waits for this grid before touching memory
  int const m0 = blockIdx.x * IGEMM_BM;
  int const n0 = blockIdx.y * BN;
  int const split = blockIdx.z;
  int const kb_begin = split * prm.kblks_per_split;
  int const kb_end = min(kb_begin + prm.kblks_per_split, prm.kblks_total);
  int const nkb = kb_end - kb_begin;
  int const chunk = prm.chunk_kblks;
  int const nchunks = (nkb + chunk - 1) / chunk;

  if (warp_id == 0 && lane == 0) {
    tma_prefetch_desc(&p_hi_map);
    tma_prefetch_desc(&q_hi_map);
    if (kPlanes == 2) { tma_prefetch_desc(&p_lo_map); tma_prefetch_desc(&q_lo_map); }
    // a stage is released by every CTA that consumed data this CTA multicast into it: cm + cn - 1 of them (itself included)
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], prm.cm + prm.cn - 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], 4); }
    fence_barrier_init();
  }
  if (warp_id == 1) { tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem); }
  tc_fence_before();
  bool const clustered = (prm.cm * prm.cn) > 1;
  if (clustered) { cluster_sync_all(); } else { __syncthreads(); }  // peers' barriers must exist before anyone multicasts into them
  tc_fence_after();
  pdl_wait();  // everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous kernel's tail
  uint32_t const tmem_base = *tmem_ptr_smem;
  // position inside the cluster and the multicast groups: CTAs with the same cx (same P tile) share P, same cy (same Q tile) share Q
  uint32_t const cx = clustered ? cluster_ctaid_x() : 0u, cy = clustered ? cluster_ctaid_y() : 0u;
  uint16_t mask_p = 0, mask_q = 0;
  for (int j = 0; j < prm.cn; ++j) { mask_p |= static_cast<uint16_t>(1u << (cx + j * prm.cm)); }
  for (int i = 0; i < prm.cm; ++i) { mask_q |= static_cast<uint16_t>(1u << (i + cy * prm.cm)); }
  int const p_rows_mine = IGEMM_BM / prm.cn, p_row0 = cy * p_rows_mine;  // the slice of the P tile this CTA fetches for its group
  int const q_rows_mine = BN / prm.cm, q_row0 = cx * q_rows_mine;

  if (warp_id == 0) {
    // ===================== TMA producer (whole warp walks the loop with warp-uniform state; one elected lane issues) ==========
    // (with `if (lane == 0)` around the loop the compiler cannot prove the operands of UTMALDG / UTCHMMA -- uniform-register instructions --
    // uniform and wraps each one in an ELECT / R2UR.BROADCAST retry loop: ~100 SASS instructions per k-block on one thread)
    {
      int img = 0, h_base = 0, w_base = 0;
      int const pm0 = m0 + p_row0;  // first P row (output pixel) of this CTA's slice
      if (prm.p_im2col) {
        img = pm0 / prm.ohw;
        int const rem = pm0 - img * prm.ohw;
        int const oy = rem / prm.ow, ox = rem - oy * prm.ow;
        h_base = oy * prm.sy - prm.py;
        w_base = ox * prm.sx - prm.px;
      }
      bool const mc_p = prm.cn > 1, mc_q = prm.cm > 1;
      // (tap, channel block) of the current k-block, advanced incrementally: a runtime integer division costs a single thread ~150 cycles
      int cb = 0, kx = 0, ky = 0;
      if (prm.p_im2col) { int const tap0 = kb_begin / prm.cblks; cb = kb_begin - tap0 * prm.cblks; ky = tap0 / prm.kw; kx = tap0 - ky * prm.kw; }
      for (int i = 0; i < ((prm.debug & 1) ? 0 : nkb); ++i) {
        int const kb = kb_begin + i;
        int const s = i % kStages;
        uint32_t const ph = (i / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (elect_one_sync()) {
        mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);  // the whole stage: own slices + the slices the group's other CTAs multicast in
        uint8_t *st = smem + s * Cfg::kStageBytes;
        uint8_t *p_hi = st + p_row0 * 128, *p_lo = p_hi + kPBytes;
        uint8_t *q_hi = st + kPlanes * kPBytes + q_row0 * 128, *q_lo = q_hi + kQBytes;
        if (prm.p_im2col) {
          if (mc_p) {
            tma_load_im2col_4d_mc(p_hi, &p_hi_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky, mask_p);
            if (kPlanes == 2) { tma_load_im2col_4d_mc(p_lo, &p_lo_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky, mask_p); }
          } else {
            tma_load_im2col_4d(p_hi, &p_hi_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky);
            if (kPlanes == 2) { tma_load_im2col_4d(p_lo, &p_lo_map, &full_bar[s], cb * IGEMM_BK, w_base, h_base, img, (uint16_t)kx, (uint16_t)ky); }
          }
        } else {
          int const c0 = prm.p_kb_rows ? 0 : kb * IGEMM_BK, c1 = pm0 + kb * prm.p_kb_rows;
          if (mc_p) {
            tma_load_2d_mc(p_hi, &p_hi_map, &full_bar[s], c0, c1, mask_p);
            if (kPlanes == 2) { tma_load_2d_mc(p_lo, &p_lo_map, &full_bar[s], c0, c1, mask_p); }
          } else {
            tma_load_2d(p_hi, &p_hi_map, &full_bar[s], c0, c1);
            if (kPlanes == 2) { tma_load_2d(p_lo, &p_lo_map, &full_bar[s], c0, c1); }
          }
        }
        {
          int const c0 = prm.q_kb_rows ? 0 : kb * IGEMM_BK, c1 = n0 + kb * prm.q_kb_rows;
          if (mc_q) {
            tma_load_2d_mc(q_hi, &q_hi_map, &full_bar[s], c0, c1 + q_row0, mask_q);
            if (kPlanes == 2) { tma_load_2d_mc(q_lo, &q_lo_map, &full_bar[s], c0, c1 + q_row0, mask_q); }
          } else {
            tma_load_2d(q_hi, &q_hi_map, &full_bar[s], c0, c1);
            if (kPlanes == 2) { tma_load_2d(q_lo, &q_lo_map, &full_bar[s], c0, c1); }
          }
        }
        }
        __syncwarp();
        if (++cb == prm.cblks) { cb = 0; if (++kx == prm.kw) { kx = 0; ++ky; } }
      }
    }
  } else if (warp_id == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    {
      uint32_t const idesc = prm.idesc;
      int const kb_mod = prm.kb_mod, ksteps_last = prm.ksteps_last;
      int kb_in_grp = kb_begin % kb_mod;  // once per CTA
      int i = 0;
      for (int c = 0; c < nchunks; ++c) {
        int const buf = c & 1;
        if (!(prm.debug & 8)) { mbar_wait(&tmem_empty_bar[buf], ((c >> 1) & 1) ^ 1); }
        tc_fence_after();
        uint32_t const tmem_d = tmem_base + buf * tmem_buf_cols(BN);
        uint32_t const tmem_x = tmem_base + 2 * tmem_buf_cols(BN);  // cross-term accumulator: 2^-11 of the main one, drained once at the end
        int const i_end = min(i + chunk, nkb);
        bool first = true;
        for (; i < i_end; ++i) {
          int const s = i % kStages;
          uint32_t const ph = (i / kStages) & 1;
          if (!(prm.debug & 1)) { mbar_wait(&full_bar[s], ph); }
          tc_fence_after();
          uint32_t const st = smem_u32(smem + s * Cfg::kStageBytes);
          uint32_t const p_hi = sw128_desc_lo(st), p_lo = sw128_desc_lo(st + kPBytes);
          uint32_t const q_hi = sw128_desc_lo(st + kPlanes * kPBytes), q_lo = sw128_desc_lo(st + kPlanes * kPBytes + kQBytes);
          int nk = IGEMM_BK / IGEMM_UMMA_K;
          if (++kb_in_grp == kb_mod) { kb_in_grp = 0; nk = ksteps_last; }
          if (prm.debug & 2) { nk = 0; }
          if (elect_one_sync()) {
            issue_kblock<kPlanes, false>(tmem_d, tmem_x, p_hi, p_lo, q_hi, q_lo, idesc, first ? 0u : 1u, i == 0 ? 0u : 1u, nk);
            // frees this smem stage -- in every CTA that multicast a slice into it -- once the MMAs above have read it
            if (clustered) { umma_commit_mc(&empty_bar[s], static_cast<uint16_t>(mask_p | mask_q)); } else { umma_commit(&empty_bar[s]); }
            if (i == i_end - 1) { umma_commit(&tmem_full_bar[buf]); }  // accumulator chunk complete
          }
          __syncwarp();
          first = false;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    int const q = warp_id & 3;  // TMEM lane quarter this warp may read
    int const row = q * 32 + lane;
    // stage this tile's bias in shared memory while the main loop runs (zeros when absent, for split-K partials, or in swapped mode)
    for (int j = row; j < BN; j += 128) {
      bool const use = prm.has_bias && prm.split_stride == 0 && !prm.swapped && (n0 + j) < prm.q_rows;
      bias_s[j] = use ? __ldg(prm.bias + n0 + j) : 0.0f;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");  // epilogue warps only
    float acc[BN];
    {
      float const *res_row = nullptr;  // (host: a residual input only on non-swapped, non-split launches)
      if (prm.res && m0 + row < prm.p_rows) {
        int const img = (m0 + row) / prm.out_hw, pix = (m0 + row) - img * prm.out_hw;
        res_row = prm.res + (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + pix;
      }
      igemm_acc_init<BN>(acc, res_row, prm.out_hw, prm.q_rows - n0, prm.p_scale[0] * prm.q_scale[0]);
    }
    for (int c = 0; c < ((prm.debug & 8) ? 0 : nchunks); ++c) {
      int const buf = c & 1;
      mbar_wait(&tmem_full_bar[buf], (c >> 1) & 1);
      tc_fence_after();
      uint32_t const taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * tmem_buf_cols(BN);
#pragma unroll
      for (int j0 = 0; j0 < BN; j0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(taddr + j0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { acc[j0 + j] += __uint_as_float(r[j]); }
      }
      if (kPlanes == 2 && c == nchunks - 1) {  // the last commit also covers every cross-term MMA
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
      if (lane == 0) { mbar_arrive(&tmem_empty_bar[buf]); }
    }
    // ---- write out: NCHW fp32 (or split-K partial) ----
    float const inv = prm.p_scale[1] * prm.q_scale[1];
    bool const final_out = (prm.split_stride == 0);
    float s_out = 1.0f;
    if (prm.out16 && prm.w_l1max) {  // (host: only for non-swapped, non-split launches) -- see IgemmParams::w_l1max
      s_out = igemm_out_scale(prm, reinterpret_cast<float *>(bar_mem + 256), row);
      if (blockIdx.x == 0 && blockIdx.y == 0 && row == 0) { prm.out16_scale2[0] = s_out; prm.out16_scale2[1] = 1.0f / s_out; }
    }
    float const floor_v = (final_out && prm.relu) ? 0.0f : -INFINITY;
    int const prow = m0 + row;
    float amax = 0.0f;
    if (prow < prm.p_rows && !(prm.debug & 4)) {
      float *outp = prm.out + static_cast<long long>(split) * prm.split_stride;
      if (!prm.swapped) {  // row = pixel, columns = channels (bias per column, staged in smem by the prologue)
        int const img = prow / prm.out_hw, pix = prow - img * prm.out_hw;
        float *o = outp + (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + pix;
        amax = igemm_store_row<BN>(acc, inv, bias_s, floor_v, o, prm.out_hw, prm.q_rows - n0);
        if (prm.out16 && final_out) {
          long long const o16 = static_cast<long long>(prow) * prm.out16_pitch + n0;
          if (prm.w_l1max) { igemm_store_row_split16<BN>(acc, inv, bias_s, floor_v, s_out, prm.out16 + o16, prm.out16_lo ? prm.out16_lo + o16 : nullptr, prm.q_rows - n0); }
          else { igemm_store_row_bf16<BN>(acc, inv, bias_s, floor_v, prm.out16 + o16, prm.q_rows - n0); }
        }
      } else {  // row = channel, columns = pixels
        int const ch = prow;
        float const b = (final_out && prm.has_bias) ? __ldg(prm.bias + ch) : 0.0f;
        // (img, pix) of column j advance by carries: one integer division per row, not per element -- with it the 64 stores of an
        // inner-product layer's tile took 9 us, half of the kernel (tools/fc_experiments.py)
        int const img0 = n0 / prm.out_hw;
        int pix = n0 - img0 * prm.out_hw;
        long long off = (static_cast<long long>(img0) * prm.out_chans + ch) * prm.out_hw + pix;
        long long const img_step = static_cast<long long>(prm.out_chans) * prm.out_hw - (prm.out_hw - 1);  // last pixel of an image -> first of the next
#pragma unroll
        for (int j = 0; j < BN; ++j) {
          if (n0 + j < prm.q_rows) {
            float const v = fmaxf(fmaf(acc[j], inv, b), floor_v);
            amax = fmaxf(amax, fabsf(v));
            outp[off] = v;
          }
          if (++pix == prm.out_hw) { pix = 0; off += img_step; } else { off += 1; }
        }
      }
    }
    if (!final_out && prm.tile_tickets) {
      // ---- fused split-K reduction, spread over the tile's split CTAs: once all `splits` partial tiles are written, CTA z sums columns
      // [z*BN/splits, ...) of the tile over the splits in split order (+ bias, ReLU) -- the arithmetic of splitk_reduce_kernel without its launch.
      // (A first version let the LAST CTA reduce the whole tile: 8 dependent L2 round trips per column group on 8 CTAs made fc8 20 us slower.)
      // All CTAs of a split-K launch are co-resident (tiles * splits <= #SMs by plan), so the bounded spin below cannot deadlock.
      unsigned int *ticket = prm.tile_tickets + 2 * (blockIdx.y * gridDim.x + blockIdx.x);  // {arrived, reduced}
      unsigned int const splits = gridDim.z;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (row == 0) {
        atomicAdd(ticket, 1u);
        unsigned int v = 0, spins = 0;
        while (true) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ticket) : "memory");
          if (v >= splits) { break; }
          __nanosleep(64);
          if (++spins > (1u << 22)) { printf("b200: split-K partial tiles never completed (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z); __trap(); }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      __threadfence();
      float const floor_f = prm.relu ? 0.0f : -INFINITY;
      int const nvalid = min(BN, prm.q_rows - n0);
      int const per = (BN + static_cast<int>(splits) - 1) / static_cast<int>(splits);
      int const j_begin = static_cast<int>(blockIdx.z) * per, j_end = min(nvalid, j_begin + per);
      amax = 0.0f;
      if (prow < prm.p_rows && !(prm.debug & 4)) {
        float b_row = 0.0f;
        long long off0 = 0, img_step = 0;
        int pix0 = 0;
        if (!prm.swapped) {  // row = pixel, columns = channels: column j at off0 + j * out_hw
          int const img = prow / prm.out_hw, pix = prow - img * prm.out_hw;
          off0 = (static_cast<long long>(img) * prm.out_chans + n0) * prm.out_hw + pix;
        } else {             // row = channel, columns = pixels
          b_row = prm.has_bias ? __ldg(prm.bias + prow) : 0.0f;
          int const p0 = n0 + j_begin, img0 = p0 / prm.out_hw;
          pix0 = p0 - img0 * prm.out_hw;
          off0 = (static_cast<long long>(img0) * prm.out_chans + prow) * prm.out_hw + pix0;
          img_step = static_cast<long long>(prm.out_chans) * prm.out_hw - (prm.out_hw - 1);
        }
        for (int j0 = j_begin; j0 < j_end; j0 += 4) {  // 4 columns x all splits in flight per thread, then the adds in split order
          long long offs[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (!prm.swapped) { offs[j] = off0 + static_cast<long long>(j0 + j) * prm.out_hw; }
            else { offs[j] = off0; if (++pix0 == prm.out_hw) { pix0 = 0; off0 += img_step; } else { off0 += 1; } }
          }
          float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 4
          for (unsigned int sp = 0; sp < splits; ++sp) {
            float const *w = prm.out + static_cast<long long>(sp) * prm.split_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) { if (j0 + j < j_end) { v[j] += __ldcg(w + offs[j]); } }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j0 + j < j_end) {
              float r = v[j];
              if (!prm.swapped) { if (prm.has_bias) { r += __ldg(prm.bias + n0 + j0 + j); } } else { r += b_row; }
              r = fmaxf(r, floor_f);
              amax = fmaxf(amax, fabsf(r));
              prm.out_final[offs[j]] = r;
            }
          }
        }
      }
      if (prm.out_absmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
        if (lane == 0 && amax > 0.0f) { atomicMax(prm.out_absmax, __float_as_uint(amax)); }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (row == 0) {  // the last CTA to finish its share re-arms both counters for the next launch
        if (atomicAdd(ticket + 1, 1u) == splits - 1) { ticket[1] = 0u; __threadfence(); ticket[0] = 0u; }
      }
    }
    if (prm.out_absmax && prm.split_stride == 0) {  // warp-uniform branch; all 32 lanes take part in the shuffle reduce
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o)); }
      if (lane == 0 && amax > 0.0f) { atomicMax(prm.out_absmax, __float_as_uint(amax)); }
    }
  }

  tc_fence_before();
  if (clustered) { cluster_sync_all(); } else { __syncthreads(); }  // peers may still be arriving on this CTA's barriers
  if (warp_id == 1) { tmem_dealloc<Cfg::kTmemCols>(tmem_base); }
}

// split-K fix-up: out[i] = relu( sum_s ws[s][i] + bias[chan(i)] )   (deterministic order). `out_img_stride` != 0: the result is a channel
// slice of a wider NCHW var (concat by offset) -- image `img` of the slice starts at out + img * out_img_stride.
__global__ void splitk_reduce_kernel(float const *__restrict__ ws, float *__restrict__ out, float const *__restrict__ bias,
                                     long long n, int splits, int out_chans, int out_hw, int relu, unsigned int *out_absmax, long long out_img_stride) {
  pdl_prologue();
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float v = 0.0f;
  if (i < n) {
    // partial sums are added in split order (deterministic); four loads in flight at a time instead of one load -> add chain per split
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
      float const a0 = ws[s * n + i], a1 = ws[(s + 1) * n + i], a2 = ws[(s + 2) * n + i], a3 = ws[(s + 3) * n + i];
      v += a0; v += a1; v += a2; v += a3;
    }
    for (; s < splits; ++s) { v += ws[s * n + i]; }
    if (bias) {
      long long const c = (n < (1ll << 31)) ? static_cast<long long>((static_cast<unsigned>(i) / static_cast<unsigned>(out_hw)) % static_cast<unsigned>(out_chans))
                                            : (i / out_hw) % out_chans;
      v += __ldg(bias + c);
    }
    if (relu) { v = fmaxf(v, 0.0f); }
    long long o = i;
    if (out_img_stride) { long long const per_img = static_cast<long long>(out_chans) * out_hw, img = i / per_img; o = img * out_img_stride + (i - img * per_img); }
    out[o] = v;
  }
  if (out_absmax) {
    float m = (i < n) ? fabsf(v) : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
    if ((threadIdx.x & 31) == 0 && m > 0.0f) { atomicMax(out_absmax, __float_as_uint(m)); }
  }
}

}  // namespace b200
