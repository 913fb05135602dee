// pointwise.cuh -- the bandwidth-bound functions of Boda's rtc_fwd path as hand-written sm_100a kernels:
// gen_data_* (test inputs), pool, lrn, relu, softmax, copy (Concat), reduce (N-ary sum), plus the operand
// "pack" kernels (abs-max -> power-of-two scale -> transpose + fp16 hi/lo split) that feed the tensor-core kernel.
// All activations are fp32 NCHW (`img:chan:y:x`) exactly as in the reference (src/conv_util.cc:482-503).
#pragma once
#include <cstdint>
#include <cfloat>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "pdl.cuh"
#include "scale.cuh"

namespace b200 {

// ---- abs-max side channel -------------------------------------------------------------------------------------
// Producers of a tensor that a convolution will consume can publish max|x| into a uint32 cell (non-negative floats order like
// their bit patterns), which saves the consumer a separate reduction pass before it picks its power-of-two operand scale.
__device__ __forceinline__ void publish_absmax_warp(float m, unsigned int *cell) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
  if ((threadIdx.x & 31) == 0 && m > 0.0f) {
    unsigned int const bits = __float_as_uint(m);
    if (bits > *reinterpret_cast<volatile unsigned int *>(cell)) { atomicMax(cell, bits); }
  }
}
// ---- gen_data (test/rtc/gen-util.h:1-9 and test/rtc/gen_data_*.cucl) -----------------------------------------
__device__ __forceinline__ float det_hash_rand(uint32_t rv) {
  uint32_t h = rv;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  return __fmaf_rn((float)(h), (10.0f / (float)(0xFFFFFFFFu)), -5.0f);  // one fused op, as NVRTC's fmad=true compiles the reference
}

// kind: 0 = Convolution_in / Convolution_filts style 4-d (x,y modes), 1 = biases, 2 = sgemm_a (K:M), 3 = sgemm_b (K:N)
__global__ void gen_data_kernel(float *__restrict__ dst, uint32_t n, int kind, uint32_t inner /*x | M | N*/,
                                uint32_t inner2 /*y | K*/, uint32_t mode, float vi, uint32_t salt) {
  pdl_prologue();
  uint32_t const i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) { return; }
  float val = vi;
  if (kind == 0) {
    uint32_t const x = i % inner, y = (i / inner) % inner2;
    if (mode == 2) { val += (float)x; }
    if (mode == 3) { val += (float)y; }
    else if (mode == 4) { if ((x == inner / 2) && (y == inner2 / 2)) { val += 1.0f; } }
    else if (mode == 5) { val += det_hash_rand(i + salt); }
  } else if (kind == 1) {
    if (mode == 5) { val += det_hash_rand(i + salt); }
  } else if (kind == 2) {
    uint32_t fin_mode = mode;
    if (fin_mode >= 100) { fin_mode = fin_mode / 100; }
    uint32_t const m = i % inner, k = i / inner;
    if (fin_mode == 2) { val += (float)m; }
    if (fin_mode == 3) { val += (float)k; }
    else if (fin_mode == 4) { if ((m == inner / 2) && (k == inner2 / 2)) { val += 1.0f; } }
    else if (fin_mode == 5) { val += det_hash_rand(i + salt); }
    else if (fin_mode == 6) { val += (float)(m * 1000 + k); }
  } else {
    uint32_t const nn = i % inner, k = i / inner;
    if (mode == 2) { val += (float)nn; }
    if (mode == 3) { val += (float)k; }
    else if (mode == 4) { if ((nn == inner / 2) && (k == inner2 / 2)) { val += 1.0f; } }
    else if (mode == 5) { val += det_hash_rand(i + salt); }
    else if (mode >= 100) { if (nn == k) { val += 1.0f; } }
  }
  dst[i] = val;
}

// ---- relu (test/rtc/relu.cucl) --------------------------------------------------------------------------------
__global__ void relu_kernel(float *__restrict__ x, long long n) {
  pdl_prologue();
  long long const i4 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i4 + 3 < n && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
    float4 v = *reinterpret_cast<float4 *>(x + i4);
    v.x = (v.x <= 0) ? 0.0f : v.x; v.y = (v.y <= 0) ? 0.0f : v.y; v.z = (v.z <= 0) ? 0.0f : v.z; v.w = (v.w <= 0) ? 0.0f : v.w;
    *reinterpret_cast<float4 *>(x + i4) = v;
  } else {
    for (long long i = i4; i < n && i < i4 + 4; ++i) { x[i] = (x[i] <= 0) ? 0.0f : x[i]; }
  }
}

// ---- copy / Concat (test/rtc/copy.cucl) -----------------------------------------------------------------------
// in: [N][C][HW] -> out[:, ocix:ocix+C]; per image the source block is contiguous, so move 128-bit words when aligned.
__global__ void concat_copy_kernel(float *__restrict__ in, float *__restrict__ out, long long per_img /*C*HW*/,
                                   long long out_img_stride, long long out_off, int n_img, int vec4, unsigned int *out_absmax, int reverse) {
  pdl_prologue();
  long long const total = per_img * n_img;
  float m = 0.0f;
  if (vec4) {
    long long const i4 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
    if (i4 < total) {
      long long const img = i4 / per_img, r = i4 - img * per_img;
      float4 *ip = reinterpret_cast<float4 *>(in + i4), *op = reinterpret_cast<float4 *>(out + img * out_img_stride + out_off + r);
      if (reverse) { *ip = *op; }  // extract: in = out[:, ocix:ocix+C]
      else {
        float4 const v = *ip;
        *op = v;
        m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
      }
    }
  } else {
    long long const i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < total) {
      long long const img = i / per_img, r = i - img * per_img;
      if (reverse) { in[i] = out[img * out_img_stride + out_off + r]; }
      else {
        float const v = in[i];
        out[img * out_img_stride + out_off + r] = v;
        m = fabsf(v);
      }
    }
  }
  if (out_absmax) { publish_absmax_warp(m, out_absmax); }
}

// ---- reduce: N-ary elementwise sum (test/rtc/reduce.cucl) ------------------------------------------------------
struct ReduceArgs { float const *ins[8]; int ins_num; };
// relu != 0 fuses the in-place ReLU that follows a residual join (Eltwise SUM + ReLU in ResNet); the sum order is the reference's.
// 128-bit loads / stores when every pointer is 16-byte aligned (n4 = n / 4 vector elements, then a scalar tail), grid-stride so a CTA
// publishes max|out| once (one atomic per warp-reduction, not one per 32 elements).
__global__ void __launch_bounds__(256)
reduce_sum_kernel(ReduceArgs a, float *__restrict__ out, long long n, int relu, unsigned int *out_absmax, int vec4) {
  pdl_prologue();
  long long const tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x, nthr = static_cast<long long>(gridDim.x) * blockDim.x;
  float m = 0.0f;
  long long done = 0;
  if (vec4) {
    long long const n4 = n >> 2;
    constexpr int kU = 4;  // four independent 128-bit loads per input in flight per thread
    for (long long i0 = tid; i0 < n4; i0 += nthr * kU) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) { v[u] = make_float4(0.f, 0.f, 0.f, 0.f); }
      for (int j = 0; j < a.ins_num; ++j) {
        float4 x[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) { long long const i = i0 + u * nthr; x[u] = (i < n4) ? __ldg(reinterpret_cast<float4 const *>(a.ins[j]) + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < kU; ++u) { v[u].x += x[u].x; v[u].y += x[u].y; v[u].z += x[u].z; v[u].w += x[u].w; }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        long long const i = i0 + u * nthr;
        if (i < n4) {
          float4 w = v[u];
          if (relu) { w.x = (w.x <= 0) ? 0.0f : w.x; w.y = (w.y <= 0) ? 0.0f : w.y; w.z = (w.z <= 0) ? 0.0f : w.z; w.w = (w.w <= 0) ? 0.0f : w.w; }
          reinterpret_cast<float4 *>(out)[i] = w;
          m = fmaxf(m, fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fmaxf(fabsf(w.z), fabsf(w.w))));
        }
      }
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nthr) {
    float v = 0;
    for (int j = 0; j < a.ins_num; ++j) { v += __ldg(a.ins[j] + i); }
    if (relu) { v = (v <= 0) ? 0.0f : v; }
    out[i] = v;
    m = fmaxf(m, fabsf(v));
  }
  if (out_absmax) { publish_absmax_warp(m, out_absmax); }
}

// ---- filter row L1 norm (parameter-only): max over out chans of sum_k |w[oc, k]|, for the output bound of IgemmParams::w_l1max. grid = out chans.
__global__ void __launch_bounds__(256)
filts_l1max_kernel(float const *__restrict__ filts, long long per_oc, unsigned int *__restrict__ out_bits) {
  pdl_prologue();
  float const *w = filts + static_cast<long long>(blockIdx.x) * per_oc;
  float s = 0.0f;
  for (long long i = threadIdx.x; i < per_oc; i += 256) { s += fabsf(__ldg(w + i)); }
  __shared__ float warp_s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); }
  if ((threadIdx.x & 31) == 0) { warp_s[threadIdx.x >> 5] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int k = 0; k < 8; ++k) { t += warp_s[k]; }
    atomicMax(out_bits, __float_as_uint(t));  // non-negative floats order like their bit patterns
  }
}

// ---- BatchNorm / Scale folding (parameter-only) ---------------------------------------------------------------
// conv -> BatchNorm(use_global_stats) -> Scale is   y = gamma * ((W.x + bias) - mean/sf) / sqrt(var/sf + eps) + beta   (Caffe semantics:
// the stored mean / var blobs are divided by the moving-average scale factor blob sf, 0 -> 0), i.e. a convolution with
//   W'[oc] = W[oc] * a[oc],  b'[oc] = (bias[oc] - mean[oc]/sf) * a[oc] + beta[oc],  a[oc] = gamma[oc] / sqrt(var[oc]/sf + eps).
// Any of {biases, (mean,var,sf), (gamma,beta)} may be absent (null). grid = (ceil(per_oc / 256), out_chans).
__global__ void bn_fold_kernel(float const *__restrict__ filts, float const *__restrict__ biases, float const *__restrict__ mean, float const *__restrict__ var,
                               float const *__restrict__ sf, float const *__restrict__ gamma, float const *__restrict__ beta, float eps,
                               float *__restrict__ out_filts, float *__restrict__ out_biases, long long per_oc) {
  pdl_prologue();
  int const oc = blockIdx.y;
  float s = 1.0f, m = 0.0f, inv_std = 1.0f;
  if (mean) {
    float const f = __ldg(sf);
    s = (f == 0.0f) ? 0.0f : __fdiv_rn(1.0f, f);
    m = __ldg(mean + oc) * s;
    inv_std = __fdiv_rn(1.0f, __fsqrt_rn(__ldg(var + oc) * s + eps));
  }
  float const a = (gamma ? __ldg(gamma + oc) : 1.0f) * inv_std;
  long long const i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < per_oc) { out_filts[oc * per_oc + i] = __ldg(filts + oc * per_oc + i) * a; }
  if (i == 0) { out_biases[oc] = ((biases ? __ldg(biases + oc) : 0.0f) - m) * a + (beta ? __ldg(beta + oc) : 0.0f); }
}

// ---- pool (test/rtc/pool.cucl:13-40) --------------------------------------------------------------------------
__global__ void pool_kernel(float const *__restrict__ in, float *__restrict__ out, long long n_out, int H, int W, int OH,
                            int OW, int KH, int KW, int sy, int sx, int py, int px, int avg_pool, unsigned int *out_absmax) {
  pdl_prologue();
  long long const i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  float out_v = 0.0f;
  if (i < n_out) {
    int const ox = static_cast<int>(i % OW), oy = static_cast<int>((i / OW) % OH);
    long long const plane = i / (static_cast<long long>(OW) * OH);
    float const *ip = in + plane * H * W;
    out_v = avg_pool ? 0.0f : -FLT_MAX;
    float avg_pool_sz = 0;
    for (int kx = 0; kx != KW; ++kx) {
      for (int ky = 0; ky != KH; ++ky) {
        int const in_y = oy * sy + ky - py, in_x = ox * sx + kx - px;
        if (in_y >= 0 && in_x >= 0 && in_x < W && in_y < H) {
          float const v = __ldg(ip + in_y * W + in_x);
          if (avg_pool) { out_v += v; avg_pool_sz += 1; }
          else if (v > out_v) { out_v = v; }
        }
      }
    }
    if (avg_pool) { out_v = __fdiv_rn(out_v, avg_pool_sz); }
    out[i] = out_v;
  }
  if (out_absmax) { publish_absmax_warp((i < n_out) ? fabsf(out_v) : 0.0f, out_absmax); }  // whole warp reaches this point together
}

// Fixed-window variant (compile-time kernel / stride, e.g. the 3x3 stride-2 pools of AlexNet / NiN / GoogLeNet): grid.y walks the
// (img,chan) planes so no thread does a 64-bit divide, the window is fully unrolled (all taps in flight), same tap order as above.
template <int K, int S>
__global__ void __launch_bounds__(256)
pool_kernel_fixed(float const *__restrict__ in, float *__restrict__ out, int H, int W, int OH, int OW, int py, int px, int avg_pool,
                  unsigned int *out_absmax) {
  pdl_prologue();
  int const p = blockIdx.x * blockDim.x + threadIdx.x;
  long long const plane = blockIdx.y + static_cast<long long>(blockIdx.z) * gridDim.y;
  bool const valid = p < OH * OW;
  float out_v = 0.0f;
  if (valid) {
    int const oy = p / OW, ox = p - oy * OW;
    float const *ip = in + plane * H * W;
    int const y0 = oy * S - py, x0 = ox * S - px;
    float v[K][K];  // [kx][ky]
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        int const in_y = y0 + ky, in_x = x0 + kx;
        bool const ok = in_y >= 0 && in_x >= 0 && in_x < W && in_y < H;
        v[kx][ky] = ok ? __ldg(ip + in_y * W + in_x) : (avg_pool ? 0.0f : -FLT_MAX);
      }
    }
    if (avg_pool) {
      float cnt = 0;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          int const in_y = y0 + ky, in_x = x0 + kx;
          if (in_y >= 0 && in_x >= 0 && in_x < W && in_y < H) { out_v += v[kx][ky]; cnt += 1; }
        }
      }
      out_v = __fdiv_rn(out_v, cnt);
    } else {
      out_v = -FLT_MAX;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) { out_v = fmaxf(out_v, v[kx][ky]); }
      }
    }
    out[plane * OH * OW + p] = out_v;
  }
  if (out_absmax) { publish_absmax_warp(valid ? fabsf(out_v) : 0.0f, out_absmax); }
}

// optional second output of pool_plane_kernel: the consumer convolution's NHWC 16-bit planes (see the end of the kernel)
struct PoolPlanes {
  uint16_t *hi, *lo;              // null hi = off; lo only in the fp32-parity mode
  float *scale2;                  // {scale, 1/scale} of the planes
  unsigned int const *in_absmax;  // max|in| bit pattern (fp16 planes); null = bf16 planes, scale 1
  int C, cpad, bf16;              // channels of the pooled node, its padded NHWC pitch, storage type
  int OW, dHp, dWp, dpy, dpx;     // dWp != 0: the planes are in the shared-padding layout of a halo-mode consumer (igemm4.cuh): pixel (y, x) of image n at
                                  // row (n*dHp + y + dpy)*dWp + x + dpx
};

// Plane-group variant: one CTA stages kPPC consecutive (img,chan) planes -- a CONTIGUOUS run of kPPC*H*W floats -- in shared memory with
// 128-bit loads and computes their kPPC*OH*OW outputs (also contiguous) from there, so the overlapping stride-S window reads never touch
// L1/L2 and both global streams are long coalesced runs. Small planes (13x13, 27x27) would otherwise mean thousands of tiny CTAs that
// are all ramp-up and no bandwidth (measured: 15 us for the 6.7 MB of AlexNet pool5 with one plane per CTA). The host picks planes-per-CTA
// (a multiple of 4, which keeps every run 16-byte aligned) so that the tile fits the dynamic shared memory it requests.
// Compute phase of the plane-group pool kernels: the np planes staged at plane_s ([np][H][W]) -> op ([np][OH][OW], global) and, when
// `stage` is non-null, a shared-memory copy of the outputs for the NHWC plane write. Returns this thread's max|out|.
template <int K, int S>
__device__ __forceinline__ float pool_planes_compute(float const *plane_s, float *stage, float *op, int np, int H, int W, int OH, int OW, int py, int px, int avg_pool) {
  int const hw = H * W, ohw = OH * OW, n_out = np * ohw;
  float amax = 0.0f;
  if (!avg_pool && S == 1) {
    // Stride-1 max pooling (GoogLeNet's inception pools): a work item is a run of kC outputs of one output row. It folds the K input rows
    // into kC + K - 1 column maxima in registers (every input is loaded once per item, bounds tests hoisted per column / per row), then
    // each output is the max of K neighbouring column maxima. Max is exact in any order, so this equals the reference's tap loop bit for
    // bit. (The column-walk form below measured 98 instructions per output on 14x14 planes -- its per-column setup is not amortised over
    // short columns -- and ran at 75 % issue utilisation, 11 % of DRAM peak: ncu, profiles/ncu_r01_googlenet_pool_fc.md.)
    constexpr int kC = 8, kIn = kC + K - 1;
    int const ncx = (OW + kC - 1) / kC, items = np * OH * ncx;
    for (int it = threadIdx.x; it < items; it += 256) {
      int const t = it / ncx, cx = it - t * ncx, pl = t / OH, oy = t - pl * OH;
      int const ox0 = cx * kC, x_in0 = ox0 - px, y0 = oy - py;
      bool xok[kIn];
#pragma unroll
      for (int c = 0; c < kIn; ++c) { xok[c] = static_cast<unsigned>(x_in0 + c) < static_cast<unsigned>(W); }
      float vm[kIn];
#pragma unroll
      for (int c = 0; c < kIn; ++c) { vm[c] = -FLT_MAX; }
      float const *pb = plane_s + pl * hw + x_in0;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        if (static_cast<unsigned>(y0 + ky) < static_cast<unsigned>(H)) {
          float const *pr = pb + (y0 + ky) * W;
#pragma unroll
          for (int c = 0; c < kIn; ++c) { if (xok[c]) { vm[c] = fmaxf(vm[c], pr[c]); } }
        }
      }
      int const obase = pl * ohw + oy * OW + ox0;
#pragma unroll
      for (int o = 0; o < kC; ++o) {
        if (ox0 + o < OW) {
          float out_v = vm[o];
#pragma unroll
          for (int k = 1; k < K; ++k) { out_v = fmaxf(out_v, vm[o + k]); }
          op[obase + o] = out_v;
          if (stage) { stage[obase + o] = out_v; }
          amax = fmaxf(amax, fabsf(out_v));
        }
      }
    }
  } else if (!avg_pool) {
    // Max pooling: a work item is one output column (plane, ox) over a segment of kSeg output rows. It keeps the horizontal maxima of the K
    // input rows under the current window in registers, so moving down one output row costs S new rows of K shared-memory loads and the
    // K - S others are reused; the x bounds tests are per item, the y test per row. Max is exact in any order, so this equals the
    // reference's tap loop bit for bit (~20 instructions per output instead of ~140: the tap loop was issue-bound, ncu).
    constexpr int kSeg = 8;
    int const nseg = (OH + kSeg - 1) / kSeg, items = np * nseg * OW;
    for (int it = threadIdx.x; it < items; it += 256) {
      int const t = it / OW, ox = it - t * OW, pl = t / nseg, seg = t - pl * nseg;
      int const oy0 = seg * kSeg, oy1 = min(oy0 + kSeg, OH), x0 = ox * S - px;
      bool xok[K];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) { xok[kx] = static_cast<unsigned>(x0 + kx) < static_cast<unsigned>(W); }
      float const *pc = plane_s + pl * hw + x0;
      auto hrow = [&](int y) -> float {
        float m = -FLT_MAX;
        if (static_cast<unsigned>(y) < static_cast<unsigned>(H)) {
          float const *pr = pc + y * W;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) { float v = -FLT_MAX; if (xok[kx]) { v = pr[kx]; } m = fmaxf(m, v); }
        }
        return m;
      };
      int y0 = oy0 * S - py;
      float hr[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { hr[k] = hrow(y0 + k); }
      int const obase = pl * ohw + ox;
      for (int oy = oy0; oy < oy1; ++oy) {
        float out_v = hr[0];
#pragma unroll
        for (int k = 1; k < K; ++k) { out_v = fmaxf(out_v, hr[k]); }
        op[obase + oy * OW] = out_v;
        if (stage) { stage[obase + oy * OW] = out_v; }  // staged for the transposed (NHWC) plane write below
        amax = fmaxf(amax, fabsf(out_v));
        y0 += S;
        if (oy + 1 < oy1) {
#pragma unroll
          for (int k = 0; k < K; ++k) { if (k + S < K) { hr[k] = hr[k + S]; } else { hr[k] = hrow(y0 + k); } }
        }
      }
    }
  } else {
    // Average pooling keeps the reference's tap order (the fp32 sum is order-dependent): one output per thread per step.
    // (plane, oy, ox) of output o = threadIdx.x + 256 k, advanced by fixed deltas with carries instead of two integer divisions per output;
    // every tap's bounds test is two unsigned compares
    int pl = threadIdx.x / ohw, p0 = threadIdx.x - pl * ohw;
    int oy = p0 / OW, ox = p0 - oy * OW;
    int const d_pl = 256 / ohw, d_p = 256 - d_pl * ohw, d_oy = d_p / OW, d_ox = d_p - d_oy * OW;
    for (int o = threadIdx.x; o < n_out; o += 256) {
      int const y0 = oy * S - py, x0 = ox * S - px;
      float const *ps = plane_s + pl * hw + y0 * W + x0;
      float out_v = 0.0f, cnt = 0.0f;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        bool const xok = static_cast<unsigned>(x0 + kx) < static_cast<unsigned>(W);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          bool const ok = xok && static_cast<unsigned>(y0 + ky) < static_cast<unsigned>(H);
          if (ok) { out_v += ps[ky * W + kx]; cnt += 1.0f; }
        }
      }
      out_v = __fdiv_rn(out_v, cnt);
      op[o] = out_v;
      if (stage) { stage[o] = out_v; }
      amax = fmaxf(amax, fabsf(out_v));
      ox += d_ox; oy += d_oy; pl += d_pl;
      if (ox >= OW) { ox -= OW; ++oy; }
      if (oy >= OH) { oy -= OH; ++pl; }
    }
  }
  return amax;
}

// NHWC 16-bit plane write of the staged outputs `os` ([np][ohw], np a multiple of 8 consecutive channels of ONE image starting at plane0).
__device__ __forceinline__ void pool_planes_write16(float const *os, PoolPlanes const &pp, long long plane0, int np, int ohw, float s) {
  {
    // Layout-transform elimination: also write the NHWC 16-bit plane(s) the consuming convolution reads (it then skips its pack kernel). The
    // CTA's planes are np (a multiple of 8) consecutive channels of ONE image, so every pixel gets 16-byte runs of 8 channels. The fp16 planes'
    // power-of-two scale comes from max|in| (published by the producer of the pooled node): pooling never increases max|x|.
    long long const img = plane0 / pp.C;
    int const chan0 = static_cast<int>(plane0 - img * pp.C), groups = np >> 3;
    for (int idx = threadIdx.x; idx < ohw * groups; idx += 256) {
      int const pix = idx / groups, g = idx - pix * groups;
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float const t0 = os[(8 * g + 2 * k) * ohw + pix] * s, t1 = os[(8 * g + 2 * k + 1) * ohw + pix] * s;
        if (pp.bf16) {
          __nv_bfloat162 const h = __floats2bfloat162_rn(t0, t1);
          wh[k] = *reinterpret_cast<uint32_t const *>(&h);
          wl[k] = 0u;
        } else {
          __half2 const h = __floats2half2_rn(t0, t1);
          float2 const hf = __half22float2(h);
          __half2 const l = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
          wh[k] = *reinterpret_cast<uint32_t const *>(&h);
          wl[k] = *reinterpret_cast<uint32_t const *>(&l);
        }
      }
      long long orow = img * ohw + pix;
      if (pp.dWp) { int const oy = pix / pp.OW, ox = pix - oy * pp.OW; orow = (img * pp.dHp + oy + pp.dpy) * pp.dWp + ox + pp.dpx; }
      long long const o16 = orow * pp.cpad + chan0 + 8 * g;
      *reinterpret_cast<uint4 *>(pp.hi + o16) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      if (pp.lo) { *reinterpret_cast<uint4 *>(pp.lo + o16) = make_uint4(wl[0], wl[1], wl[2], wl[3]); }
    }
  }
}

template <int K, int S>
__global__ void __launch_bounds__(256)
pool_plane_kernel(float const *__restrict__ in, float *__restrict__ out, int H, int W, int OH, int OW, int py, int px, int avg_pool,
                  unsigned int *out_absmax, int ppc, long long n_planes, PoolPlanes pp) {
  pdl_prologue();
  extern __shared__ __align__(16) float plane_s[];
  long long const plane0 = static_cast<long long>(blockIdx.x) * ppc;
  int const np = static_cast<int>(min(static_cast<long long>(ppc), n_planes - plane0));
  int const hw = H * W, n_in = np * hw, ohw = OH * OW, n_out = np * ohw;
  float const *ip = in + plane0 * hw;
  if ((reinterpret_cast<uintptr_t>(ip) & 15) == 0) {
    // all loads of a batch are issued before the first shared-memory store (8 x 16 B in flight per thread)
    float4 const *ip4 = reinterpret_cast<float4 const *>(ip);
    float4 *s4 = reinterpret_cast<float4 *>(plane_s);
    int const n4 = n_in >> 2;
    constexpr int kU = 8;
    for (int base = threadIdx.x; base < n4; base += 256 * kU) {
      float4 v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) { int const i = base + u * 256; if (i < n4) { v[u] = __ldg(ip4 + i); } }
#pragma unroll
      for (int u = 0; u < kU; ++u) { int const i = base + u * 256; if (i < n4) { s4[i] = v[u]; } }
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n_in; i += 256) { plane_s[i] = __ldg(ip + i); }
  } else {
    for (int i = threadIdx.x; i < n_in; i += 256) { plane_s[i] = __ldg(ip + i); }
  }
  __syncthreads();
  float amax = 0.0f;
  float *op = out + plane0 * ohw;
  bool const stage_out = pp.hi != nullptr;
  amax = pool_planes_compute<K, S>(plane_s, stage_out ? plane_s + n_in : nullptr, op, np, H, W, OH, OW, py, px, avg_pool);
  if (out_absmax) { publish_absmax_warp(amax, out_absmax); }
  if (pp.hi) {
    // Layout-transform elimination: also write the NHWC 16-bit plane(s) the consuming convolution reads (it then skips its pack kernel). The
    // CTA's planes are np (a multiple of 8) consecutive channels of ONE image, so every pixel gets 16-byte runs of 8 channels. The fp16 planes'
    // power-of-two scale comes from max|in| (published by the producer of the pooled node): pooling never increases max|x|.
    float const s = pp.in_absmax ? scale_from_absmax_bits(*pp.in_absmax) : 1.0f;
    if (blockIdx.x == 0 && threadIdx.x == 0) { pp.scale2[0] = s; pp.scale2[1] = 1.0f / s; }
    __syncthreads();
    pool_planes_write16(plane_s + n_in, pp, plane0, np, ohw, s);
  }
}

// Persistent, double-buffered form of the same kernel: each CTA walks plane groups g = blockIdx.x, blockIdx.x + gridDim.x, ... and one
// thread fetches group g+1 with a 1-D bulk copy (cp.async.bulk, completion on an mbarrier) while all threads pool group g, so the HBM read
// stream never stops for the compute phase (with one group per CTA, co-resident CTAs load together and then compute together: 21 % of DRAM
// peak on AlexNet pool1, ncu). Host guarantees: every group's byte count is a multiple of 16 and `in` is 16-byte aligned.
template <int K, int S>
__global__ void __launch_bounds__(256)
pool_plane_pipe_kernel(float const *__restrict__ in, float *__restrict__ out, int H, int W, int OH, int OW, int py, int px, int avg_pool,
                       unsigned int *out_absmax, int ppc, long long n_planes, PoolPlanes pp) {
  extern __shared__ __align__(128) float plane_s[];
  __shared__ __align__(8) uint64_t full_bar[2];
  int const hw = H * W, ohw = OH * OW;
  int const buf_floats = (ppc * hw + 31) & ~31;  // 128-byte aligned buffers
  float *stage = pp.hi ? plane_s + 2 * buf_floats : nullptr;
  long long const n_groups = (n_planes + ppc - 1) / ppc;
  if (threadIdx.x == 0) { mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1); fence_barrier_init(); }
  __syncthreads();
  pdl_prologue();
  auto fetch = [&](long long g, int b) {  // one thread
    long long const p0 = g * ppc;
    uint32_t const bytes = static_cast<uint32_t>(min(static_cast<long long>(ppc), n_planes - p0)) * static_cast<uint32_t>(hw) * 4u;
    mbar_expect_tx(&full_bar[b], bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(plane_s + b * buf_floats)),
                 "l"(in + p0 * hw), "r"(bytes), "r"(smem_u32(&full_bar[b]))
                 : "memory");
  };
  float const s = (pp.hi && pp.in_absmax) ? scale_from_absmax_bits(*pp.in_absmax) : 1.0f;
  if (pp.hi && blockIdx.x == 0 && threadIdx.x == 0) { pp.scale2[0] = s; pp.scale2[1] = 1.0f / s; }
  float amax = 0.0f;
  long long g = blockIdx.x;
  if (threadIdx.x == 0 && g < n_groups) { fetch(g, 0); }
  for (int it = 0; g < n_groups; g += gridDim.x, ++it) {
    int const b = it & 1;
    if (threadIdx.x == 0 && g + gridDim.x < n_groups) { fetch(g + gridDim.x, b ^ 1); }  // (buffer b^1 was released by the barrier below)
    mbar_wait(&full_bar[b], (it >> 1) & 1);
    long long const plane0 = g * ppc;
    int const np = static_cast<int>(min(static_cast<long long>(ppc), n_planes - plane0));
    amax = fmaxf(amax, pool_planes_compute<K, S>(plane_s + b * buf_floats, stage, out + plane0 * ohw, np, H, W, OH, OW, py, px, avg_pool));
    __syncthreads();  // everybody is done with buffer b; the staged outputs are complete
    if (pp.hi) {
      pool_planes_write16(stage, pp, plane0, np, ohw, s);
      __syncthreads();  // ... before the next group overwrites the stage
    }
  }
  if (out_absmax) { publish_absmax_warp(amax, out_absmax); }
}

// ---- lrn (test/rtc/lrn.cucl:35-50, LRN_MATCH_CAFFE branch) ----------------------------------------------------
// The reference runs one thread per (img,y,x) that walks ALL channels with a running add-new / subtract-old sum of squares.
// At B200 widths that is latency-bound (a 27x27 map has too few pixels to fill 148 SMs), so a thread here owns one pixel and
// one CHUNK of kChunk output channels: it rebuilds the window sum at the chunk start (ascending-order FMAs, exactly the
// reference's sequence for chunk 0) and then runs the reference's running update inside the chunk. Consecutive threads are
// consecutive x, so every channel step reads/writes coalesced 128-byte rows. kLS = local_size (ring buffer in registers).
// x^(-beta) for the LRN scale: exp2(-beta * log2(x)), the same two approximations and the same product as __powf (the reference compiles
// lrn.cucl with --use_fast_math, src/nvrtc_util.cc:251), in their flush-to-zero forms: __powf's non-ftz expansion wraps both in subnormal
// range fix-ups (8 instructions, 2 predicates) that a scale base >= k can never need. Identical bits for normal-range arguments.
__device__ __forceinline__ float lrn_pow_neg(float x, float neg_beta) {
  float l, e;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l * neg_beta));
  return e;
}

// One pixel, one chunk of kChunk output channels. kInterior: the chunk and its window halo lie inside [0, C) -- no channel bounds tests.
template <int kLS, int kChunk, bool kInterior>
__device__ __forceinline__ float lrn_chunk(float const *__restrict__ in, float *__restrict__ out, long long base, int c_begin, int C, int HW, float alpha_over_ls,
                                           float neg_beta, float k) {
  constexpr int hls = kLS >> 1;
  constexpr int kTot = kChunk + 2 * hls;  // input channels c_begin-hls .. c_begin+kChunk-1+hls
  float v[kTot];
  float const *ip = in + base + static_cast<long long>(c_begin - hls) * HW;
#pragma unroll
  for (int s = 0; s < kTot; ++s) {  // all loads first: kTot independent requests in flight per thread
    int const ic = c_begin - hls + s;
    v[s] = (kInterior || (ic >= 0 && ic < C)) ? __ldg(ip + static_cast<long long>(s) * HW) : 0.0f;
  }
  float *op = out + base + static_cast<long long>(c_begin) * HW;
  float ls_sum = 0.0f, amax = 0.0f;
#pragma unroll
  for (int s = 0; s < kTot; ++s) {  // the reference's running update: add the newest square, subtract the one leaving the window
    ls_sum = __fmaf_rn(v[s], v[s], ls_sum);
    if (s >= kLS) { ls_sum = __fmaf_rn(-v[s - kLS], v[s - kLS], ls_sum); }
    if (s >= 2 * hls) {
      int const j = s - 2 * hls;  // output channel c_begin + j
      if (kInterior || c_begin + j < C) {
        float const scale_base = __fmaf_rn(ls_sum, alpha_over_ls, k);
        float const ov = v[s - hls] * lrn_pow_neg(scale_base, neg_beta);
        op[static_cast<long long>(j) * HW] = ov;
        amax = fmaxf(amax, fabsf(ov));
      }
    }
  }
  return amax;
}

template <int kLS, int kChunk>
__global__ void __launch_bounds__(128)
lrn_kernel(float const *__restrict__ in, float *__restrict__ out, long long n_pels, int C, int HW, float alpha, float beta, float k,
           unsigned int *out_absmax) {
  pdl_prologue();
  long long const pel = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  constexpr int hls = kLS >> 1;
  int const c_begin = blockIdx.y * kChunk;
  float amax = 0.0f;
  if (pel < n_pels) {
    long long const img = pel / HW;
    long long const base = img * C * HW + (pel - img * HW);
    float const alpha_over_ls = alpha / (float)kLS;
    // (measured before this split: 67 SASS instructions per output, most of them channel-bounds predicates and __powf's subnormal fix-ups)
    if (c_begin - hls >= 0 && c_begin + kChunk + hls <= C) { amax = lrn_chunk<kLS, kChunk, true>(in, out, base, c_begin, C, HW, alpha_over_ls, -beta, k); }
    else { amax = lrn_chunk<kLS, kChunk, false>(in, out, base, c_begin, C, HW, alpha_over_ls, -beta, k); }
  }
  if (out_absmax) { publish_absmax_warp(amax, out_absmax); }
}

// generic local_size fallback (ring in local memory)
__global__ void lrn_kernel_generic(float const *__restrict__ in, float *__restrict__ out, long long n_pels, int C, int HW,
                                   int local_size, float alpha, float beta, float k, unsigned int *out_absmax) {
  pdl_prologue();
  long long const pel = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  bool const valid = pel < n_pels;
  long long const img = valid ? pel / HW : 0;
  long long const base = img * C * HW + (valid ? (pel - img * HW) : 0);
  float amax = 0.0f;
  if (!valid) { C = -local_size; }  // no iterations
  int const hls = local_size >> 1;
  float const alpha_over_ls = alpha / (float)local_size;
  float ls_buf[32];
  for (int i = 0; i < local_size; ++i) { ls_buf[i] = 0.0f; }
  float ls_sum = 0.0f;
  for (int ic = 0; ic < C + hls; ++ic) {
    int const lsb_ix = ic % local_size;
    float const ls_old = ls_buf[lsb_ix];
    ls_buf[lsb_ix] = (ic < C) ? __ldg(in + base + static_cast<long long>(ic) * HW) : 0.0f;
    ls_sum = __fmaf_rn(ls_buf[lsb_ix], ls_buf[lsb_ix], ls_sum);
    ls_sum = __fmaf_rn(-ls_old, ls_old, ls_sum);
    if (ic >= hls) {
      float const scale_base = __fmaf_rn(ls_sum, alpha_over_ls, k);
      float const ov = ls_buf[(lsb_ix + local_size - hls) % local_size] * powf(scale_base, -beta);
      out[base + static_cast<long long>(ic - hls) * HW] = ov;
      amax = fmaxf(amax, fabsf(ov));
    }
  }
  if (out_absmax) { publish_absmax_warp(amax, out_absmax); }
}

// ---- lrn + max-pool(3x3, stride 2) in one kernel (test/rtc/lrn.cucl:35-50 followed by test/rtc/pool.cucl:1-41) ---------------------------
// AlexNet-ng norm1 -> pool1 and norm2 -> pool2 (GoogLeNet: norm2 -> pool2): the LRN output is read by nothing but the pool, and both kernels
// are bandwidth kernels -- apart, the normalised map makes a round trip through memory (37 MB written, then read, for norm1) and costs a launch.
// A CTA owns kLpRows pooled rows x all columns x kLpCC channels of one image. Phase 1 is lrn_kernel's: one thread per pixel of the
// 2*kLpRows+1 input rows loads its kLpCC+4 channel values straight from memory (consecutive threads = consecutive x: coalesced rows; the
// window reaches two channels to either side), runs lrn_chunk<5, 16>'s arithmetic in registers -- so the values are bit-identical to the plain
// kernel's -- and leaves the normalised values in shared memory. Phase 2 pools them and writes the pooled node and (optionally) the consumer
// convolution's NHWC 16-bit planes, in the consumer's layout. The LRN node itself is not written (the whole-net driver recomputes it on demand
// with the plain lrn function when somebody asks for it).
// (Two earlier forms staged the inputs in shared memory -- per-element copies, then one bulk copy per channel into a persistent double-buffered
// ring: 87 KB per CTA left two CTAs = 16 warps per SM for code that is a chain of dependent FMAs and exp2/log2, 1.3 TB/s on 37 MB and SLOWER
// than the two separate kernels on GoogLeNet's 154 MB map, r02 ncu. This form needs 39 KB; small CTAs (128 threads, four per SM) measured best.
// It is still bound by its instruction stream (LRN's FMA / exp2 / log2 chain plus the pooling loads), not by memory: on AlexNet's 37 / 24 MB maps
// it saves the launch and the round trip (42 -> 31 us, 35 -> 25 us); on a 154 MB map it only equals the two kernels, so the whole-net driver
// fuses maps up to 64 MB.)
constexpr int kLpCC = 16, kLpRows = 4, kLpHalo = 2;  // channels per CTA (= the chunk of lrn_kernel<5, 16>), pooled rows per CTA, LRN half window (local_size 5)

template <int kT, int kB>
__global__ void __launch_bounds__(kT, kB)
lrn_maxpool_kernel(float const *__restrict__ in, float *__restrict__ out, int C, int H, int W, int OH, int OW, float alpha_over_ls, float neg_beta, float k,
                   unsigned int *out_absmax, PoolPlanes pp) {
  extern __shared__ __align__(16) float lp_s[];
  constexpr int kCh = kLpCC + 2 * kLpHalo;
  int const oy0 = blockIdx.x * kLpRows, c0 = blockIdx.y * kLpCC;
  long long const img = blockIdx.z;
  int const n_orows = min(kLpRows, OH - oy0);
  int const iy0 = oy0 * 2, n_irows = min(H - iy0, 2 * n_orows + 1);
  int const npix = n_irows * W;        // rows iy0 .. iy0 + n_irows - 1 are one contiguous run of a channel plane
  int const nstride = (2 * kLpRows + 1) * W;  // floats per normalised channel in shared memory
  float *stage = lp_s + kLpCC * nstride;      // pooled outputs [kLpCC][n_orows * OW] for the plane write
  long long const HW = static_cast<long long>(H) * W;
  constexpr int nthreads = kT;
  pdl_prologue();
  // ---- phase 1: normalise. lrn_chunk<5, 16>'s arithmetic: the window sum rebuilt at the chunk start, then the running update ----
  // (a thread's next pixel is loaded while the current one is normalised: the loads are the long pole and a 128-thread CTA walks ~4 pixels per thread)
  float const *src0 = in + ((img * C + (c0 - kLpHalo)) * H + iy0) * static_cast<long long>(W);  // only dereferenced for channels inside [0, C)
  float nxt[kCh];
#pragma unroll
  for (int t = 0; t < kCh; ++t) {
    nxt[t] = (static_cast<int>(threadIdx.x) < npix && static_cast<unsigned>(c0 - kLpHalo + t) < static_cast<unsigned>(C)) ? __ldg(src0 + threadIdx.x + t * HW) : 0.0f;
  }
  for (int p = threadIdx.x; p < npix; p += nthreads) {
    float v[kCh];
#pragma unroll
    for (int t = 0; t < kCh; ++t) { v[t] = nxt[t]; }
    if (p + nthreads < npix) {
#pragma unroll
      for (int t = 0; t < kCh; ++t) { nxt[t] = (static_cast<unsigned>(c0 - kLpHalo + t) < static_cast<unsigned>(C)) ? __ldg(src0 + p + nthreads + t * HW) : 0.0f; }
    }
    float ls_sum = 0.0f;
#pragma unroll
    for (int t = 0; t < kCh; ++t) {
      ls_sum = __fmaf_rn(v[t], v[t], ls_sum);
      if (t >= 5) { ls_sum = __fmaf_rn(-v[t - 5], v[t - 5], ls_sum); }
      if (t >= 2 * kLpHalo) {
        float const scale_base = __fmaf_rn(ls_sum, alpha_over_ls, k);
        lp_s[(t - 2 * kLpHalo) * nstride + p] = v[t - kLpHalo] * lrn_pow_neg(scale_base, neg_beta);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: 3x3 / 2 max pooling (windows clipped at the map's edge, Caffe ceil rule) of channels c0 .. c0 + kLpCC - 1. A thread keeps ONE
  // output pixel and walks the channels (the pixel's index arithmetic and edge tests happen once); nthreads / n_out thread groups share the channels ----
  float amax = 0.0f;
  int const n_out = n_orows * OW;
  int const n_ch = min(kLpCC, C - c0);
  {
    int const n_grp = n_out <= nthreads ? nthreads / n_out : 1;
    int const grp = n_out <= nthreads ? static_cast<int>(threadIdx.x) / n_out : 0;
    long long const ohw = static_cast<long long>(OH) * OW;
    for (int rem = static_cast<int>(threadIdx.x) - grp * n_out; rem < n_out && grp < n_grp; rem += nthreads) {
      int const oyl = (rem >= OW) + (rem >= 2 * OW) + (rem >= 3 * OW), ox = rem - oyl * OW;  // (kLpRows == 4)
      int const ny = min(3, n_irows - oyl * 2), nx = min(3, W - ox * 2);
      float *op = out + ((img * C + c0) * OH + oy0) * static_cast<long long>(OW) + rem + grp * ohw;
      float const *pl = lp_s + grp * nstride + (oyl * 2) * W + ox * 2;
      float *sp = stage + grp * n_out + rem;
      if (ny == 3 && nx == 3) {
        float const *r1 = pl + W, *r2 = r1 + W;
#pragma unroll 2
        for (int ch = grp; ch < n_ch; ch += n_grp, pl += n_grp * nstride, r1 += n_grp * nstride, r2 += n_grp * nstride, op += n_grp * ohw, sp += n_grp * n_out) {
          float const m = fmaxf(fmaxf(fmaxf(pl[0], pl[1]), fmaxf(pl[2], r1[0])), fmaxf(fmaxf(r1[1], r1[2]), fmaxf(r2[0], fmaxf(r2[1], r2[2]))));
          *op = m;
          if (pp.hi) { *sp = m; }
          amax = fmaxf(amax, fabsf(m));
        }
      } else {
        for (int ch = grp; ch < n_ch; ch += n_grp, pl += n_grp * nstride, op += n_grp * ohw, sp += n_grp * n_out) {
          float m = -FLT_MAX;
          for (int y = 0; y < ny; ++y) { for (int x = 0; x < nx; ++x) { m = fmaxf(m, pl[y * W + x]); } }
          *op = m;
          if (pp.hi) { *sp = m; }
          amax = fmaxf(amax, fabsf(m));
        }
      }
    }
  }
  if (out_absmax) { publish_absmax_warp(amax, out_absmax); }
  if (pp.hi) {  // the consumer's planes: 16-byte runs of 8 channels per pixel (host: C a multiple of 8; fp16 planes: the scale from max|in|, the LRN scale factor is <= 1 for k >= 1)
    float const sc = pp.in_absmax ? scale_from_absmax_bits(*pp.in_absmax) : 1.0f;
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) { pp.scale2[0] = sc; pp.scale2[1] = 1.0f / sc; }
    __syncthreads();
    int const groups = n_ch >> 3;
    static_assert(kLpCC == 16 && kLpRows == 4, "index arithmetic below");
    for (int idx = threadIdx.x; idx < n_out * 2; idx += nthreads) {
      int const pix = idx >> 1, g = idx & 1;
      if (g >= groups) { continue; }
      uint32_t wh[4], wl[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float const t0 = stage[(8 * g + 2 * q) * n_out + pix] * sc, t1 = stage[(8 * g + 2 * q + 1) * n_out + pix] * sc;
        if (pp.bf16) {
          __nv_bfloat162 const h = __floats2bfloat162_rn(t0, t1);
          wh[q] = *reinterpret_cast<uint32_t const *>(&h);
          wl[q] = 0u;
        } else {
          __half2 const h = __floats2half2_rn(t0, t1);
          float2 const hf = __half22float2(h);
          __half2 const l = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
          wh[q] = *reinterpret_cast<uint32_t const *>(&h);
          wl[q] = *reinterpret_cast<uint32_t const *>(&l);
        }
      }
      int const oyl = (pix >= OW) + (pix >= 2 * OW) + (pix >= 3 * OW), oy = oy0 + oyl, ox = pix - oyl * OW;
      long long const orow = pp.dWp ? (img * pp.dHp + oy + pp.dpy) * pp.dWp + ox + pp.dpx : (img * OH + oy) * static_cast<long long>(OW) + ox;
      long long const o16 = orow * pp.cpad + c0 + 8 * g;
      *reinterpret_cast<uint4 *>(pp.hi + o16) = make_uint4(wh[0], wh[1], wh[2], wh[3]);
      if (pp.lo) { *reinterpret_cast<uint4 *>(pp.lo + o16) = make_uint4(wl[0], wl[1], wl[2], wl[3]); }
    }
  }
}

// ---- softmax over chan (test/rtc/softmax.cucl:6-21; running max starts at 0.0f) -------------------------------
// One warp per (img,y,x): lanes stride the channel dim, shuffle-reduce max and sum.
__global__ void softmax_kernel(float const *__restrict__ in, float *__restrict__ prob, long long n_pels, int C, int HW) {
  pdl_prologue();
  long long const pel = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  int const lane = threadIdx.x & 31;
  if (pel >= n_pels) { return; }
  long long const img = pel / HW;
  long long const base = img * C * HW + (pel - img * HW);
  float pel_max = 0.0f;
  if (C <= 1024) {
    // up to 32 channels per lane stay in registers: one pass over memory with all loads in flight (the three-pass form below is a chain of
    // ~100 dependent global round trips per pixel: 27 us for the 32 x 1000 logits of ResNet-50, ncu). Same per-lane order of the exp sum and
    // the same shuffle trees as below, so the result is bit-identical.
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { int const c = lane + 32 * k; v[k] = (c < C) ? __ldg(in + base + static_cast<long long>(c) * HW) : 0.0f; }
#pragma unroll
    for (int k = 0; k < 32; ++k) { if (lane + 32 * k < C) { pel_max = fmaxf(pel_max, v[k]); } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { pel_max = fmaxf(pel_max, __shfl_xor_sync(0xffffffffu, pel_max, o)); }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 32; ++k) { if (lane + 32 * k < C) { v[k] = expf(v[k] - pel_max); sum += v[k]; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); }
#pragma unroll
    for (int k = 0; k < 32; ++k) { int const c = lane + 32 * k; if (c < C) { prob[base + static_cast<long long>(c) * HW] = __fdiv_rn(v[k], sum); } }
    return;
  }
  for (int c = lane; c < C; c += 32) { pel_max = fmaxf(pel_max, __ldg(in + base + static_cast<long long>(c) * HW)); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { pel_max = fmaxf(pel_max, __shfl_xor_sync(0xffffffffu, pel_max, o)); }
  float pel_sum = 0.0f;
  for (int c = lane; c < C; c += 32) {
    float const v = expf(__ldg(in + base + static_cast<long long>(c) * HW) - pel_max);
    prob[base + static_cast<long long>(c) * HW] = v;
    pel_sum += v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { pel_sum += __shfl_xor_sync(0xffffffffu, pel_sum, o); }
  for (int c = lane; c < C; c += 32) { prob[base + static_cast<long long>(c) * HW] = __fdiv_rn(prob[base + static_cast<long long>(c) * HW], pel_sum); }
}

// ---- operand pack: abs-max -> power-of-two scale -> transpose + 16-bit split ----------------------------------
// absmax over a tensor (non-negative floats order like their bit patterns, so atomicMax on uint works) + the power-of-two operand scale
// derived from it: cells = {max bits, blocks-done counter}; the LAST block to finish turns the max into {scale, 1/scale} and re-arms
// the cells, so no separate one-thread "finalize" launch is needed. Four independent 128-bit loads in flight per thread.
__global__ void __launch_bounds__(256)
absmax_kernel(float const *__restrict__ x, long long n, unsigned int *__restrict__ cells, float *__restrict__ scale2) {
  pdl_prologue();
  float m = 0.0f;
  long long const stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long const tid = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    long long const n4 = n >> 2;
    float4 const *x4 = reinterpret_cast<float4 const *>(x);
    for (long long i0 = tid; i0 < n4; i0 += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { long long const i = i0 + u * stride; v[u] = (i < n4) ? __ldg(x4 + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < 4; ++u) { m = fmaxf(m, fmaxf(fmaxf(fabsf(v[u].x), fabsf(v[u].y)), fmaxf(fabsf(v[u].z), fabsf(v[u].w)))); }
    }
    for (long long i = (n4 << 2) + tid; i < n; i += stride) { m = fmaxf(m, fabsf(__ldg(x + i))); }
  } else {
    for (long long i = tid; i < n; i += stride) { m = fmaxf(m, fabsf(__ldg(x + i))); }
  }
  __shared__ float warp_m[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
  if ((threadIdx.x & 31) == 0) { warp_m[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = 0.0f;
    for (int w = 0; w < 8; ++w) { bm = fmaxf(bm, warp_m[w]); }
    if (bm > 0.0f) { atomicMax(cells, __float_as_uint(bm)); }
    __threadfence();
    unsigned int const ticket = atomicAdd(cells + 1, 1u);
    if (ticket == gridDim.x - 1) {  // every block's max is in: publish the scale, re-arm the cells for the next use
      __threadfence();
      float const s = scale_from_absmax_bits(*reinterpret_cast<volatile unsigned int *>(cells));
      scale2[0] = s;
      scale2[1] = 1.0f / s;
      cells[0] = 0u;
      cells[1] = 0u;
    }
  }
}
// HW == 1 activations (fc7 / fc8 inputs): NCHW [img][chan][1][1] already IS the K-major matrix [img][chan]; only scale + split.
template <bool kBf16>
__global__ void pack_rows_split_kernel(float const *__restrict__ src, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, float const *__restrict__ scale2,
                                       int R, long long dst_b_stride, long long n, unsigned int const *__restrict__ absmax_bits, long long kmajor_rows) {
  pdl_prologue();
  float const s = absmax_bits ? scale_from_absmax_bits(*absmax_bits) : scale2[0];
  if (absmax_bits && threadIdx.x == 0 && blockIdx.x == 0) { const_cast<float *>(scale2)[0] = s; const_cast<float *>(scale2)[1] = 1.0f / s; }
  long long const i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) { return; }
  long long const b = i / R;
  int const r = static_cast<int>(i - b * R);
  float const v = __ldg(src + i) * s;
  long long const o = kmajor_rows ? ((static_cast<long long>(r) >> 6) * kmajor_rows + b) * 64 + (r & 63) : b * dst_b_stride + r;  // see pack_xpose_split_kernel
  if (kBf16) {
    __nv_bfloat16 const h = __float2bfloat16_rn(v);
    reinterpret_cast<__nv_bfloat16 *>(hi)[o] = h;
    if (lo) { reinterpret_cast<__nv_bfloat16 *>(lo)[o] = __float2bfloat16_rn(v - __bfloat162float(h)); }
  } else {
    __half const h = __float2half_rn(v);
    reinterpret_cast<__half *>(hi)[o] = h;
    if (lo) { reinterpret_cast<__half *>(lo)[o] = __float2half_rn(v - __half2float(h)); }
  }
}

// Few-channel activations for the row-merged conv path (network inputs: chan = 3): one thread per pixel gathers its kCp-padded channel
// vector (each channel read is coalesced across the warp) and writes it as ONE 8- or 16-byte word per plane:
// NCHW -> [img][y][x_pitch][kCp], pixel (y,x) at column x + px_off.
template <int kCp, bool kBf16>
__global__ void __launch_bounds__(256)
pack_smallc_kernel(float const *__restrict__ src, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, float const *__restrict__ scale2, int C, int H, int W,
                   int Wp, int px_off, long long n_pix, unsigned int const *__restrict__ absmax_bits) {
  pdl_prologue();
  float const s = absmax_bits ? scale_from_absmax_bits(*absmax_bits) : scale2[0];
  if (absmax_bits && threadIdx.x == 0 && blockIdx.x == 0) { const_cast<float *>(scale2)[0] = s; const_cast<float *>(scale2)[1] = 1.0f / s; }
  long long const i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_pix) { return; }
  long long const hw = static_cast<long long>(H) * W;
  long long const img = i / hw;
  int const pix = static_cast<int>(i - img * hw);
  int const y = pix / W, x = pix - y * W;
  float const *sp = src + img * C * hw + pix;
  uint16_t h[kCp], l[kCp];
#pragma unroll
  for (int c = 0; c < kCp; ++c) {
    float const v = (c < C) ? __ldg(sp + c * hw) * s : 0.0f;
    if (kBf16) {
      __nv_bfloat16 const hv = __float2bfloat16_rn(v);
      h[c] = __bfloat16_as_ushort(hv);
      l[c] = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(hv)));
    } else {
      __half const hv = __float2half_rn(v);
      h[c] = __half_as_ushort(hv);
      l[c] = __half_as_ushort(__float2half_rn(v - __half2float(hv)));
    }
  }
  long long const o = ((img * H + y) * Wp + x + px_off) * kCp;
  if (kCp == 4) {
    *reinterpret_cast<uint2 *>(hi + o) = *reinterpret_cast<uint2 *>(h);
    if (lo) { *reinterpret_cast<uint2 *>(lo + o) = *reinterpret_cast<uint2 *>(l); }
  } else {
    *reinterpret_cast<uint4 *>(hi + o) = *reinterpret_cast<uint4 *>(h);
    if (lo) { *reinterpret_cast<uint4 *>(lo + o) = *reinterpret_cast<uint4 *>(l); }
  }
}

// absmax_kernel + pack_smallc_kernel in one launch (fp16 planes need max|x| of the WHOLE tensor before the first element can be written):
// a CTA owns `rows` image rows of one image, all C channels; it fetches them into shared memory with one bulk copy per channel (runs start at
// any float offset of a 16-byte unit: the enclosing units are fetched), folds its max|x| into cells[0], waits at a grid-wide barrier
// (cells[1]; every CTA is resident -- the host sizes the grid by occupancy), derives the scale and writes its pixels' planes from shared
// memory. The input is read from memory once; cells[2] counts the CTAs out, the last one re-arms the cells.
__host__ __device__ constexpr int smallc_row_stride(int rows, int W) { return (rows * W + 3 + 3) & ~3; }

template <int kCp>
__global__ void __launch_bounds__(256)
absmax_pack_smallc_kernel(float const *__restrict__ src, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, float *__restrict__ scale2, unsigned int *cells,
                          int C, int H, int W, int Wp, int px_off, int rows, int upi, long long n_total) {
  extern __shared__ __align__(128) float ap_s[];
  __shared__ __align__(8) uint64_t full_bar;
  __shared__ float warp_m[8];
  __shared__ float s_bcast;
  int const img = blockIdx.x / upi, y0 = (blockIdx.x - img * upi) * rows;
  int const nrows = min(rows, H - y0), npix = nrows * W;
  int const rs = smallc_row_stride(rows, W);
  long long const hw = static_cast<long long>(H) * W;
  long long const g0 = (static_cast<long long>(img) * C * H + y0) * W;  // float index of channel 0's run; channel c: + c * hw
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); fence_barrier_init(); }
  __syncthreads();
  pdl_prologue();
  if (threadIdx.x < 32) {  // lane c fetches channel c
    int const c = threadIdx.x;
    bool const mine = c < C;
    long long const g = g0 + c * hw;
    int const off = static_cast<int>(g & 3);
    // whole 16-byte units, but never past the end of the tensor (its last floats, if any, are read one by one below)
    long long const units_end = min((g - off) + ((off + npix + 3) & ~3), n_total & ~3ll);
    uint32_t const bytes = mine ? static_cast<uint32_t>(max(0ll, units_end - (g - off))) * 4u : 0u;
    uint32_t const total = __reduce_add_sync(0xffffffffu, bytes);
    if (c == 0) { mbar_expect_tx(&full_bar, total); }
    __syncwarp();
    if (mine && bytes) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ap_s + c * rs)), "l"(src + (g - off)),
                   "r"(bytes), "r"(smem_u32(&full_bar))
                   : "memory");
    }
    if (mine) { for (long long i = max(units_end, g); i < g + npix; ++i) { ap_s[c * rs + off + static_cast<int>(i - g)] = __ldg(src + i); } }
  }
  mbar_wait(&full_bar, 0);
  __syncthreads();
  float m = 0.0f;
  for (int c = 0; c < C; ++c) {
    float const *row = ap_s + c * rs + static_cast<int>((g0 + c * hw) & 3);
    for (int i = threadIdx.x; i < npix; i += 256) { m = fmaxf(m, fabsf(row[i])); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o)); }
  if ((threadIdx.x & 31) == 0) { warp_m[threadIdx.x >> 5] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bm = 0.0f;
    for (int w = 0; w < 8; ++w) { bm = fmaxf(bm, warp_m[w]); }
    if (bm > 0.0f) { atomicMax(cells, __float_as_uint(bm)); }
    __threadfence();
    atomicAdd(cells + 1, 1u);
    unsigned int v = 0, spins = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cells + 1) : "memory");
      if (v >= gridDim.x) { break; }
      if (++spins > (1u << 22)) { printf("b200: input pack barrier never completed (block %d: %u of %u)\n", blockIdx.x, v, gridDim.x); __trap(); }
    }
    float const s = scale_from_absmax_bits(*reinterpret_cast<volatile unsigned int *>(cells));
    s_bcast = s;
    if (blockIdx.x == 0) { scale2[0] = s; scale2[1] = 1.0f / s; }
    __threadfence();
    if (atomicAdd(cells + 2, 1u) == gridDim.x - 1) { cells[0] = 0u; cells[1] = 0u; cells[2] = 0u; }  // everybody has read the maximum: re-arm for the next launch
  }
  __syncthreads();
  float const s = s_bcast;
  int offc[kCp];
#pragma unroll
  for (int c = 0; c < kCp; ++c) { offc[c] = c * rs + static_cast<int>((g0 + c * hw) & 3); }
  for (int pix = threadIdx.x; pix < npix; pix += 256) {
    int const y = pix / W, x = pix - y * W;
    __align__(16) uint16_t h[kCp], l[kCp];
#pragma unroll
    for (int c = 0; c < kCp; ++c) {
      float const v = (c < C) ? ap_s[offc[c] + pix] * s : 0.0f;
      __half const hv = __float2half_rn(v);
      h[c] = __half_as_ushort(hv);
      l[c] = __half_as_ushort(__float2half_rn(v - __half2float(hv)));
    }
    long long const o = ((static_cast<long long>(img) * H + y0 + y) * Wp + x + px_off) * kCp;
    if (kCp == 4) {
      *reinterpret_cast<uint2 *>(hi + o) = *reinterpret_cast<uint2 *>(h);
      if (lo) { *reinterpret_cast<uint2 *>(lo + o) = *reinterpret_cast<uint2 *>(l); }
    } else {
      *reinterpret_cast<uint4 *>(hi + o) = *reinterpret_cast<uint4 *>(h);
      if (lo) { *reinterpret_cast<uint4 *>(lo + o) = *reinterpret_cast<uint4 *>(l); }
    }
  }
}

// src [B][R][C] fp32 (C contiguous)  ->  dst planes [B][C][Rpad] 16-bit (R contiguous, zero padded to Rpad):
//   NCHW activations  (B=img, R=chan, C=y*x)        -> NHWC  [img][y*x][chan_pad]
//   OIHW filters      (B=out_chan, R=in_chan, C=ky*kx) -> [out_chan][ky*kx][chan_pad]  (K-major rows for the Q operand)
//   sgemm a  K:M      (B=1, R=K, C=M)               -> [M][Kpad]
//   row-merged small-chan convs: NCHW -> [img][y][x_pitch][4|8], OIHW -> [out_chan][ky][(kx,chan) padded to 64]
// hi = cvt(s*x), lo = cvt(s*x - hi) (lo plane optional). kBf16 selects bf16 instead of fp16 storage.
// Tile: 64 (R) x 32 (C); loads are 128-byte rows along C, stores are 128-byte rows of half2 along R.
template <bool kBf16>
__global__ void __launch_bounds__(256)
pack_xpose_split_kernel(float const *__restrict__ src, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, float const *__restrict__ scale2,
                        int R, int C, int Rpad, long long dst_c_stride, long long dst_b_stride, int c_inner, long long dst_chi_stride, long long dst_base,
                        unsigned int const *__restrict__ absmax_bits, long long kmajor_rows, int tap_minor) {
  pdl_prologue();
  __shared__ float tile[64][33];
  int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  int const r0 = blockIdx.x * 64, c0 = blockIdx.y * 32;
  long long const b = blockIdx.z;
  // scale: either finalised earlier into scale2, or derived here from the abs-max cell the tensor's producer published
  // (then one thread also writes {scale, 1/scale} for the contraction kernel's epilogue)
  float const s = absmax_bits ? scale_from_absmax_bits(*absmax_bits) : scale2[0];
  if (absmax_bits && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { const_cast<float *>(scale2)[0] = s; const_cast<float *>(scale2)[1] = 1.0f / s; }
  float const *sp = src + b * static_cast<long long>(R) * C;
#pragma unroll
  for (int rr = ty; rr < 64; rr += 8) {
    int const r = r0 + rr, c = c0 + tx;
    tile[rr][tx] = (r < R && c < C) ? __ldg(sp + static_cast<long long>(r) * C + c) * s : 0.0f;
  }
  __syncthreads();
  int const r = r0 + 2 * tx;
  if (r < Rpad) {
#pragma unroll
    for (int cc = ty; cc < 32; cc += 8) {
      int const c = c0 + cc;
      if (c < C) {
        float const v0 = tile[2 * tx][cc], v1 = tile[2 * tx + 1][cc];
        // destination of source column c: (c / c_inner) * dst_chi_stride + (c % c_inner) * dst_c_stride  (c_inner >= C: plain c * dst_c_stride);
        // the two-level form lays image rows out with a padded pitch / filter rows as (ky)(kx,chan) for the row-merged conv path
        // k = position inside the destination row; plain layout: row-major [b][k]; k-block-major (kmajor_rows > 0, used for filters):
        // [k / 64][b (padded to kmajor_rows)][k % 64], so that the 128-row x 64-element tile one TMA load fetches is 16 KB CONTIGUOUS
        // tap_minor > 0 (filters of a halo-mode convolution, igemm4.cuh; = number of taps): k-blocks ordered channel block major, tap minor --
        // k = ((chan / 64) * taps + tap) * 64 + chan % 64 -- so that the ku consecutive k-blocks of a stage are ONE 3-d TMA box
        long long const k = tap_minor ? (static_cast<long long>(r >> 6) * tap_minor + c) * 64 + (r & 63)
                                      : dst_base + static_cast<long long>(c / c_inner) * dst_chi_stride + static_cast<long long>(c % c_inner) * dst_c_stride + r;
        long long const o = kmajor_rows ? ((k >> 6) * kmajor_rows + b) * 64 + (k & 63) : b * dst_b_stride + k;
        if (kBf16) {
          __nv_bfloat16 const h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          *reinterpret_cast<__nv_bfloat162 *>(hi + o) = __nv_bfloat162(h0, h1);
          if (lo) {
            *reinterpret_cast<__nv_bfloat162 *>(lo + o) =
                __nv_bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1)));
          }
        } else {
          __half const h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          *reinterpret_cast<__half2 *>(hi + o) = __half2(h0, h1);
          if (lo) {
            *reinterpret_cast<__half2 *>(lo + o) = __half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
          }
        }
      }
    }
  }
}

// Wide-tile variant for the big activation tensors (ResNet / GoogLeNet: 56x56, 28x28, 14x14 maps, pixel counts divisible by 4): a CTA moves a
// 64-channel x 128-pixel tile, every thread has eight 128-bit loads in flight (512-byte runs per channel row instead of 128), the
// transposed read of shared memory is conflict-free at a row pitch of 129 floats, stores are 128-byte rows of half2 as before. Plain NHWC
// destination only: dst[b][pixel][chan] with pitch dst_c_stride.
template <bool kBf16>
__global__ void __launch_bounds__(256)
pack_xpose_split_v4_kernel(float const *__restrict__ src, uint16_t *__restrict__ hi, uint16_t *__restrict__ lo, float const *__restrict__ scale2,
                           int R, int C, int Rpad, long long dst_c_stride, long long dst_b_stride, unsigned int const *__restrict__ absmax_bits) {
  pdl_prologue();
  constexpr int kPitch = 129;
  extern __shared__ float tile_v4[];  // [64][kPitch]
  int const tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  int const r0 = blockIdx.x * 64, c0 = blockIdx.y * 128;
  long long const b = blockIdx.z;
  float const s = absmax_bits ? scale_from_absmax_bits(*absmax_bits) : scale2[0];
  if (absmax_bits && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) { const_cast<float *>(scale2)[0] = s; const_cast<float *>(scale2)[1] = 1.0f / s; }
  float const *sp = src + b * static_cast<long long>(R) * C;
  int const c = c0 + 4 * tx;
  float4 v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {  // all loads first
    int const r = r0 + ty + 8 * k;
    v[k] = (r < R && c < C) ? __ldg(reinterpret_cast<float4 const *>(sp + static_cast<long long>(r) * C + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float *t = tile_v4 + (ty + 8 * k) * kPitch + 4 * tx;
    t[0] = v[k].x * s; t[1] = v[k].y * s; t[2] = v[k].z * s; t[3] = v[k].w * s;
  }
  __syncthreads();
  int const r = r0 + 2 * tx;
  if (r < Rpad) {
#pragma unroll 4
    for (int cc = ty; cc < 128; cc += 8) {
      int const cpix = c0 + cc;
      if (cpix < C) {
        float const v0 = tile_v4[(2 * tx) * kPitch + cc], v1 = tile_v4[(2 * tx + 1) * kPitch + cc];
        long long const o = b * dst_b_stride + static_cast<long long>(cpix) * dst_c_stride + r;
        if (kBf16) {
          __nv_bfloat16 const h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
          *reinterpret_cast<__nv_bfloat162 *>(hi + o) = __nv_bfloat162(h0, h1);
          if (lo) { *reinterpret_cast<__nv_bfloat162 *>(lo + o) = __nv_bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)), __float2bfloat16_rn(v1 - __bfloat162float(h1))); }
        } else {
          __half const h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
          *reinterpret_cast<__half2 *>(hi + o) = __half2(h0, h1);
          if (lo) { *reinterpret_cast<__half2 *>(lo + o) = __half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1))); }
        }
      }
    }
  }
}

}  // namespace b200
