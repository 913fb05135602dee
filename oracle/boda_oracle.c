/* boda_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker; never the product path).
 *
 * A plain-C CPU restatement of the operator semantics of Boda's rtc_fwd hot path, written from the
 * behaviour of the reference's CUCL templates. Every function cites the reference file:line it follows
 * (paths relative to the reference root). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED against the reference's own golden digests
 * (test/good_tr/{sgemm-gen5,sgemm-gen600,conv-gen5,conv-debug,conv-full-gen5,ops-prof-conv-3x3-cudnn-boda}/wisdom.wis,
 * decoded into tests/golden/wisdom_digests.json by tests/golden/make_golden.py) -- see tests/test_oracle_golden.py.
 *
 * All tensors are dense row-major fp32 in the reference ("ref") layouts: activations img:chan:y:x,
 * filters out_chan:in_chan:y:x, sgemm a K:M, b K:N, c M:N.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ---- deterministic input generation -------------------------------------------------------------- */

/* test/rtc/gen-util.h:1-9 : murmur3 finaliser of the flat index, mapped to [-5,5]. */
static inline float det_hash_rand(uint32_t rv) {
  uint32_t h = rv;
  h ^= h >> 16;
  h *= 0x85ebca6bu;
  h ^= h >> 13;
  h *= 0xc2b2ae35u;
  h ^= h >> 16;
  /* the reference builds with NVRTC's default fmad=true (src/nvrtc_util.cc:251), i.e. one fused multiply-add */
  return fmaf((float)(h), (10.0f / (float)(UINT32_MAX)), -5.0f);
}

ORACLE_API float oracle_det_hash_rand(uint32_t rv) { return det_hash_rand(rv); }

/* Salts: test/rtc/gen_data_Convolution_in.cucl:16 (234234567), gen_data_Convolution_filts.cucl:16 (8753985),
 * gen_data_Convolution_biases.cucl:11 (39475612), gen_data_sgemm_a.cucl:17 and gen_data_sgemm_b.cucl:15 (12738732). */
enum { SALT_CONV_IN = 234234567, SALT_CONV_FILTS = 8753985, SALT_CONV_BIASES = 39475612, SALT_SGEMM = 12738732 };

/* Generic 4-d generator for Convolution in / filts (test/rtc/gen_data_Convolution_in.cucl:9-19,
 * gen_data_Convolution_filts.cucl:9-19): mode 2 adds x, mode 3 adds y, mode 4 is a centre impulse, mode 5 the hash.
 * Note the reference's `if(mode==2) ...; if(mode==3) ... else if(mode==4) ... else if(mode==5)` chain is preserved. */
static void gen_4d(float *dst, uint32_t d0, uint32_t d1, uint32_t ysz, uint32_t xsz, uint32_t mode, float vi,
                   uint32_t salt) {
  uint64_t const n = (uint64_t)d0 * d1 * ysz * xsz;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t const x = (uint32_t)(i % xsz);
    uint32_t const y = (uint32_t)((i / xsz) % ysz);
    float val = vi;
    if (mode == 2) { val += (float)x; }
    if (mode == 3) { val += (float)y; }
    else if (mode == 4) { if ((x == xsz / 2) && (y == ysz / 2)) { val += 1.0f; } }
    else if (mode == 5) { val += det_hash_rand((uint32_t)i + salt); }
    dst[i] = val;
  }
}

ORACLE_API void oracle_gen_conv_in(float *in, uint32_t img, uint32_t chan, uint32_t y, uint32_t x, uint32_t mode,
                                   float vi) {
  gen_4d(in, img, chan, y, x, mode, vi, SALT_CONV_IN);
}
ORACLE_API void oracle_gen_conv_filts(float *filts, uint32_t out_chan, uint32_t in_chan, uint32_t y, uint32_t x,
                                      uint32_t mode, float vi) {
  gen_4d(filts, out_chan, in_chan, y, x, mode, vi, SALT_CONV_FILTS);
}
/* test/rtc/gen_data_Convolution_biases.cucl:7-13 : only mode 5 adds anything. */
ORACLE_API void oracle_gen_conv_biases(float *biases, uint32_t out_chan, uint32_t mode, float vi) {
  for (uint32_t i = 0; i < out_chan; ++i) {
    float val = vi;
    if (mode == 5) { val += det_hash_rand(i + SALT_CONV_BIASES); }
    biases[i] = val;
  }
}
/* test/rtc/gen_data_sgemm_a.cucl:7-21 : a is K:M (row = k). mode>=100 -> mode/100; mode 6 = M*1000+K. */
ORACLE_API void oracle_gen_sgemm_a(float *a, uint32_t K, uint32_t M, uint32_t mode, float vi) {
  uint32_t fin_mode = mode;
  if (fin_mode >= 100) { fin_mode = fin_mode / 100; }
  uint64_t const n = (uint64_t)K * M;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t const m = (uint32_t)(i % M), k = (uint32_t)(i / M);
    float val = vi;
    if (fin_mode == 2) { val += (float)m; }
    if (fin_mode == 3) { val += (float)k; }
    else if (fin_mode == 4) { if ((m == M / 2) && (k == K / 2)) { val += 1.0f; } }
    else if (fin_mode == 5) { val += det_hash_rand((uint32_t)i + SALT_SGEMM); }
    else if (fin_mode == 6) { val += (float)(m * 1000 + k); }
    a[i] = val;
  }
}
/* test/rtc/gen_data_sgemm_b.cucl:7-22 : b is K:N. mode>=100 -> identity. */
ORACLE_API void oracle_gen_sgemm_b(float *b, uint32_t K, uint32_t N, uint32_t mode, float vi) {
  uint64_t const n = (uint64_t)K * N;
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t const nn = (uint32_t)(i % N), k = (uint32_t)(i / N);
    float val = vi;
    if (mode == 2) { val += (float)nn; }
    if (mode == 3) { val += (float)k; }
    else if (mode == 4) { if ((nn == N / 2) && (k == K / 2)) { val += 1.0f; } }
    else if (mode == 5) { val += det_hash_rand((uint32_t)i + SALT_SGEMM); }
    else if (mode >= 100) { if (nn == k) { val += 1.0f; } }
    b[i] = val;
  }
}

/* ---- Convolution ---------------------------------------------------------------------------------- */

/* out[n,oc,oy,ox] = relu?( bias[oc] + sum_{ic,ky,kx} in[n,ic,oy*s+ky-p,ox*s+kx-p] * filts[oc,ic,ky,kx] )
 * Follows test/rtc/conv.cucl:25-52 (loop over filts_ix_out_chan_elem = in_chan:y:x row-major, one FMA per tap into
 * an fp32 accumulator that starts at 0) and src/cnn_codegen.cc:35-42,204-214 (bias added after the K loop, then
 * max(0,.) when conv_has_relu). Padding is mathematically-correct zero padding (every other reference variant
 * zero-fills: test/rtc/tconv_xpose_in.cucl:28-42, k1conv_xpose_in.cucl:9-16).
 * Output size: src/conv_util.cc:167-173. The per-output accumulation order (ic, then ky, then kx) is the
 * reference's; the loop nest is arranged so the innermost loop runs over ox for vectorisation. */
ORACLE_API int oracle_conv_fwd(float const *in, float const *filts, float const *biases, float *out, uint32_t N,
                               uint32_t C, uint32_t H, uint32_t W, uint32_t OC, uint32_t KH, uint32_t KW, uint32_t sy,
                               uint32_t sx, uint32_t py, uint32_t px, int relu) {
  if (H + 2 * py < KH || W + 2 * px < KW) { return -1; }
  int32_t const OH = (int32_t)((H + 2 * py - KH) / sy + 1), OW = (int32_t)((W + 2 * px - KW) / sx + 1);
  int64_t const jobs = (int64_t)N * OC;
#pragma omp parallel
  {
    float *acc = (float *)malloc(sizeof(float) * (size_t)OH * OW);
#pragma omp for schedule(dynamic, 4)
    for (int64_t job = 0; job < jobs; ++job) {
      uint32_t const n = (uint32_t)(job / OC), oc = (uint32_t)(job % OC);
      memset(acc, 0, sizeof(float) * (size_t)OH * OW);
      for (uint32_t ic = 0; ic < C; ++ic) {
        float const *inp = in + ((size_t)n * C + ic) * H * W;
        for (uint32_t ky = 0; ky < KH; ++ky) {
          for (uint32_t kx = 0; kx < KW; ++kx) {
            float const w = filts[(((size_t)oc * C + ic) * KH + ky) * KW + kx];
            /* valid ox range: 0 <= ox*sx + kx - px < W */
            int32_t ox_lo = 0, ox_hi = OW;
            { int32_t const off = (int32_t)kx - (int32_t)px;
              if (off < 0) { ox_lo = (-off + (int32_t)sx - 1) / (int32_t)sx; }
              int32_t const lim = (int32_t)W - off; /* need ox*sx < lim */
              int32_t const hi = (lim + (int32_t)sx - 1) / (int32_t)sx;
              if (hi < ox_hi) { ox_hi = hi; } }
            for (int32_t oy = 0; oy < OH; ++oy) {
              int32_t const iy = oy * (int32_t)sy + (int32_t)ky - (int32_t)py;
              if (iy < 0 || iy >= (int32_t)H) { continue; }
              float const *irow = inp + (size_t)iy * W + ((int32_t)kx - (int32_t)px);
              float *arow = acc + (size_t)oy * OW;
              if (sx == 1) {
                for (int32_t ox = ox_lo; ox < ox_hi; ++ox) { arow[ox] = fmaf(irow[ox], w, arow[ox]); }
              } else {
                for (int32_t ox = ox_lo; ox < ox_hi; ++ox) { arow[ox] = fmaf(irow[(size_t)ox * sx], w, arow[ox]); }
              }
            }
          }
        }
      }
      float const b = biases ? biases[oc] : 0.0f;
      float *o = out + ((size_t)n * OC + oc) * OH * OW;
      for (int32_t i = 0; i < OH * OW; ++i) {
        float v = acc[i] + b;
        if (relu) { v = (v > 0.0f) ? v : 0.0f; }
        o[i] = v;
      }
    }
    free(acc);
  }
  return 0;
}

/* Same algorithm and accumulation ORDER as oracle_conv_fwd, but the accumulator is double and the result is rounded to
 * fp32 once: the reference's arithmetic with its fp32 accumulation rounding noise removed. Used to separate "differs
 * from the reference's math" from "differs from the reference's rounding" (for K in the thousands and the +-5 hash
 * inputs, fp32 sequential accumulation alone is ~2e-3 mrd away from this; see DESIGN.md "Numerics"). */
ORACLE_API int oracle_conv_fwd_acc64(float const *in, float const *filts, float const *biases, float *out, uint32_t N,
                                     uint32_t C, uint32_t H, uint32_t W, uint32_t OC, uint32_t KH, uint32_t KW, uint32_t sy,
                                     uint32_t sx, uint32_t py, uint32_t px, int relu) {
  if (H + 2 * py < KH || W + 2 * px < KW) { return -1; }
  int32_t const OH = (int32_t)((H + 2 * py - KH) / sy + 1), OW = (int32_t)((W + 2 * px - KW) / sx + 1);
  int64_t const jobs = (int64_t)N * OC;
#pragma omp parallel
  {
    double *acc = (double *)malloc(sizeof(double) * (size_t)OH * OW);
#pragma omp for schedule(dynamic, 4)
    for (int64_t job = 0; job < jobs; ++job) {
      uint32_t const n = (uint32_t)(job / OC), oc = (uint32_t)(job % OC);
      memset(acc, 0, sizeof(double) * (size_t)OH * OW);
      for (uint32_t ic = 0; ic < C; ++ic) {
        float const *inp = in + ((size_t)n * C + ic) * H * W;
        for (uint32_t ky = 0; ky < KH; ++ky) {
          for (uint32_t kx = 0; kx < KW; ++kx) {
            double const w = filts[(((size_t)oc * C + ic) * KH + ky) * KW + kx];
            for (int32_t oy = 0; oy < OH; ++oy) {
              int32_t const iy = oy * (int32_t)sy + (int32_t)ky - (int32_t)py;
              if (iy < 0 || iy >= (int32_t)H) { continue; }
              for (int32_t ox = 0; ox < OW; ++ox) {
                int32_t const ix = ox * (int32_t)sx + (int32_t)kx - (int32_t)px;
                if (ix < 0 || ix >= (int32_t)W) { continue; }
                acc[(size_t)oy * OW + ox] += (double)inp[(size_t)iy * W + ix] * w;
              }
            }
          }
        }
      }
      double const b = biases ? biases[oc] : 0.0;
      float *o = out + ((size_t)n * OC + oc) * OH * OW;
      for (int32_t i = 0; i < OH * OW; ++i) {
        float v = (float)(acc[i] + b);
        if (relu) { v = (v > 0.0f) ? v : 0.0f; }
        o[i] = v;
      }
    }
    free(acc);
  }
  return 0;
}

/* ---- SGEMM ---------------------------------------------------------------------------------------- */

/* c[m,n] = sum_k a[k,m] * b[k,n]; a is K:M (pre-transposed A), b is K:N, c is M:N; fp32 FMA, k ascending.
 * test/rtc/sgemm.cucl:1-45 and src/cnn_codegen.cc:469-484. Blocked over (m,n) for cache; the per-output k order
 * is unchanged. */
ORACLE_API void oracle_sgemm(float const *a, float const *b, float *c, uint32_t M, uint32_t N, uint32_t K) {
  enum { MB = 8, NB = 256 };
  int64_t const mblks = (M + MB - 1) / MB, nblks = (N + NB - 1) / NB;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int64_t mb = 0; mb < mblks; ++mb) {
    for (int64_t nb = 0; nb < nblks; ++nb) {
      float acc[MB][NB];
      uint32_t const m0 = (uint32_t)mb * MB, n0 = (uint32_t)nb * NB;
      uint32_t const mlen = (M - m0 < MB) ? (M - m0) : MB, nlen = (N - n0 < NB) ? (N - n0) : NB;
      memset(acc, 0, sizeof(acc));
      for (uint32_t k = 0; k < K; ++k) {
        float const *brow = b + (size_t)k * N + n0;
        float const *arow = a + (size_t)k * M + m0;
        for (uint32_t mi = 0; mi < mlen; ++mi) {
          float const av = arow[mi];
          for (uint32_t ni = 0; ni < nlen; ++ni) { acc[mi][ni] = fmaf(av, brow[ni], acc[mi][ni]); }
        }
      }
      for (uint32_t mi = 0; mi < mlen; ++mi) {
        memcpy(c + (size_t)(m0 + mi) * N + n0, acc[mi], sizeof(float) * nlen);
      }
    }
  }
}

/* oracle_sgemm with a double accumulator (same k order), rounded to fp32 once. */
ORACLE_API void oracle_sgemm_acc64(float const *a, float const *b, float *c, uint32_t M, uint32_t N, uint32_t K) {
  enum { MB = 8, NB = 256 };
  int64_t const mblks = (M + MB - 1) / MB, nblks = (N + NB - 1) / NB;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
  for (int64_t mb = 0; mb < mblks; ++mb) {
    for (int64_t nb = 0; nb < nblks; ++nb) {
      double acc[MB][NB];
      uint32_t const m0 = (uint32_t)mb * MB, n0 = (uint32_t)nb * NB;
      uint32_t const mlen = (M - m0 < MB) ? (M - m0) : MB, nlen = (N - n0 < NB) ? (N - n0) : NB;
      memset(acc, 0, sizeof(acc));
      for (uint32_t k = 0; k < K; ++k) {
        float const *brow = b + (size_t)k * N + n0;
        float const *arow = a + (size_t)k * M + m0;
        for (uint32_t mi = 0; mi < mlen; ++mi) {
          double const av = arow[mi];
          for (uint32_t ni = 0; ni < nlen; ++ni) { acc[mi][ni] += av * (double)brow[ni]; }
        }
      }
      for (uint32_t mi = 0; mi < mlen; ++mi) {
        for (uint32_t ni = 0; ni < nlen; ++ni) { c[(size_t)(m0 + mi) * N + n0 + ni] = (float)acc[mi][ni]; }
      }
    }
  }
}

/* ---- Pooling -------------------------------------------------------------------------------------- */

/* test/rtc/pool.cucl:13-40 : window clipped to the input (padding never read), max init -FLT_MAX, avg divides by
 * the clipped count; loop order kx outer, ky inner (only matters for float avg summation order, preserved).
 * Output size (caller computes): src/conv_util.cc:198-204 ceil rule. */
ORACLE_API void oracle_pool_fwd(float const *in, float *out, uint32_t N, uint32_t C, uint32_t H, uint32_t W,
                                uint32_t OH, uint32_t OW, uint32_t KH, uint32_t KW, uint32_t sy, uint32_t sx,
                                uint32_t py, uint32_t px, int avg_pool) {
  int64_t const planes = (int64_t)N * C;
#pragma omp parallel for schedule(static)
  for (int64_t p = 0; p < planes; ++p) {
    float const *ip = in + (size_t)p * H * W;
    float *op = out + (size_t)p * OH * OW;
    for (uint32_t oy = 0; oy < OH; ++oy) {
      for (uint32_t ox = 0; ox < OW; ++ox) {
        float out_v = avg_pool ? 0.0f : -FLT_MAX;
        float avg_pool_sz = 0;
        for (int32_t kx = 0; kx != (int32_t)KW; ++kx) {
          for (int32_t ky = 0; ky != (int32_t)KH; ++ky) {
            int32_t const in_y = (int32_t)(oy * sy) + ky - (int32_t)py;
            int32_t const in_x = (int32_t)(ox * sx) + kx - (int32_t)px;
            if (in_y >= 0 && in_x >= 0 && in_x < (int32_t)W && in_y < (int32_t)H) {
              float const v = ip[(size_t)in_y * W + in_x];
              if (avg_pool) { out_v += v; avg_pool_sz += 1; }
              else if (v > out_v) { out_v = v; }
            }
          }
        }
        if (avg_pool) { out_v /= avg_pool_sz; }
        op[(size_t)oy * OW + ox] = out_v;
      }
    }
  }
}

/* ---- LRN (across channels) ------------------------------------------------------------------------ */

/* test/rtc/lrn.cucl:35-50 (LRN_MATCH_CAFFE branch): running sum of squares over a ring buffer of local_size
 * channels with add-new / subtract-old, scale = powf(k + ls_sum*(alpha/local_size), -beta). */
ORACLE_API void oracle_lrn_fwd(float const *in, float *out, uint32_t N, uint32_t C, uint32_t H, uint32_t W,
                               uint32_t local_size, float alpha, float beta, float k) {
  int64_t const pels = (int64_t)N * H * W;
  size_t const cs = (size_t)H * W;
  int32_t const hls = (int32_t)(local_size >> 1);
  float const alpha_over_ls = (float)(alpha) / (float)(local_size);
#pragma omp parallel for schedule(static)
  for (int64_t pel = 0; pel < pels; ++pel) {
    uint32_t const n = (uint32_t)(pel / (int64_t)cs);
    size_t const base = (size_t)n * C * cs + (size_t)(pel % (int64_t)cs);
    float ls_buf[64];
    for (uint32_t i = 0; i < local_size; ++i) { ls_buf[i] = 0.0f; }
    float ls_sum = 0.0f;
    for (int32_t ic = 0; ic < (int32_t)C + hls; ++ic) {
      int32_t const lsb_ix = ic % (int32_t)local_size;
      float const ls_old = ls_buf[lsb_ix];
      ls_buf[lsb_ix] = (ic < (int32_t)C) ? in[base + (size_t)ic * cs] : 0.0f;
      ls_sum = fmaf(ls_buf[lsb_ix], ls_buf[lsb_ix], ls_sum); /* fmad contraction, as NVRTC compiles lrn.cucl:43 */
      ls_sum = fmaf(-ls_old, ls_old, ls_sum);
      if (ic >= hls) {
        int32_t const oc = ic - hls;
        float const scale_base = fmaf(ls_sum, alpha_over_ls, k);
        float const scale = powf(scale_base, -beta);
        out[base + (size_t)oc * cs] = ls_buf[(lsb_ix + (int32_t)local_size - hls) % (int32_t)local_size] * scale;
      }
    }
  }
}

/* ---- ReLU / Softmax / Concat copy / N-ary sum ----------------------------------------------------- */

/* test/rtc/relu.cucl:1-5 */
ORACLE_API void oracle_relu(float *inout, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) { inout[i] = (inout[i] <= 0) ? 0.0f : inout[i]; }
}

/* test/rtc/softmax.cucl:6-21 : per (img,y,x) over chan; NOTE running max starts at 0.0f, not -inf (line 8). */
ORACLE_API void oracle_softmax(float const *in, float *prob, uint32_t N, uint32_t C, uint32_t H, uint32_t W) {
  size_t const cs = (size_t)H * W;
  int64_t const pels = (int64_t)N * H * W;
#pragma omp parallel for schedule(static)
  for (int64_t pel = 0; pel < pels; ++pel) {
    size_t const base = (size_t)(pel / (int64_t)cs) * C * cs + (size_t)(pel % (int64_t)cs);
    float pel_sum = 0.0f, pel_max = 0.0f;
    for (uint32_t c = 0; c < C; ++c) { float const v = in[base + c * cs]; pel_max = (v > pel_max) ? v : pel_max; }
    for (uint32_t c = 0; c < C; ++c) { float const v = expf(in[base + c * cs] - pel_max); prob[base + c * cs] = v; pel_sum += v; }
    for (uint32_t c = 0; c < C; ++c) { prob[base + c * cs] /= pel_sum; }
  }
}

/* test/rtc/copy.cucl:5-9 + src/rtc_fwd.cc:267-280 : Concat = one copy per input into out[:, ocix:ocix+C]. */
ORACLE_API void oracle_concat_copy(float const *in, float *out, uint32_t N, uint32_t C, uint32_t H, uint32_t W,
                                   uint32_t out_C, uint32_t ocix) {
  size_t const cs = (size_t)H * W;
  for (uint32_t n = 0; n < N; ++n) {
    memcpy(out + ((size_t)n * out_C + ocix) * cs, in + (size_t)n * C * cs, sizeof(float) * C * cs);
  }
}

/* test/rtc/reduce.cucl:1-14 + src/cnn_codegen.cc:28-34 : v = 0; v += ins[i][ix] for each input in order. */
ORACLE_API void oracle_reduce_sum(float const *const *ins, uint32_t ins_num, float *out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    float v = 0;
    for (uint32_t j = 0; j < ins_num; ++j) { v += ins[j][i]; }
    out[i] = v;
  }
}

/* ---- comparison metric + digest helpers ----------------------------------------------------------- */

/* src/boda_base.cc:140-154 : |v2-v1| / max(min_sig_mag,|v1|,|v2|) in double. */
static inline double min_sig_mag_rel_diff(double min_sig_mag, double v1, double v2) {
  double const a1 = (v1 < 0) ? -v1 : v1, a2 = (v2 < 0) ? -v2 : v2;
  double amax = (a1 > a2) ? a1 : a2;
  if (min_sig_mag > amax) { amax = min_sig_mag; }
  double const d = v2 - v1;
  return ((d < 0) ? -d : d) / amax;
}

/* src/boda_base.cc:157-206 (ssds_diff_t): res = {ssds, sds, mad, mrd, sum1, sum2, num_diff, has_nan}. */
ORACLE_API void oracle_ssds_diff(float const *o1, float const *o2, uint64_t sz, double *res) {
  double ssds = 0, sds = 0, mad = 0, mrd = 0, sum1 = 0, sum2 = 0, num_diff = 0;
  for (uint64_t i = 0; i < sz; ++i) {
    sum1 += (double)o1[i];
    sum2 += (double)o2[i];
    double const d = (double)o2[i] - (double)o1[i];
    sds += d;
    ssds += d * d;
    double const ad = (d < 0) ? -d : d;
    if (ad > mad) { mad = ad; }
    double const rd = min_sig_mag_rel_diff(1.0, o1[i], o2[i]);
    if (rd > mrd) { mrd = rd; }
    if (o1[i] != o2[i]) { num_diff += 1; }
  }
  res[0] = ssds; res[1] = sds; res[2] = mad; res[3] = mrd; res[4] = sum1; res[5] = sum2; res[6] = num_diff;
  res[7] = (isnan(ssds) || isnan(sds) || isnan(mad)) ? 1.0 : 0.0;
}

/* src/boda_base.cc:266-272 (nda_digest_T::get_samp): sequential *float* sum with a uint32_t index. */
ORACLE_API float oracle_strided_sum(float const *ve, uint64_t sz, uint64_t offset, uint64_t stride) {
  float sv = 0.0f;
  for (uint32_t i = (uint32_t)offset; i < sz; i += (uint32_t)stride) { sv += ve[i]; }
  return sv;
}

/* src/boda_base.cc:253-256 : min/max scan. */
ORACLE_API void oracle_min_max(float const *ve, uint64_t sz, float *min_v, float *max_v) {
  float mn = FLT_MAX, mx = -FLT_MAX;
  for (uint64_t i = 0; i < sz; ++i) { if (ve[i] < mn) { mn = ve[i]; } if (ve[i] > mx) { mx = ve[i]; } }
  *min_v = mn; *max_v = mx;
}

/* bench.py's reference arm: use all host cores even when the launcher (torch.distributed.run) exported OMP_NUM_THREADS=1 */
ORACLE_API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) { omp_set_num_threads(n); }
#else
  (void)n;
#endif
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
