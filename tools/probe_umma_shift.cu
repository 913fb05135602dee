// probe_umma_shift.cu -- hardware probe: can a tcgen05 shared-memory descriptor (K-major, SWIZZLE_128B) start at an arbitrary
// 128-byte ROW offset inside a TMA-written tile, and which `base_offset` does it need?  This decides whether one halo tile of
// activations can feed all (ky,kx) taps of a stride-1 convolution ("tap reuse") without re-loading it per tap.
//
// One CTA: TMA loads A[256 x 64] fp16 (two 128-row boxes, contiguous, 1024-aligned) and B[128 x 64]; for each shift s the MMA reads
// A rows [s, s+128) through a descriptor whose start address is advanced by s*128 bytes; D = A[s:s+128] * B^T is written out.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I boda_b200/csrc tools/probe_umma_shift.cu -o /tmp/probe_umma_shift
#include "umma.cuh"
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

using namespace b200;

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, float *out, int shift, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_sm = smem;               // 256 rows x 128 B = 32 KB
  uint8_t *b_sm = smem + 256 * 128;   // 128 rows x 128 B = 16 KB
  uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + 48 * 1024);
  uint64_t *done_bar = full_bar + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(full_bar + 2);
  int const warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(full_bar, 1); mbar_init(done_bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc<128>(tmem_ptr); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t const tmem_base = *tmem_ptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(full_bar, 48 * 1024);
    tma_load_2d(a_sm, &amap, full_bar, 0, 0);
    tma_load_2d(a_sm + 128 * 128, &amap, full_bar, 0, 128);
    tma_load_2d(b_sm, &bmap, full_bar, 0, 0);
    mbar_wait(full_bar, 0);
    tc_fence_after();
    uint32_t const a_addr = smem_u32(a_sm) + shift * 128;
    uint64_t adesc = make_kmajor_sw128_desc(a_addr);
    if (mode == 1) { adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49; }  // base_offset = row phase inside the 1024-byte swizzle atom
    uint64_t const bdesc = make_kmajor_sw128_desc(smem_u32(b_sm));
    uint32_t const idesc = make_idesc_f16(0, 128, 128);
    for (int k = 0; k < 4; ++k) { umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u); }
    umma_commit(done_bar);
  }
  mbar_wait(done_bar, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int j0 = 0; j0 < 128; j0 += 32) {
    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + j0, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) { out[(warp * 32 + lane) * 128 + j0 + j] = __uint_as_float(r[j]); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tmem_dealloc<128>(tmem_base); }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  int const AR = 256, BR = 128, K = 64;
  std::vector<__half> ha(AR * K), hb(BR * K);
  std::vector<float> fa(AR * K), fb(BR * K);
  for (int i = 0; i < AR * K; ++i) { fa[i] = float((i * 7 + (i / K) * 3) % 17 - 8); ha[i] = __float2half(fa[i]); }
  for (int i = 0; i < BR * K; ++i) { fb[i] = float((i * 5 + (i / K)) % 13 - 6); hb[i] = __float2half(fb[i]); }
  __half *da, *db; float *dout;
  CK(cudaMalloc(&da, AR * K * 2)); CK(cudaMalloc(&db, BR * K * 2)); CK(cudaMalloc(&dout, 128 * 128 * 4));
  CK(cudaMemcpy(da, ha.data(), AR * K * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, hb.data(), BR * K * 2, cudaMemcpyHostToDevice));
  auto mk = [&](void *base, uint64_t rows) {
    CUtensorMap m; cuuint64_t gd[2] = {64, rows}; cuuint64_t gs[1] = {128}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", int(r)); exit(1); }
    return m;
  };
  CUtensorMap amap = mk(da, AR), bmap = mk(db, BR);
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
  std::vector<float> out(128 * 128);
  int const shifts[] = {0, 8, 1, 2, 3, 7, 9, 15, 16, 17, 33, 100, 127};
  for (int mode = 0; mode < 2; ++mode) {
    for (int s : shifts) {
      CK(cudaMemset(dout, 0, 128 * 128 * 4));
      probe_kernel<<<1, 128, 52 * 1024>>>(amap, bmap, dout, s, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %3d: kernel error %s\n", mode, s, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(out.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 128; ++n) {
        double ref = 0; for (int k = 0; k < K; ++k) { ref += double(fa[(m + s) * K + k]) * fb[n * K + k]; }
        double d = std::fabs(ref - out[m * 128 + n]); if (d > maxerr) { maxerr = d; } if (d > 1e-3) { ++bad; }
      }
      printf("mode(base_offset=%s) shift %3d: max|err| %.3f bad %d/16384 -> %s\n", mode ? "row&7" : "0", s, maxerr, bad, bad ? "MISMATCH" : "ok");
    }
  }
  return 0;
}
