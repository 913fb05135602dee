// scale.cuh -- the per-tensor power-of-two operand scale of the fp16 hi/lo planes, shared by the pack kernels and the contraction epilogues.
#pragma once

namespace b200 {

// scale = 2^(13 - floor(log2(absmax))): scaled values land in [2^13, 2^14), well inside fp16 range with most lo-plane residuals normal
__device__ __forceinline__ float scale_from_absmax_bits(unsigned int bits) {
  float const m = __uint_as_float(bits);
  float s = 1.0f;
  if (m > 0.0f && isfinite(m)) {
    int e;
    frexpf(m, &e);  // m = f * 2^e, f in [0.5,1)  -> floor(log2 m) = e-1
    int sh = 13 - (e - 1);
    sh = max(-100, min(100, sh));
    s = ldexpf(1.0f, sh);
  }
  return s;
}

}  // namespace b200
