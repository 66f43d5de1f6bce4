// exact-form GELU for the tensor-core epilogues (gemm_tc.cu, mlp_fused_tc.cu)
#pragma once
#include <cuda_runtime.h>

namespace b200 {

// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf(z) ~ z * P(z^2) on |z| <= 3.3 (degree-9 minimax fit, |erf error| < 1.2e-5,
// |GELU error| < 1.3e-4 absolute / 3e-5 relative for |x| > 1 -- far below the fp16 resolution of the stored activation)
// and erf = +-1 beyond.  16 FP32 pipe operations and no MUFU: the erf/exp formulation was bound by the 16-lane SFU.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fminf(fmaxf(x * 0.70710678118654752f, -3.3f), 3.3f);
  const float u = z * z;
  float p = -1.910707594e-09f;
  p = fmaf(p, u, 1.189451195e-07f);
  p = fmaf(p, u, -3.287776810e-06f);
  p = fmaf(p, u, 5.362731304e-05f);
  p = fmaf(p, u, -5.806723683e-04f);
  p = fmaf(p, u, 4.467851002e-03f);
  p = fmaf(p, u, -2.555716617e-02f);
  p = fmaf(p, u, 1.115591214e-01f);
  p = fmaf(p, u, -3.755447127e-01f);
  p = fmaf(p, u, 1.128300576e+00f);
  const float h = 0.5f * x;
  return fmaf(h, z * p, h);
}

}  // namespace b200
