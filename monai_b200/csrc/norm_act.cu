// InstanceNorm statistics + fused normalise / affine / residual / activation.
//
// Reference semantics: nn.InstanceNorm3d(eps=1e-5, biased variance, per (sample, channel) over D*H*W) followed by
// PReLU / LeakyReLU as ordered by ADN("NDA") (monai/networks/blocks/acti_norm.py:19-101), and the residual
// tail of UnetResBlock (monai/networks/blocks/dynunet_block.py:97-111: out = lrelu(norm2(conv2) + norm3(conv3(x)))).
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

// One block handles `chunk` consecutive elements of one (n,c) plane; partial sums go through fp32 atomics.
template <typename T>
__global__ void __launch_bounds__(256) instnorm_stats_kernel(const T* __restrict__ x, int C, long long S,
                                                             long long stride_n, long long chunk, float* __restrict__ stats) {
  const int nc = blockIdx.y;
  const int n = nc / C, c = nc % C;
  const T* p = x + (long long)n * stride_n + (long long)c * S;
  const long long lo = (long long)blockIdx.x * chunk, hi = min(S, lo + chunk);
  float s = 0.f, q = 0.f;
  for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float v = io<T>::ld(p + i);
    s += v; q = fmaf(v, v, q);
  }
  s = warp_sum(s); q = warp_sum(q);
  __shared__ float ss[8], sq[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { ss[wid] = s; sq[wid] = q; }
  __syncthreads();
  if (wid == 0) {
    s = lane < (blockDim.x >> 5) ? ss[lane] : 0.f;
    q = lane < (blockDim.x >> 5) ? sq[lane] : 0.f;
    s = warp_sum(s); q = warp_sum(q);
    if (lane == 0) { atomicAdd(stats + 2 * nc, s); atomicAdd(stats + 2 * nc + 1, q); }
  }
}

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case 1: case 2: return v >= 0.f ? v : v * slope;
    case 3: return fmaxf(v, 0.f);
    case 4: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

struct NormActP {
  const void* x; void* y; const void* res;
  int C; long long S, xs_n, ys_n, rs_n;
  const float* stats; const float* res_stats; float eps;
  const float* gamma; const float* beta;
  int act; float slope; const float* slope_ptr; int n_slope;
};

template <typename T>
__global__ void __launch_bounds__(256) norm_act_kernel(NormActP p) {
  const int c = blockIdx.y, n = blockIdx.z;
  float scale = 1.f, shift = 0.f, rscale = 1.f, rshift = 0.f;
  const float invS = 1.f / (float)p.S;
  if (p.stats) {
    const float s = p.stats[2 * (n * p.C + c)], q = p.stats[2 * (n * p.C + c) + 1];
    const float mean = s * invS;
    const float var = fmaxf(q * invS - mean * mean, 0.f);
    const float rstd = 1.f / sqrtf(var + p.eps);
    const float g = p.gamma ? p.gamma[c] : 1.f, b = p.beta ? p.beta[c] : 0.f;
    scale = rstd * g; shift = b - mean * rstd * g;
  } else {  // pre-folded per-channel affine (e.g. eval-mode BatchNorm)
    scale = p.gamma ? p.gamma[c] : 1.f; shift = p.beta ? p.beta[c] : 0.f;
  }
  if (p.res_stats) {
    const float s = p.res_stats[2 * (n * p.C + c)], q = p.res_stats[2 * (n * p.C + c) + 1];
    const float mean = s * invS;
    const float var = fmaxf(q * invS - mean * mean, 0.f);
    const float rstd = 1.f / sqrtf(var + p.eps);
    rscale = rstd; rshift = -mean * rstd;
  }
  const float slope = (p.act == 2 && p.slope_ptr) ? p.slope_ptr[c % p.n_slope] : p.slope;
  const T* x = (const T*)p.x + (long long)n * p.xs_n + (long long)c * p.S;
  T* y = (T*)p.y + (long long)n * p.ys_n + (long long)c * p.S;
  const T* r = p.res ? (const T*)p.res + (long long)n * p.rs_n + (long long)c * p.S : nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.S; i += (long long)gridDim.x * blockDim.x) {
    float v = fmaf(io<T>::ld(x + i), scale, shift);
    if (r) v += fmaf(io<T>::ld(r + i), rscale, rshift);
    io<T>::st(y + i, apply_act(v, p.act, slope));
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_instnorm_stats(const void* x, int dtype, int N, int C, long long S, long long x_stride_n,
                                   float* stats, void* stream) {
  B200_REQUIRE(x && stats, "instnorm_stats: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && S > 0, "instnorm_stats: empty problem");
  B200_REQUIRE((long long)N * C <= 65535, "instnorm_stats: N*C too large for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * N * C, st));
  // aim for ~4 waves of blocks over the machine, at least 2048 elements per block
  long long want = (long long)num_sms() * 8 / ((long long)N * C) + 1;
  long long chunk = std::max<long long>(2048, (S + want - 1) / want);
  chunk = (chunk + 255) / 256 * 256;
  dim3 grid(ceil_div(S, chunk), N * C);
  if (dtype == B200_DT_F16) instnorm_stats_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, C, S, x_stride_n, chunk, stats);
  else if (dtype == B200_DT_F32) instnorm_stats_kernel<float><<<grid, 256, 0, st>>>((const float*)x, C, S, x_stride_n, chunk, stats);
  else return set_err(B200_ERR_INVALID, "instnorm_stats: bad dtype");
  B200_LAUNCH_CHECK("instnorm_stats_kernel");
  return B200_OK;
}

extern "C" int b200_norm_act(const void* x, int dtype, int N, int C, long long S, long long x_stride_n,
                             const float* stats, float eps, const float* gamma, const float* beta, const void* res,
                             long long res_stride_n, const float* res_stats, int act, float slope,
                             const float* slope_ptr, int n_slope, void* y, long long y_stride_n, void* stream) {
  B200_REQUIRE(x && y, "norm_act: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && S > 0, "norm_act: empty problem");
  B200_REQUIRE(act >= 0 && act <= 4, "norm_act: unknown activation %d", act);
  B200_REQUIRE(C <= 65535 && N <= 65535, "norm_act: N or C too large for one launch");
  B200_REQUIRE(!(act == 2 && slope_ptr) || n_slope >= 1, "norm_act: prelu needs n_slope >= 1");
  NormActP p;
  p.x = x; p.y = y; p.res = res; p.C = C; p.S = S; p.xs_n = x_stride_n; p.ys_n = y_stride_n; p.rs_n = res_stride_n;
  p.stats = stats; p.res_stats = res_stats; p.eps = eps; p.gamma = gamma; p.beta = beta;
  p.act = act; p.slope = slope; p.slope_ptr = slope_ptr; p.n_slope = n_slope;
  long long want = (long long)num_sms() * 16 / ((long long)N * C) + 1;
  int bx = (int)std::min<long long>(want, (S + 255) / 256);
  dim3 grid(max(bx, 1), C, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F16) norm_act_kernel<__half><<<grid, 256, 0, st>>>(p);
  else if (dtype == B200_DT_F32) norm_act_kernel<float><<<grid, 256, 0, st>>>(p);
  else return set_err(B200_ERR_INVALID, "norm_act: bad dtype");
  B200_LAUNCH_CHECK("norm_act_kernel");
  return B200_OK;
}
