// InstanceNorm statistics + fused normalise / affine / residual / activation.
//
// Reference semantics: nn.InstanceNorm3d(eps=1e-5, biased variance, per (sample, channel) over D*H*W) followed by
// PReLU / LeakyReLU as ordered by ADN("NDA") (monai/networks/blocks/acti_norm.py:19-101), and the residual
// tail of UnetResBlock (monai/networks/blocks/dynunet_block.py:97-111: out = lrelu(norm2(conv2) + norm3(conv3(x)))).
#include "common.cuh"
#include "stats.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

// One block per (n,c) plane: every thread adds a fixed strided subset, the warp totals are combined in warp order --
// the same bits on every run (no floating-point atomics).
template <typename T>
__global__ void __launch_bounds__(1024) instnorm_stats_kernel(const T* __restrict__ x, int C, long long S,
                                                              long long stride_n, float* __restrict__ stats) {
  const int nc = blockIdx.x;
  const int n = nc / C, c = nc % C;
  const T* p = x + (long long)n * stride_n + (long long)c * S;
  float s = 0.f, q = 0.f;
  for (long long i = threadIdx.x; i < S; i += blockDim.x) {
    const float v = io<T>::ld(p + i);
    s += v; q = fmaf(v, v, q);
  }
  s = warp_sum(s); q = warp_sum(q);
  __shared__ float ss[32], sq[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { ss[wid] = s; sq[wid] = q; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ts = 0.0, tq = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { ts += (double)ss[w]; tq += (double)sq[w]; }
    stats[2 * nc] = (float)ts; stats[2 * nc + 1] = (float)tq;
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

// chunked form: block (chunk, plane) adds its slice in a fixed order and writes one {sum, sumsq} pair; stats_finish_kernel then adds
// the chunks of a plane in chunk order in double precision
template <typename T>
__global__ void __launch_bounds__(256) instnorm_partial_kernel(const T* __restrict__ x, int C, long long S, long long stride_n, long long chunk,
                                                               float* __restrict__ part) {
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
  if (threadIdx.x == 0) {
    float ts = 0.f, tq = 0.f;
    for (int w = 0; w < 8; ++w) { ts += ss[w]; tq += sq[w]; }
    part[((long long)nc * gridDim.x + blockIdx.x) * 2] = ts;
    part[((long long)nc * gridDim.x + blockIdx.x) * 2 + 1] = tq;
  }
}

static int instnorm_chunks(int N, int C, long long S) {
  // enough blocks to fill the chip when there are few planes, at least 16 K elements per block
  const long long want = (long long)num_sms() * 4 / std::max(1LL, (long long)N * C);
  return (int)std::max<long long>(1, std::min<long long>(std::min<long long>(want, 256), S / 16384));
}

extern "C" long long b200_instnorm_stats_workspace_bytes(int N, int C, long long S) {
  if (N <= 0 || C <= 0 || S <= 0) return -1;
  const int ch = instnorm_chunks(N, C, S);
  return ch > 1 ? (long long)N * C * ch * 2 * (long long)sizeof(float) : 0;
}

extern "C" int b200_instnorm_stats(const void* x, int dtype, int N, int C, long long S, long long x_stride_n,
                                   float* stats, void* workspace, void* stream) {
  B200_REQUIRE(x && stats, "instnorm_stats: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && S > 0, "instnorm_stats: empty problem");
  B200_REQUIRE(dtype == B200_DT_F16 || dtype == B200_DT_F32, "instnorm_stats: bad dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const int ch = workspace ? instnorm_chunks(N, C, S) : 1;
  if (ch > 1) {
    B200_REQUIRE((long long)N * C <= 65535, "instnorm_stats: N*C too large for one launch");
    const long long chunk = ((S + ch - 1) / ch + 255) / 256 * 256;
    dim3 grid(ch, N * C);
    if (dtype == B200_DT_F16) instnorm_partial_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, C, S, x_stride_n, chunk, (float*)workspace);
    else instnorm_partial_kernel<float><<<grid, 256, 0, st>>>((const float*)x, C, S, x_stride_n, chunk, (float*)workspace);
    B200_LAUNCH_CHECK("instnorm_partial_kernel");
    return launch_stats_finish((const float*)workspace, (long long)N * C, ch, 1, 1, 1, stats, st);
  }
  const int threads = S >= 32768 ? 1024 : (S >= 4096 ? 256 : 64);
  dim3 grid(N * C);
  if (dtype == B200_DT_F16) instnorm_stats_kernel<__half><<<grid, threads, 0, st>>>((const __half*)x, C, S, x_stride_n, stats);
  else instnorm_stats_kernel<float><<<grid, threads, 0, st>>>((const float*)x, C, S, x_stride_n, stats);
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

// ---------------------------------------------------------------------------------------------------------------------
// finishing pass of the deterministic statistics (stats.cuh)
// ---------------------------------------------------------------------------------------------------------------------
namespace b200 {

// one block per (batch item, N tile) group; blockDim = (2*NT, L): thread (i, l) adds the rows r = l, l+L, ... of column i
// in increasing r, then the L lane totals are added in increasing l -- a fixed order, in double precision.
__global__ void stats_finish_kernel(const float* __restrict__ partials, int rows, int nt2, int n_tiles, int NT, int Cout,
                                    float* __restrict__ stats) {
  extern __shared__ double s_fin[];   // [L][nt2]
  const long long g = blockIdx.x;
  const int i = threadIdx.x, l = threadIdx.y, L = blockDim.y;
  const float* src = partials + g * (long long)rows * nt2;
  double acc = 0.0;
  for (int r = l; r < rows; r += L) acc += (double)src[(long long)r * nt2 + i];
  s_fin[l * nt2 + i] = acc;
  __syncthreads();
  if (l == 0) {
    double t = 0.0;
    for (int k = 0; k < L; ++k) t += s_fin[k * nt2 + i];
    const int n = (int)(g / n_tiles), nt = (int)(g % n_tiles);
    const int col = nt * NT + i / 2;
    if (col < Cout) stats[((long long)n * Cout + col) * 2 + (i & 1)] = (float)t;
  }
}

int launch_stats_finish(const float* partials, long long groups, int rows, int NT, int n_tiles, int Cout, float* stats, cudaStream_t st) {
  const int nt2 = 2 * NT;
  B200_REQUIRE(nt2 <= 1024 && groups > 0 && groups < (1LL << 31), "stats_finish: bad sizes");
  int L = 1024 / nt2;
  L = L > 8 ? 8 : (L < 1 ? 1 : L);
  dim3 block(nt2, L);
  stats_finish_kernel<<<(unsigned)groups, block, (size_t)L * nt2 * sizeof(double), st>>>(partials, rows, nt2, n_tiles, NT, Cout, stats);
  B200_LAUNCH_CHECK("stats_finish_kernel");
  return B200_OK;
}

}  // namespace b200
