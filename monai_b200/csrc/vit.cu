// Channels-first transformer pieces for the ViT encoder of UNETR: tokens are kept as [N, C, S] (the layout
// the convolution kernels use, S = number of patches), so nn.Linear runs as a 1x1x1 b200_conv3d_direct and only two operations need
// kernels of their own:
//   * b200_layernorm_cf   nn.LayerNorm(C) over the channel axis of every token (monai/networks/blocks/transformerblock.py:94-99,
//                         monai/networks/nets/vit.py:128) -- two-pass mean / variance in fp32, affine per channel;
//   * b200_mhsa_cf        softmax(q k^T * scale) v per (batch item, head) (SABlock.forward, monai/networks/blocks/selfattention.py:
//                         170-217 without mask / relative positions): the projection's channels are ordered (q|k|v, head, dim)
//                         ("b h (qkv l d) -> qkv b l h d"), the output's (head, dim) ("b l h d -> b h (l d)").
// These are the fp32-faithful generic forms (CUDA cores, fp32 accumulation, fp16 or fp32 storage); the windowed attention of
// SwinUNETR has its own tensor-core kernels (attn_tc.cu).
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

template <typename T>
__global__ void __launch_bounds__(256) layernorm_cf_kernel(const T* __restrict__ x, T* __restrict__ y, int C, long long S, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (s >= S) return;
  const T* xs = x + (long long)n * C * S + s;
  T* ys = y + (long long)n * C * S + s;
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum += io<T>::ld(xs + (long long)c * S);
  const float mean = sum / (float)C;
  float var = 0.f;
  for (int c = 0; c < C; ++c) { const float d = io<T>::ld(xs + (long long)c * S) - mean; var = fmaf(d, d, var); }
  const float rstd = 1.f / sqrtf(var / (float)C + eps);
  for (int c = 0; c < C; ++c) {
    const float v = (io<T>::ld(xs + (long long)c * S) - mean) * rstd;
    io<T>::st(ys + (long long)c * S, v * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f));
  }
}

// Non-overlapping patches as channels: out[n, ((c*pd + a)*ph + b)*pw + e, ((gd*Gh + gh)*Gw + gw)] = x[n, c, gd*pd + a, gh*ph + b, gw*pw + e].
// With it the patch projection of PatchEmbeddingBlock (a convolution with kernel = stride = patch: blocks/patchembedding.py:104-108)
// is a Linear over the channel axis, weight.reshape(hidden, C * pd * ph * pw).
template <typename T>
__global__ void __launch_bounds__(256) patchify_kernel(const T* __restrict__ x, T* __restrict__ y, int C, int D, int H, int W, int pd, int ph, int pw) {
  const int Gd = D / pd, Gh = H / ph, Gw = W / pw;
  const long long S = (long long)Gd * Gh * Gw, P = (long long)pd * ph * pw;
  const long long total = (long long)C * P * S;
  const int n = blockIdx.y;
  const T* xn = x + (long long)n * C * D * H * W;
  T* yn = y + (long long)n * total;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S, k = i / S;
    const int gw = (int)(s % Gw), gh = (int)((s / Gw) % Gh), gd = (int)(s / ((long long)Gw * Gh));
    const int e = (int)(k % pw), b = (int)((k / pw) % ph), a = (int)((k / ((long long)pw * ph)) % pd), c = (int)(k / P);
    yn[i] = xn[(((long long)c * D + gd * pd + a) * H + gh * ph + b) * W + gw * pw + e];
  }
}

constexpr int kMhsaQ = 64;    // queries per block (one per thread)
constexpr int kMhsaKC = 32;   // keys staged per step

// block = (query tile, head, batch item); thread = one query with q[D] and the output accumulator in registers; K / V chunks of 32 keys
// are staged in shared memory as fp32 [D][32] and read as broadcasts; online softmax in the natural-exponent domain.
// Windowed form (win > 0; WindowAttention.forward, monai/networks/nets/swin_unetr.py:509-532): the token axis holds nW windows of
// `win` tokens, a query attends to the keys of its window only, `bias[head][i][j]` (window-local i, j; the gathered relative position
// bias) is added to the scaled scores, and with `region` ([nW][win] labels of compute_mask, swin_unetr.py:779-816) pairs from
// different regions get -100 (the shifted-window mask).  blockIdx.x = window * tiles_per_window + tile.
template <typename T, int D>
__global__ void __launch_bounds__(kMhsaQ) mhsa_cf_kernel(const T* __restrict__ qkv, T* __restrict__ out, int heads, long long S, float scale, int win,
                                                         const float* __restrict__ bias, const int32_t* __restrict__ region) {
  __shared__ float s_k[D][kMhsaKC];
  __shared__ float s_v[D][kMhsaKC];
  __shared__ int s_reg[kMhsaKC];
  const int head = blockIdx.y, n = blockIdx.z;
  long long k_lo = 0, k_hi = S, qi;
  int w = 0, qloc;
  if (win > 0) {
    const int tiles = (win + kMhsaQ - 1) / kMhsaQ;
    w = blockIdx.x / tiles;
    qloc = (blockIdx.x % tiles) * kMhsaQ + threadIdx.x;
    k_lo = (long long)w * win; k_hi = k_lo + win;
    qi = k_lo + qloc;
  } else {
    qi = (long long)blockIdx.x * kMhsaQ + threadIdx.x;
    qloc = (int)qi;
  }
  const bool live = win > 0 ? qloc < win : qi < S;
  const long long HD = (long long)heads * D;
  const T* qp = qkv + ((long long)n * 3 * HD + (long long)head * D) * S;
  const T* kp = qp + HD * S;
  const T* vp = kp + HD * S;
  float q[D], acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { q[d] = live ? io<T>::ld(qp + (long long)d * S + qi) * scale : 0.f; acc[d] = 0.f; }
  const float* brow = (bias && live) ? bias + ((long long)head * win + qloc) * win : nullptr;
  const int myreg = (region && live) ? region[(long long)w * win + qloc] : 0;
  float m = -INFINITY, l = 0.f;
  for (long long j0 = k_lo; j0 < k_hi; j0 += kMhsaKC) {
    __syncthreads();
    for (int e = threadIdx.x; e < D * kMhsaKC; e += kMhsaQ) {
      const int d = e / kMhsaKC, jj = e % kMhsaKC;
      const bool ok = j0 + jj < k_hi;
      s_k[d][jj] = ok ? io<T>::ld(kp + (long long)d * S + j0 + jj) : 0.f;
      s_v[d][jj] = ok ? io<T>::ld(vp + (long long)d * S + j0 + jj) : 0.f;
    }
    if (region && threadIdx.x < kMhsaKC) s_reg[threadIdx.x] = j0 + threadIdx.x < k_hi ? region[j0 + threadIdx.x] : 0;   // region is [nW][win]: index = w*win + local
    __syncthreads();
    const int nk = (int)min((long long)kMhsaKC, k_hi - j0);
    for (int jj = 0; jj < nk; ++jj) {
      float sc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) sc = fmaf(q[d], s_k[d][jj], sc);
      if (brow) sc += brow[(int)(j0 - k_lo) + jj];
      if (region && s_reg[jj] != myreg) sc += -100.f;
      const float mn = fmaxf(m, sc);
      const float corr = __expf(m - mn), p = __expf(sc - mn);
      l = l * corr + p;
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = fmaf(p, s_v[d][jj], acc[d] * corr);
      m = mn;
    }
  }
  if (live) {
    const float inv = 1.f / l;
    T* op = out + ((long long)n * HD + (long long)head * D) * S + qi;
#pragma unroll
    for (int d = 0; d < D; ++d) io<T>::st(op + (long long)d * S, acc[d] * inv);
  }
}

// out[n, c, r] = src[r] >= 0 ? x[n, c, src[r]] : 0: window partition + cyclic shift + zero padding of SwinTransformerBlock
// (swin_unetr.py:596-625) and, with the inverse table, window_reverse + roll back + crop (:626-648), on channels-first tokens.
template <typename T>
__global__ void __launch_bounds__(256) gather_cf_kernel(const T* __restrict__ x, T* __restrict__ y, int C, long long S_in, long long S_out,
                                                        const int32_t* __restrict__ src) {
  const int n = blockIdx.z, c = blockIdx.y;
  const T* xs = x + ((long long)n * C + c) * S_in;
  T* ys = y + ((long long)n * C + c) * S_out;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < S_out; r += (long long)gridDim.x * blockDim.x) {
    const int s = src[r];
    ys[r] = s >= 0 ? xs[s] : T(0.f);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_layernorm_cf(const void* x, int dtype, int N, int C, long long S, const float* gamma, const float* beta, float eps,
                                 void* y, void* stream) {
  B200_REQUIRE(x && y, "layernorm_cf: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && S > 0 && N <= 65535, "layernorm_cf: bad sizes");
  dim3 grid((unsigned)ceil_div(S, 256), N);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F32) layernorm_cf_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, C, S, gamma, beta, eps);
  else if (dtype == B200_DT_F16) layernorm_cf_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, (__half*)y, C, S, gamma, beta, eps);
  else return set_err(B200_ERR_INVALID, "layernorm_cf: bad dtype");
  B200_LAUNCH_CHECK("layernorm_cf_kernel");
  return B200_OK;
}

extern "C" int b200_patchify(const void* x, int dtype, int N, int C, int D, int H, int W, int pd, int ph, int pw, void* y, void* stream) {
  B200_REQUIRE(x && y, "patchify: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && pd > 0 && ph > 0 && pw > 0 && D >= pd && H >= ph && W >= pw && N <= 65535, "patchify: bad sizes");
  const long long total = (long long)C * (D / pd * pd) * (H / ph * ph) * (W / pw * pw);
  dim3 grid((unsigned)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), N);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F32) patchify_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, C, D, H, W, pd, ph, pw);
  else if (dtype == B200_DT_F16) patchify_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, (__half*)y, C, D, H, W, pd, ph, pw);
  else return set_err(B200_ERR_INVALID, "patchify: bad dtype");
  B200_LAUNCH_CHECK("patchify_kernel");
  return B200_OK;
}

extern "C" int b200_mhsa_cf(const void* qkv, int dtype, int N, int heads, int dim_head, long long S, float scale, int win, const float* bias,
                            const int32_t* region, void* out, void* stream) {
  B200_REQUIRE(qkv && out, "mhsa_cf: null pointer");
  B200_REQUIRE(N > 0 && heads > 0 && S > 0 && N <= 65535 && heads <= 65535, "mhsa_cf: bad sizes");
  B200_REQUIRE(dtype == B200_DT_F32 || dtype == B200_DT_F16, "mhsa_cf: bad dtype");
  B200_REQUIRE(win >= 0 && (win == 0 || S % win == 0), "mhsa_cf: the token count %lld is not a multiple of the window %d", S, win);
  B200_REQUIRE(win > 0 || (!bias && !region), "mhsa_cf: bias / region need a window size");
  const long long blocks_x = win > 0 ? (S / win) * ceil_div(win, kMhsaQ) : ceil_div(S, kMhsaQ);
  B200_REQUIRE(blocks_x <= 2147483647LL, "mhsa_cf: too many query tiles");
  dim3 grid((unsigned)blocks_x, heads, N);
  cudaStream_t st = (cudaStream_t)stream;
#define LM(D) do { if (dtype == B200_DT_F32) mhsa_cf_kernel<float, D><<<grid, kMhsaQ, 0, st>>>((const float*)qkv, (float*)out, heads, S, scale, win, bias, region); \
                   else mhsa_cf_kernel<__half, D><<<grid, kMhsaQ, 0, st>>>((const __half*)qkv, (__half*)out, heads, S, scale, win, bias, region); } while (0)
  switch (dim_head) {
    case 8: LM(8); break;
    case 16: LM(16); break;
    case 24: LM(24); break;
    case 32: LM(32); break;
    case 48: LM(48); break;
    case 64: LM(64); break;
    default: return set_err(B200_ERR_UNSUPPORTED, "mhsa_cf: dim_head must be 8, 16, 24, 32, 48 or 64 (got %d)", dim_head);
  }
#undef LM
  B200_LAUNCH_CHECK("mhsa_cf_kernel");
  return B200_OK;
}

extern "C" int b200_gather_cf(const void* x, int dtype, int N, int C, long long S_in, const int32_t* src, long long S_out, void* y, void* stream) {
  B200_REQUIRE(x && y && src, "gather_cf: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && S_in > 0 && S_out > 0 && N <= 65535 && C <= 65535, "gather_cf: bad sizes");
  dim3 grid((unsigned)std::min<long long>(ceil_div(S_out, 256), 1024), C, N);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F32) gather_cf_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, C, S_in, S_out, src);
  else if (dtype == B200_DT_F16) gather_cf_kernel<__half><<<grid, 256, 0, st>>>((const __half*)x, (__half*)y, C, S_in, S_out, src);
  else return set_err(B200_ERR_INVALID, "gather_cf: bad dtype");
  B200_LAUNCH_CHECK("gather_cf_kernel");
  return B200_OK;
}
