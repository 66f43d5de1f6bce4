// SwinUNETR token-side kernels on NC8 activations: LayerNorm + window gather,
// PatchMerging gather + LayerNorm, windowed attention, the single-input-channel stems and the 1x1x1 output head.
// Reference: monai/networks/nets/swin_unetr.py (WindowAttention 426-532, SwinTransformerBlock 535-698,
// PatchMerging 701-773, compute_mask 779-816, proj_out 1040-1053), monai/networks/blocks/patchembedding.py:141-219,
// monai/networks/blocks/dynunet_block.py:247-267.
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

__device__ __forceinline__ void ld8(const __half* p, float (&f)[8]) {
  __align__(16) __half v[8];
  *reinterpret_cast<uint4*>(v) = *reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __half2float(v[j]);
}
__device__ __forceinline__ void st8(__half* p, const float (&f)[8]) {
  __align__(16) __half v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __float2half_rn(f[j]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(v);
}

// ---------------------------------------------------------------------------------------------------- LayerNorm
__global__ void __launch_bounds__(256) layernorm_nc8_kernel(const __half* __restrict__ x, __half* __restrict__ y, int C,
                                                            long long S_in, const int* __restrict__ src, long long S_out,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float eps) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S_out) return;
  const int n = blockIdx.y, C8 = C / 8;
  const long long s = src ? (long long)src[r] : r;
  __half* yo = y + ((long long)n * C8 * S_out + r) * 8;
  if (s < 0) {  // padded token: exact zeros (F.pad after norm1, swin_unetr.py:603-606)
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int c = 0; c < C8; ++c) *reinterpret_cast<uint4*>(yo + (long long)c * S_out * 8) = z;
    return;
  }
  const __half* xi = x + ((long long)n * C8 * S_in + s) * 8;
  float sum = 0.f;
  for (int c = 0; c < C8; ++c) {
    float f[8]; ld8(xi + (long long)c * S_in * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += f[j];
  }
  const float mean = sum / (float)C;
  float var = 0.f;
  for (int c = 0; c < C8; ++c) {
    float f[8]; ld8(xi + (long long)c * S_in * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) { const float d = f[j] - mean; var = fmaf(d, d, var); }
  }
  const float rstd = 1.f / sqrtf(var / (float)C + eps);
  for (int c = 0; c < C8; ++c) {
    float f[8]; ld8(xi + (long long)c * S_in * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[j] = (f[j] - mean) * rstd;
      if (gamma) f[j] = fmaf(f[j], gamma[c * 8 + j], beta ? beta[c * 8 + j] : 0.f);
    }
    st8(yo + (long long)c * S_out * 8, f);
  }
}

// ------------------------------------------------------------------------------------------ PatchMerging gather + LN
__constant__ int kMergeV1[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
__constant__ int kMergeV2[8][3] = {{0, 0, 0}, {0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 0}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};

__global__ void __launch_bounds__(128) patch_merge_ln_nc8_kernel(const __half* __restrict__ x, __half* __restrict__ y, int C, int D,
                                                                 int H, int W, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, int v2) {
  const int D2 = (D + 1) / 2, H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const long long S2 = (long long)D2 * H2 * W2, S = (long long)D * H * W;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S2) return;
  const int n = blockIdx.y, C8 = C / 8;
  const int w2 = (int)(r % W2), h2 = (int)((r / W2) % H2), d2 = (int)(r / ((long long)W2 * H2));
  const __half* xn = x + (long long)n * C8 * S * 8;
  long long off[8]; bool ok[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int* o = v2 ? kMergeV2[q] : kMergeV1[q];
    const int d = 2 * d2 + o[0], h = 2 * h2 + o[1], w = 2 * w2 + o[2];
    ok[q] = d < D && h < H && w < W;
    off[q] = (((long long)d * H + h) * W + w) * 8;
  }
  const float Ct = 8.f * (float)C;
  float sum = 0.f;
  for (int q = 0; q < 8; ++q)
    if (ok[q])
      for (int c = 0; c < C8; ++c) {
        float f[8]; ld8(xn + (long long)c * S * 8 + off[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) sum += f[j];
      }
  const float mean = sum / Ct;
  float var = 0.f;
  for (int q = 0; q < 8; ++q)
    for (int c = 0; c < C8; ++c) {
      float f[8];
      if (ok[q]) ld8(xn + (long long)c * S * 8 + off[q], f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float dd = f[j] - mean; var = fmaf(dd, dd, var); }
    }
  const float rstd = 1.f / sqrtf(var / Ct + eps);
  __half* yo = y + ((long long)n * (8 * C8) * S2 + r) * 8;
  for (int q = 0; q < 8; ++q)
    for (int c = 0; c < C8; ++c) {
      float f[8];
      if (ok[q]) ld8(xn + (long long)c * S * 8 + off[q], f);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
      }
      const int ch = (q * C8 + c) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaf((f[j] - mean) * rstd, gamma ? gamma[ch + j] : 1.f, beta ? beta[ch + j] : 0.f);
      st8(yo + (long long)(q * C8 + c) * S2 * 8, f);
    }
}

// ---------------------------------------------------------------------------------------------- window attention
// one block = one (window, head, batch item); K and V of the head are staged in shared memory as fp32.
__global__ void __launch_bounds__(128) window_attention_nc8_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int C,
                                                                   int nW, int n, float scale, const float* __restrict__ biasT,
                                                                   const int* __restrict__ region) {
  extern __shared__ float s_kv[];  // K[n][16], V[n][16], region[n]
  float* sK = s_kv;
  float* sV = s_kv + (size_t)n * 16;
  int* sR = reinterpret_cast<int*>(sV + (size_t)n * 16);
  const int w = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int C8 = C / 8;
  const long long T = (long long)nW * n;
  const __half* base = qkv + (long long)b * (3 * C8) * T * 8;
  const long long row0 = (long long)w * n;
  for (int i = threadIdx.x; i < n * 2; i += blockDim.x) {
    const int t = i >> 1, half_ = i & 1;
    float f[8];
    ld8(base + ((long long)(C8 + 2 * h + half_) * T + row0 + t) * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sK[t * 16 + half_ * 8 + j] = f[j];
    ld8(base + ((long long)(2 * C8 + 2 * h + half_) * T + row0 + t) * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sV[t * 16 + half_ * 8 + j] = f[j];
  }
  if (region)
    for (int i = threadIdx.x; i < n; i += blockDim.x) sR[i] = region[(long long)w * n + i];
  __syncthreads();
  const float* bh = biasT + (long long)h * n * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float q[16], o[16];
    {
      float f[8];
      ld8(base + ((long long)(2 * h) * T + row0 + i) * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[j] = f[j] * scale;
      ld8(base + ((long long)(2 * h + 1) * T + row0 + i) * 8, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) q[8 + j] = f[j] * scale;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) o[j] = 0.f;
    float m = -INFINITY, l = 0.f;
    const int ri = region ? sR[i] : 0;
    for (int j = 0; j < n; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(sK + j * 16);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 kv = kp[t];
        s = fmaf(q[4 * t], kv.x, s); s = fmaf(q[4 * t + 1], kv.y, s); s = fmaf(q[4 * t + 2], kv.z, s); s = fmaf(q[4 * t + 3], kv.w, s);
      }
      s += bh[(long long)j * n + i];
      if (region && sR[j] != ri) s += -100.0f;
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), pj = __expf(s - mn);
      l = l * corr + pj;
      const float4* vp = reinterpret_cast<const float4*>(sV + j * 16);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float4 vv = vp[t];
        o[4 * t] = fmaf(o[4 * t], corr, pj * vv.x); o[4 * t + 1] = fmaf(o[4 * t + 1], corr, pj * vv.y);
        o[4 * t + 2] = fmaf(o[4 * t + 2], corr, pj * vv.z); o[4 * t + 3] = fmaf(o[4 * t + 3], corr, pj * vv.w);
      }
      m = mn;
    }
    const float inv = 1.f / l;
    float f0[8], f1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f0[j] = o[j] * inv; f1[j] = o[8 + j] * inv; }
    __half* ob = out + (long long)b * C8 * T * 8;
    st8(ob + ((long long)(2 * h) * T + row0 + i) * 8, f0);
    st8(ob + ((long long)(2 * h + 1) * T + row0 + i) * 8, f1);
  }
}

// -------------------------------------------------------------------------------------- single-input-channel convs
template <typename T>
__global__ void __launch_bounds__(128) conv_cin1_nc8_kernel(const T* __restrict__ x, __half* __restrict__ y, const float* __restrict__ wgt,
                                                            const float* __restrict__ bias, int D, int H, int W, int Do, int Ho, int Wo,
                                                            int Cout, int k, int stride, int pad, int out_ctot, int out_coff,
                                                            float* __restrict__ stats) {
  extern __shared__ float s_w[];  // [taps][Cout], then stats [2*Cout]
  const int taps = k * k * k;
  float* s_st = s_w + taps * Cout;
  for (int i = threadIdx.x; i < taps * Cout; i += blockDim.x) {
    const int co = i % Cout, t = i / Cout;
    s_w[i] = wgt[(long long)co * taps + t];
  }
  for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) s_st[i] = 0.f;
  __syncthreads();
  const int n = blockIdx.y;
  const long long So = (long long)Do * Ho * Wo;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = r < So;
  float xv[27];
  if (ok) {
    const int ox = (int)(r % Wo), oy = (int)((r / Wo) % Ho), oz = (int)(r / ((long long)Wo * Ho));
    const T* xn = x + (long long)n * D * H * W;
    for (int t = 0; t < taps; ++t) {
      const int kz = t / (k * k), ky = (t / k) % k, kx = t % k;
      const int iz = oz * stride - pad + kz, iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      xv[t] = (iz >= 0 && iz < D && iy >= 0 && iy < H && ix >= 0 && ix < W) ? io<T>::ld(xn + ((long long)iz * H + iy) * W + ix) : 0.f;
    }
  }
  const int lane = threadIdx.x & 31;
  __half* yo = y + (((long long)n * (out_ctot / 8) + out_coff / 8) * So + r) * 8;
  for (int c0 = 0; c0 < Cout; c0 += 8) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = bias ? bias[c0 + j] : 0.f;
    if (ok)
      for (int t = 0; t < taps; ++t) {
        const float xt = xv[t];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xt, s_w[t * Cout + c0 + j], acc[j]);
      }
    if (ok) st8(yo + (long long)(c0 / 8) * So * 8, acc);
    if (stats) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a = ok ? acc[j] : 0.f;
        const float s1 = warp_sum(a), s2 = warp_sum(a * a);
        if (lane == 0) { atomicAdd(&s_st[2 * (c0 + j)], s1); atomicAdd(&s_st[2 * (c0 + j) + 1], s2); }
      }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) atomicAdd(&stats[(long long)n * Cout * 2 + i], s_st[i]);
  }
}

template <typename TO>
__global__ void __launch_bounds__(256) head_conv_nc8_kernel(const __half* __restrict__ x, TO* __restrict__ y, const float* __restrict__ wgt,
                                                            const float* __restrict__ bias, int C, long long S, int Cout) {
  extern __shared__ float s_hw[];  // [Cout][C]
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) s_hw[i] = wgt[i];
  __syncthreads();
  const int n = blockIdx.y;
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= S) return;
  float acc[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc[o] = (o < Cout && bias) ? bias[o] : 0.f;
  const __half* xi = x + ((long long)n * (C / 8) * S + r) * 8;
  for (int c = 0; c < C / 8; ++c) {
    float f[8]; ld8(xi + (long long)c * S * 8, f);
#pragma unroll
    for (int o = 0; o < 16; ++o)
      if (o < Cout) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[o] = fmaf(f[j], s_hw[o * C + c * 8 + j], acc[o]);
      }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o)
    if (o < Cout) io<TO>::st(y + ((long long)n * Cout + o) * S + r, acc[o]);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_layernorm_nc8(const void* x, int N, int C, long long S_in, const int32_t* src, long long S_out,
                                  const float* gamma, const float* beta, float eps, void* y, void* stream) {
  B200_REQUIRE(x && y, "layernorm_nc8: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && S_in > 0 && S_out > 0, "layernorm_nc8: bad sizes");
  dim3 grid(ceil_div(S_out, 256), N);
  layernorm_nc8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, C, S_in, src, S_out, gamma, beta, eps);
  B200_LAUNCH_CHECK("layernorm_nc8_kernel");
  return B200_OK;
}

extern "C" int b200_patch_merge_ln_nc8(const void* x, int N, int C, int D, int H, int W, const float* gamma,
                                       const float* beta, float eps, int v2, void* y, void* stream) {
  B200_REQUIRE(x && y, "patch_merge_ln_nc8: null pointer");
  B200_REQUIRE(N > 0 && C > 0 && C % 8 == 0 && D > 0 && H > 0 && W > 0, "patch_merge_ln_nc8: bad sizes");
  const long long S2 = (long long)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
  dim3 grid(ceil_div(S2, 128), N);
  patch_merge_ln_nc8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const __half*)x, (__half*)y, C, D, H, W, gamma, beta, eps, v2);
  B200_LAUNCH_CHECK("patch_merge_ln_nc8_kernel");
  return B200_OK;
}

extern "C" int b200_window_attention_nc8(const void* qkv, int N, int C, int heads, int nW, int n, float scale,
                                         const float* bias, const int32_t* region, void* out, void* stream) {
  B200_REQUIRE(qkv && out && bias, "window_attention_nc8: null pointer");
  B200_REQUIRE(N > 0 && heads > 0 && nW > 0 && n > 0, "window_attention_nc8: empty problem");
  B200_REQUIRE(C == heads * 16, "window_attention_nc8: head_dim must be 16 (C = %d, heads = %d)", C, heads);
  B200_REQUIRE(nW <= 2147483647 / n && heads <= 65535 && N <= 65535, "window_attention_nc8: grid too large");
  const size_t smem = (size_t)n * 32 * sizeof(float) + (size_t)n * sizeof(int);
  B200_REQUIRE(smem <= 160 * 1024, "window_attention_nc8: window of %d tokens does not fit in shared memory", n);
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(window_attention_nc8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  dim3 grid(nW, heads, N);
  window_attention_nc8_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const __half*)qkv, (__half*)out, C, nW, n, scale, bias, region);
  B200_LAUNCH_CHECK("window_attention_nc8_kernel");
  return B200_OK;
}

extern "C" int b200_conv_cin1_nc8(const void* x, int dtype, int N, int D, int H, int W, const float* weight, const float* bias,
                                  int Cout, int k, int stride, int pad, void* y, int out_ctot, int out_coff, float* stats,
                                  void* stream) {
  B200_REQUIRE(x && y && weight, "conv_cin1_nc8: null pointer");
  B200_REQUIRE(k >= 1 && k <= 3 && stride >= 1 && pad >= 0, "conv_cin1_nc8: kernel size must be 1..3");
  B200_REQUIRE(Cout > 0 && Cout % 8 == 0 && out_ctot % 8 == 0 && out_coff % 8 == 0 && out_coff + Cout <= out_ctot,
               "conv_cin1_nc8: channel counts must be multiples of 8");
  const int Do = (D + 2 * pad - k) / stride + 1, Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  B200_REQUIRE(Do > 0 && Ho > 0 && Wo > 0, "conv_cin1_nc8: empty output");
  const long long So = (long long)Do * Ho * Wo;
  dim3 grid(ceil_div(So, 128), N);
  const size_t smem = ((size_t)k * k * k * Cout + 2 * Cout) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F16)
    conv_cin1_nc8_kernel<__half><<<grid, 128, smem, st>>>((const __half*)x, (__half*)y, weight, bias, D, H, W, Do, Ho, Wo, Cout, k, stride, pad, out_ctot, out_coff, stats);
  else if (dtype == B200_DT_F32)
    conv_cin1_nc8_kernel<float><<<grid, 128, smem, st>>>((const float*)x, (__half*)y, weight, bias, D, H, W, Do, Ho, Wo, Cout, k, stride, pad, out_ctot, out_coff, stats);
  else return set_err(B200_ERR_INVALID, "conv_cin1_nc8: bad dtype");
  B200_LAUNCH_CHECK("conv_cin1_nc8_kernel");
  return B200_OK;
}

extern "C" int b200_head_conv_nc8(const void* x, int N, int C, long long S, const float* weight, const float* bias, int Cout,
                                  void* y, int out_dtype, void* stream) {
  B200_REQUIRE(x && y && weight, "head_conv_nc8: null pointer");
  B200_REQUIRE(C % 8 == 0 && Cout >= 1 && Cout <= 16, "head_conv_nc8: C must be a multiple of 8 and Cout <= 16 (got %d, %d)", C, Cout);
  dim3 grid(ceil_div(S, 256), N);
  const size_t smem = (size_t)Cout * C * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == B200_DT_F16) head_conv_nc8_kernel<__half><<<grid, 256, smem, st>>>((const __half*)x, (__half*)y, weight, bias, C, S, Cout);
  else if (out_dtype == B200_DT_F32) head_conv_nc8_kernel<float><<<grid, 256, smem, st>>>((const __half*)x, (float*)y, weight, bias, C, S, Cout);
  else return set_err(B200_ERR_INVALID, "head_conv_nc8: bad dtype");
  B200_LAUNCH_CHECK("head_conv_nc8_kernel");
  return B200_OK;
}
