// Affine-driven 3-D resampling and zero-padded separable filtering.
//
// Resample: the reference materialises a dense coordinate grid (F.affine_grid or create_grid @ affine) and
// calls F.grid_sample (monai/networks/layers/spatial_transforms.py:584-591, monai/transforms/spatial/array.py:2102-2115).
// Coordinates are an affine function of the output index, so this kernel evaluates them on the fly in fp64
// (the reference's default coordinate dtype, spatial/array.py:355,1972) and never builds the grid.
// Sampling semantics follow ATen grid_sampler_3d (unnormalised coordinates are produced by the host):
//   padding zeros / border / reflection (reflection bounds depend on align_corners), nearest = round-half-even.
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

struct ResampleP {
  const void* src; void* dst;
  int C, Di, Hi, Wi, Do, Ho, Wo;
  double m[12];
  int interp, pad, align;
};

__device__ __forceinline__ double reflect_coord(double in, double twice_low, double twice_high) {
  if (twice_low == twice_high) return 0.0;
  const double mn = twice_low / 2.0, span = (twice_high - twice_low) / 2.0;
  in = fabs(in - mn);
  const double extra = fmod(in, span);
  const int flips = (int)floor(in / span);
  return (flips & 1) ? span - extra + mn : extra + mn;
}

__device__ __forceinline__ double pad_coord(double x, int size, int pad, int align) {
  if (pad == 1) {
    x = fmin((double)(size - 1), fmax(x, 0.0));
  } else if (pad == 2) {
    x = align ? reflect_coord(x, 0.0, 2.0 * (size - 1)) : reflect_coord(x, -1.0, 2.0 * size - 1.0);
    x = fmin((double)(size - 1), fmax(x, 0.0));
  }
  return x;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) resample_affine_kernel(ResampleP p) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;  // fastest output axis
  const int j = blockIdx.y, i = blockIdx.z;
  if (k >= p.Wo) return;
  // input coordinates (a, b, c) along (D, H, W) of the source
  double a = fma(p.m[0], (double)i, fma(p.m[1], (double)j, fma(p.m[2], (double)k, p.m[3])));
  double b = fma(p.m[4], (double)i, fma(p.m[5], (double)j, fma(p.m[6], (double)k, p.m[7])));
  double c = fma(p.m[8], (double)i, fma(p.m[9], (double)j, fma(p.m[10], (double)k, p.m[11])));
  a = pad_coord(a, p.Di, p.pad, p.align);
  b = pad_coord(b, p.Hi, p.pad, p.align);
  c = pad_coord(c, p.Wi, p.pad, p.align);
  const long long in_cs = (long long)p.Di * p.Hi * p.Wi, out_cs = (long long)p.Do * p.Ho * p.Wo;
  const long long o = ((long long)i * p.Ho + j) * p.Wo + k;
  const TI* src = (const TI*)p.src;
  TO* dst = (TO*)p.dst;
  if (p.interp == 0) {
    const int ia = (int)nearbyint(a), ib = (int)nearbyint(b), ic = (int)nearbyint(c);
    const bool ok = ia >= 0 && ia < p.Di && ib >= 0 && ib < p.Hi && ic >= 0 && ic < p.Wi;
    const long long off = ((long long)ia * p.Hi + ib) * p.Wi + ic;
    for (int ch = 0; ch < p.C; ++ch) io<TO>::st(dst + ch * out_cs + o, ok ? io<TI>::ld(src + ch * in_cs + off) : 0.f);
    return;
  }
  const double fa = floor(a), fb = floor(b), fc = floor(c);
  const int a0 = (int)fa, b0 = (int)fb, c0 = (int)fc;
  const float ta = (float)(a - fa), tb = (float)(b - fb), tc = (float)(c - fc);
  const float wa[2] = {1.f - ta, ta}, wb[2] = {1.f - tb, tb}, wc[2] = {1.f - tc, tc};
  float wgt[8]; long long off[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int da = q >> 2, db = (q >> 1) & 1, dc = q & 1;
    const int ia = a0 + da, ib = b0 + db, ic = c0 + dc;
    const bool ok = ia >= 0 && ia < p.Di && ib >= 0 && ib < p.Hi && ic >= 0 && ic < p.Wi;
    wgt[q] = ok ? wa[da] * wb[db] * wc[dc] : 0.f;
    off[q] = ok ? ((long long)ia * p.Hi + ib) * p.Wi + ic : 0;
  }
  for (int ch = 0; ch < p.C; ++ch) {
    const TI* s = src + ch * in_cs;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v = fmaf(io<TI>::ld(s + off[q]), wgt[q], v);
    io<TO>::st(dst + ch * out_cs + o, v);
  }
}

// out[i] = sum_t taps[t] * in[i + (t - r) * stride] along one axis, zero outside.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) filter1d_kernel(const TI* __restrict__ in, TO* __restrict__ out,
                                                       const float* __restrict__ taps, int n, long long total,
                                                       long long stride, int extent) {
  __shared__ float s_t[128];
  for (int t = threadIdx.x; t < n; t += blockDim.x) s_t[t] = taps[t];
  __syncthreads();
  const int r = (n - 1) / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pos = (int)((i / stride) % extent);
    float acc = 0.f;
    for (int t = 0; t < n; ++t) {
      const int q = pos + t - r;
      if (q >= 0 && q < extent) acc = fmaf(s_t[t], io<TI>::ld(in + i + (long long)(t - r) * stride), acc);
    }
    io<TO>::st(out + i, acc);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_resample_affine(const void* src, int src_dtype, int C, int Di, int Hi, int Wi, void* dst,
                                    int dst_dtype, int Do, int Ho, int Wo, const double* mat3x4, int interp, int pad,
                                    int align_corners, void* stream) {
  B200_REQUIRE(src && dst && mat3x4, "resample_affine: null pointer");
  B200_REQUIRE(C > 0 && Di > 0 && Hi > 0 && Wi > 0, "resample_affine: empty source");
  if ((long long)Do * Ho * Wo == 0) return B200_OK;
  B200_REQUIRE(interp == 0 || interp == 1, "resample_affine: interp must be 0 (nearest) or 1 (trilinear)");
  B200_REQUIRE(pad >= 0 && pad <= 2, "resample_affine: pad must be 0 (zeros), 1 (border) or 2 (reflection)");
  B200_REQUIRE(Do <= 65535 && Ho <= 65535, "resample_affine: output too large for the launch grid");
  ResampleP p;
  p.src = src; p.dst = dst; p.C = C; p.Di = Di; p.Hi = Hi; p.Wi = Wi; p.Do = Do; p.Ho = Ho; p.Wo = Wo;
  for (int q = 0; q < 12; ++q) p.m[q] = mat3x4[q];
  p.interp = interp; p.pad = pad; p.align = align_corners;
  dim3 block(Wo >= 192 ? 256 : (Wo >= 96 ? 128 : 64)), grid(ceil_div(Wo, block.x), Ho, Do);
  cudaStream_t st = (cudaStream_t)stream;
#define LR(TI, TO) resample_affine_kernel<TI, TO><<<grid, block, 0, st>>>(p)
  if (src_dtype == B200_DT_F32 && dst_dtype == B200_DT_F32) LR(float, float);
  else if (src_dtype == B200_DT_F16 && dst_dtype == B200_DT_F32) LR(__half, float);
  else if (src_dtype == B200_DT_F32 && dst_dtype == B200_DT_F16) LR(float, __half);
  else if (src_dtype == B200_DT_F16 && dst_dtype == B200_DT_F16) LR(__half, __half);
  else return set_err(B200_ERR_INVALID, "resample_affine: bad dtype");
#undef LR
  B200_LAUNCH_CHECK("resample_affine_kernel");
  return B200_OK;
}

extern "C" int b200_separable_filter3d(const void* src, int dtype, int C, int D, int H, int W, const float* taps_d,
                                       int n_d, const float* taps_h, int n_h, const float* taps_w, int n_w, float* tmp,
                                       void* dst, void* stream) {
  B200_REQUIRE(src && dst && tmp, "separable_filter3d: null pointer");
  B200_REQUIRE(n_d <= 127 && n_h <= 127 && n_w <= 127, "separable_filter3d: more than 127 taps");
  B200_REQUIRE(n_d % 2 && n_h % 2 && n_w % 2, "separable_filter3d: tap counts must be odd");
  const long long total = (long long)C * D * H * W;
  if (total == 0) return B200_OK;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 32);
  cudaStream_t st = (cudaStream_t)stream;
  // reference order (simplelayers.py:170-204): first spatial axis first, last spatial axis last.
  // src -(D)-> tmp[0:total] -(H)-> tmp[total:2*total] -(W)-> dst; intermediates stay fp32.
  const bool f16 = dtype == B200_DT_F16;
  B200_REQUIRE(f16 || dtype == B200_DT_F32, "separable_filter3d: bad dtype");
  B200_REQUIRE(taps_d && taps_h && taps_w, "separable_filter3d: all three axis kernels are required");
  float* tmp2 = tmp + total;
  if (f16) filter1d_kernel<__half, float><<<blocks, 256, 0, st>>>((const __half*)src, tmp, taps_d, n_d, total, (long long)H * W, D);
  else filter1d_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, tmp, taps_d, n_d, total, (long long)H * W, D);
  B200_LAUNCH_CHECK("filter1d_kernel(d)");
  filter1d_kernel<float, float><<<blocks, 256, 0, st>>>(tmp, tmp2, taps_h, n_h, total, (long long)W, H);
  B200_LAUNCH_CHECK("filter1d_kernel(h)");
  if (f16) filter1d_kernel<float, __half><<<blocks, 256, 0, st>>>(tmp2, (__half*)dst, taps_w, n_w, total, 1, W);
  else filter1d_kernel<float, float><<<blocks, 256, 0, st>>>(tmp2, (float*)dst, taps_w, n_w, total, 1, W);
  B200_LAUNCH_CHECK("filter1d_kernel(w)");
  return B200_OK;
}
