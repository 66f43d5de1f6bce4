// Dense-grid spline resampling ("pull") with the boundary conditions of the discrete transforms.
//
// Replaces monai._C.grid_pull (monai/csrc/ext.cpp:66-74 -> monai/csrc/resample/pushpull.h:58-110, pushpull_cuda.cu:2140-2205;
// python wrapper monai/networks/layers/spatial_transforms.py:35-132) and serves the dense-grid form of Resample.__call__
// (monai/transforms/spatial/array.py:2015-2117):
//     out[b, c, o] = sum_k  sign_k * src[b, c, wrap(idx_k)] * w_k(x(o))
// where x(o) = scale * grid[b, :, o] + shift is a VOXEL coordinate per axis (the per-axis affine lets the caller hand over the
// reference's centred / normalised grids without a pass that rewrites them), w are the centred cardinal B-spline weights of
// order 0..7 per axis (no prefilter, as the reference), and (wrap, sign) implement the seven boundary conditions of
// monai/csrc/resample/bounds_common.h: replicate, dct1, dct2, dst1, dst2, dft, zero.
//
// One thread per output voxel, all channels.  Coordinates and weights are evaluated in the grid's precision (float64 grids --
// the reference's default coordinate dtype -- keep float64 coordinates); accumulation in fp32.  This is an HBM-bound gather
// kernel (grid read + source footprint + output write); orders 0 and 1 use a two-tap instantiation without local arrays.
#include "common.cuh"
#include "../../include/monai_b200.h"
#include <type_traits>

namespace b200 {

enum { kBoundReplicate = 0, kBoundDct1 = 1, kBoundDct2 = 2, kBoundDst1 = 3, kBoundDst2 = 4, kBoundDft = 5, kBoundZero = 7 };

struct GridPullP {
  const void* src; const void* grid; void* out;
  int B, C, X, Y, Z, Xo, Yo, Zo;
  long long g_sb, g_sc, g_sv;     // grid element strides: batch, component, voxel
  double scale[3], shift[3];      // voxel coordinate = scale * grid value + shift
  int bound[3], order[3];
  int extrapolate, half_even;     // order 0: round half to even (ATen) instead of half away from zero (monai._C)
};

// ---- boundary conditions (index wrap + sign), after monai/csrc/resample/bounds_common.h ------------------------------
__device__ __forceinline__ int bound_index(int type, int c, int n) {
  switch (type) {
    case kBoundReplicate: return c <= 0 ? 0 : (c >= n ? n - 1 : c);
    case kBoundDct1: {                       // reflect about the centre of the border voxels, period 2(n-1)
      if (n == 1) return 0;
      const int p = 2 * (n - 1);
      c = c < 0 ? -c : c;
      c %= p;
      return c >= n ? p - c : c;
    }
    case kBoundDct2: case kBoundDst2: {      // reflect about the edge of the border voxels, period 2n
      const int p = 2 * n;
      c = c < 0 ? p - ((-c - 1) % p) - 1 : c % p;
      return c >= n ? p - c - 1 : c;
    }
    case kBoundDst1: {                       // antisymmetric about the first out-of-bound voxel, period 2(n+1)
      if (n == 1) return 0;
      const int p = 2 * (n + 1);
      c = c == -1 ? 0 : (c < 0 ? -c - 2 : c);
      c %= p;
      return c == n ? n - 1 : (c > n ? p - c - 2 : c);
    }
    case kBoundDft: return c < 0 ? (n + c % n) % n : c % n;
    default: return c;                       // zero: the sign is 0 outside, the index is never dereferenced
  }
}
__device__ __forceinline__ int bound_sign(int type, int c, int n) {
  switch (type) {
    case kBoundDst1: {
      if (n == 1) return 1;
      const int p = 2 * (n + 1);
      c = c < 0 ? n - c - 1 : c;
      c %= p;
      if (c % (n + 1) == n) return 0;
      return ((c / (n + 1)) % 2) ? -1 : 1;
    }
    case kBoundDst2: {
      c = c < 0 ? n - c - 1 : c;
      return ((c / n) % 2) ? -1 : 1;
    }
    case kBoundZero: return (c < 0 || c >= n) ? 0 : 1;
    default: return 1;
  }
}

// centred cardinal B-spline of order n at t:  1/n! sum_j (-1)^j C(n+1, j) (t + (n+1)/2 - j)_+^n   (evaluated in double)
__device__ double bspline(int n, double t) {
  if (n == 0) return fabs(t) < 0.5 ? 1.0 : 0.0;
  if (n == 1) { const double a = fabs(t); return a < 1.0 ? 1.0 - a : 0.0; }
  const double h = 0.5 * (n + 1);
  if (fabs(t) >= h) return 0.0;
  double sum = 0.0, binom = 1.0, fact = 1.0;
  for (int k = 2; k <= n; ++k) fact *= k;
  for (int j = 0; j <= n + 1; ++j) {
    const double u = t + h - j;
    if (u > 0.0) {
      double pw = 1.0;
      for (int k = 0; k < n; ++k) pw *= u;
      sum += ((j & 1) ? -binom : binom) * pw;
    }
    binom = binom * (n + 1 - j) / (j + 1);
  }
  return sum / fact;
}

template <typename T> __device__ __forceinline__ double ld_coord(const T* p);
template <> __device__ __forceinline__ double ld_coord<float>(const float* p) { return (double)__ldg(p); }
template <> __device__ __forceinline__ double ld_coord<double>(const double* p) { return __ldg(p); }
template <> __device__ __forceinline__ double ld_coord<__half>(const __half* p) { return (double)__half2float(__ldg(p)); }

// taps of one axis: first index, count, weights, wrapped indices and signs
template <int MAXT>
struct AxisTaps { int n; float w[MAXT]; int idx[MAXT]; int sgn[MAXT]; };

template <int MAXT, typename TC>
__device__ __forceinline__ void axis_taps(TC x, int order, int bound, int size, int half_even, AxisTaps<MAXT>& a) {
  int low;
  if (order == 0) {
    low = half_even ? (int)nearbyint((double)x) : (int)round((double)x);
    a.n = 1;
    a.w[0] = 1.f;
  } else {
    low = (int)floor((double)x - 0.5 * (order - 1));
    a.n = order + 1;
    if (MAXT == 2 || order == 1) {
      const TC t = x - (TC)low;          // the reference's dx1 = x - ix0, dx0 = 1 - dx1 (pushpull_cpu.cpp:1489-1495)
      a.w[1] = (float)t;
      a.w[0] = (float)((TC)1 - t);
    } else {
#pragma unroll 1
      for (int k = 0; k < MAXT; ++k)
        if (k < a.n) a.w[k] = (float)bspline(order, (double)x - (double)(low + k));
    }
  }
#pragma unroll
  for (int k = 0; k < MAXT; ++k) {
    if (k < a.n) {
      a.sgn[k] = bound_sign(bound, low + k, size);       // sign before wrapping (bounds_common.h)
      a.idx[k] = bound_index(bound, low + k, size);
      if (a.sgn[k] == 0) a.idx[k] = 0;                   // never dereferenced with a non-zero weight; keep the address valid
    }
  }
}

template <typename TS, typename TG, typename TO, int MAXT>
__global__ void __launch_bounds__(128) grid_pull_kernel(GridPullP p) {
  using TC = typename std::conditional<std::is_same<TG, double>::value, double, float>::type;   // coordinate precision
  const long long Vo = (long long)p.Xo * p.Yo * p.Zo;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= Vo) return;
  const TG* g = (const TG*)p.grid + (long long)b * p.g_sb + v * p.g_sv;
  const TC x = (TC)fma(p.scale[0], ld_coord<TG>(g), p.shift[0]);
  const TC y = (TC)fma(p.scale[1], ld_coord<TG>(g + p.g_sc), p.shift[1]);
  const TC z = (TC)fma(p.scale[2], ld_coord<TG>(g + 2 * p.g_sc), p.shift[2]);
  const long long Vs = (long long)p.X * p.Y * p.Z;
  TO* out = (TO*)p.out + (long long)b * p.C * Vo + v;
  const TC tiny = (TC)5e-2;   // the reference's inbounds() tolerance (pushpull_cpu.cpp:67, bounds_common.h:170-173)
  const bool inb = x >= -tiny && x < (TC)(p.X - 1) + tiny && y >= -tiny && y < (TC)(p.Y - 1) + tiny && z >= -tiny && z < (TC)(p.Z - 1) + tiny;
  if (!p.extrapolate && !inb) {
    for (int c = 0; c < p.C; ++c) io<TO>::st(out + (long long)c * Vo, 0.f);
    return;
  }
  AxisTaps<MAXT> ax, ay, az;
  axis_taps<MAXT, TC>(x, p.order[0], p.bound[0], p.X, p.half_even, ax);
  axis_taps<MAXT, TC>(y, p.order[1], p.bound[1], p.Y, p.half_even, ay);
  axis_taps<MAXT, TC>(z, p.order[2], p.bound[2], p.Z, p.half_even, az);
  const TS* src = (const TS*)p.src + (long long)b * p.C * Vs;
  for (int c = 0; c < p.C; ++c) {
    const TS* s = src + (long long)c * Vs;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < MAXT; ++i) {
      if (i >= ax.n) break;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        if (j >= ay.n) break;
        const float wxy = ax.w[i] * ay.w[j];
        const int sxy = ax.sgn[i] * ay.sgn[j];
        const long long oxy = ((long long)ax.idx[i] * p.Y + ay.idx[j]) * p.Z;
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
          if (k >= az.n) break;
          const int sg = sxy * az.sgn[k];
          if (sg == 0) continue;
          const float val = io<TS>::ld(s + oxy + az.idx[k]);
          acc = fmaf(sg < 0 ? -val : val, wxy * az.w[k], acc);
        }
      }
    }
    io<TO>::st(out + (long long)c * Vo, acc);
  }
}

// The adjoint ("push" / splat; monai._C.grid_push and grid_count, pushpull.h:112-216): every voxel v of the input, whose
// coordinate is grid[b, :, v], adds  sign_k * w_k * in[b, c, v]  to its taps in out[b, c, X, Y, Z].  p.src is the input
// [B, C, Xo*Yo*Zo] (null for grid_count: the value is 1 and C = 1), p.out the zero-initialised fp32 volume.  Scattered
// float atomics: the sum order is not fixed, so results agree with the reference to rounding, not bit for bit.
template <typename TS, typename TG, int MAXT>
__global__ void __launch_bounds__(128) grid_push_kernel(GridPullP p) {
  using TC = typename std::conditional<std::is_same<TG, double>::value, double, float>::type;
  const long long Vi = (long long)p.Xo * p.Yo * p.Zo;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= Vi) return;
  const TG* g = (const TG*)p.grid + (long long)b * p.g_sb + v * p.g_sv;
  const TC x = (TC)fma(p.scale[0], ld_coord<TG>(g), p.shift[0]);
  const TC y = (TC)fma(p.scale[1], ld_coord<TG>(g + p.g_sc), p.shift[1]);
  const TC z = (TC)fma(p.scale[2], ld_coord<TG>(g + 2 * p.g_sc), p.shift[2]);
  const TC tiny = (TC)5e-2;
  const bool inb = x >= -tiny && x < (TC)(p.X - 1) + tiny && y >= -tiny && y < (TC)(p.Y - 1) + tiny && z >= -tiny && z < (TC)(p.Z - 1) + tiny;
  if (!p.extrapolate && !inb) return;
  AxisTaps<MAXT> ax, ay, az;
  axis_taps<MAXT, TC>(x, p.order[0], p.bound[0], p.X, p.half_even, ax);
  axis_taps<MAXT, TC>(y, p.order[1], p.bound[1], p.Y, p.half_even, ay);
  axis_taps<MAXT, TC>(z, p.order[2], p.bound[2], p.Z, p.half_even, az);
  const long long Vs = (long long)p.X * p.Y * p.Z;
  float* out = (float*)p.out + (long long)b * p.C * Vs;
  const TS* in = p.src ? (const TS*)p.src + (long long)b * p.C * Vi + v : nullptr;
  for (int c = 0; c < p.C; ++c) {
    const float val = in ? io<TS>::ld(in + (long long)c * Vi) : 1.f;
    float* o = out + (long long)c * Vs;
#pragma unroll
    for (int i = 0; i < MAXT; ++i) {
      if (i >= ax.n) break;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        if (j >= ay.n) break;
        const float wxy = ax.w[i] * ay.w[j];
        const int sxy = ax.sgn[i] * ay.sgn[j];
        const long long oxy = ((long long)ax.idx[i] * p.Y + ay.idx[j]) * p.Z;
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
          if (k >= az.n) break;
          const int sg = sxy * az.sgn[k];
          if (sg == 0) continue;
          const float w = wxy * az.w[k];
          atomicAdd(o + oxy + az.idx[k], (sg < 0 ? -val : val) * w);
        }
      }
    }
  }
}

// derivatives of the tap weights with respect to the coordinate (interpolation_common.h fastgrad0..7):
// B_n'(t) = B_{n-1}(t + 1/2) - B_{n-1}(t - 1/2); order 0 -> 0; order 1 -> -1, +1
template <int MAXT, typename TC>
__device__ __forceinline__ void axis_grads(TC x, int order, float (&g)[MAXT]) {
  if (order == 0) { g[0] = 0.f; return; }
  if (MAXT == 2 || order == 1) { g[0] = -1.f; g[1] = 1.f; return; }
  const int low = (int)floor((double)x - 0.5 * (order - 1));
#pragma unroll 1
  for (int k = 0; k < MAXT; ++k)
    if (k <= order) {
      const double t = (double)x - (double)(low + k);
      g[k] = (float)(bspline(order - 1, t + 0.5) - bspline(order - 1, t - 0.5));
    }
}

// Spatial gradients of the interpolated volume (monai._C.grid_grad; the do_sgrad branches of pushpull_cpu.cpp:1070-1080):
// out[b, c, o, d] = sum_k sign_k * src[b, c, idx_k] * dw_k/dx_d.  One thread per output voxel; out [B, C, Vo, 3].
template <typename TS, typename TG, typename TO, int MAXT>
__global__ void __launch_bounds__(128) grid_grad_kernel(GridPullP p) {
  using TC = typename std::conditional<std::is_same<TG, double>::value, double, float>::type;
  const long long Vo = (long long)p.Xo * p.Yo * p.Zo;
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (v >= Vo) return;
  const TG* g = (const TG*)p.grid + (long long)b * p.g_sb + v * p.g_sv;
  const TC x = (TC)fma(p.scale[0], ld_coord<TG>(g), p.shift[0]);
  const TC y = (TC)fma(p.scale[1], ld_coord<TG>(g + p.g_sc), p.shift[1]);
  const TC z = (TC)fma(p.scale[2], ld_coord<TG>(g + 2 * p.g_sc), p.shift[2]);
  const long long Vs = (long long)p.X * p.Y * p.Z;
  TO* out = (TO*)p.out + ((long long)b * p.C * Vo + v) * 3;
  const TC tiny = (TC)5e-2;
  const bool inb = x >= -tiny && x < (TC)(p.X - 1) + tiny && y >= -tiny && y < (TC)(p.Y - 1) + tiny && z >= -tiny && z < (TC)(p.Z - 1) + tiny;
  if (!p.extrapolate && !inb) {
    for (int c = 0; c < p.C; ++c)
      for (int d = 0; d < 3; ++d) io<TO>::st(out + (long long)c * Vo * 3 + d, 0.f);
    return;
  }
  AxisTaps<MAXT> ax, ay, az;
  float gx[MAXT], gy[MAXT], gz[MAXT];
  axis_taps<MAXT, TC>(x, p.order[0], p.bound[0], p.X, p.half_even, ax);
  axis_taps<MAXT, TC>(y, p.order[1], p.bound[1], p.Y, p.half_even, ay);
  axis_taps<MAXT, TC>(z, p.order[2], p.bound[2], p.Z, p.half_even, az);
  axis_grads<MAXT, TC>(x, p.order[0], gx);
  axis_grads<MAXT, TC>(y, p.order[1], gy);
  axis_grads<MAXT, TC>(z, p.order[2], gz);
  const TS* src = (const TS*)p.src + (long long)b * p.C * Vs;
  for (int c = 0; c < p.C; ++c) {
    const TS* s = src + (long long)c * Vs;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int i = 0; i < MAXT; ++i) {
      if (i >= ax.n) break;
#pragma unroll
      for (int j = 0; j < MAXT; ++j) {
        if (j >= ay.n) break;
        const float gw = gx[i] * ay.w[j], wg = ax.w[i] * gy[j], ww = ax.w[i] * ay.w[j];
        const int sxy = ax.sgn[i] * ay.sgn[j];
        const long long oxy = ((long long)ax.idx[i] * p.Y + ay.idx[j]) * p.Z;
#pragma unroll
        for (int k = 0; k < MAXT; ++k) {
          if (k >= az.n) break;
          const int sg = sxy * az.sgn[k];
          if (sg == 0) continue;
          float val = io<TS>::ld(s + oxy + az.idx[k]);
          val = sg < 0 ? -val : val;
          a0 = fmaf(val, gw * az.w[k], a0);
          a1 = fmaf(val, wg * az.w[k], a1);
          a2 = fmaf(val, ww * gz[k], a2);
        }
      }
    }
    TO* o = out + (long long)c * Vo * 3;
    io<TO>::st(o, a0); io<TO>::st(o + 1, a1); io<TO>::st(o + 2, a2);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_grid_pull(const void* src, int src_dtype, int B, int C, int X, int Y, int Z, const void* grid, int grid_dtype,
                              long long grid_stride_b, long long grid_stride_c, long long grid_stride_v, int Xo, int Yo, int Zo,
                              const double* scale3, const double* shift3, const int* bound3, const int* order3, int extrapolate,
                              int nearest_half_even, void* out, int out_dtype, void* stream) {
  B200_REQUIRE(src && grid && out && bound3 && order3, "grid_pull: null pointer");
  B200_REQUIRE(B > 0 && C > 0 && X > 0 && Y > 0 && Z > 0 && Xo > 0 && Yo > 0 && Zo > 0, "grid_pull: empty problem");
  B200_REQUIRE(B <= 65535, "grid_pull: batch too large for one launch");
  B200_REQUIRE((long long)X * Y * Z < (1LL << 31), "grid_pull: source volume too large");
  GridPullP p;
  p.src = src; p.grid = grid; p.out = out;
  p.B = B; p.C = C; p.X = X; p.Y = Y; p.Z = Z; p.Xo = Xo; p.Yo = Yo; p.Zo = Zo;
  p.g_sb = grid_stride_b; p.g_sc = grid_stride_c; p.g_sv = grid_stride_v;
  int max_order = 0;
  for (int a = 0; a < 3; ++a) {
    p.scale[a] = scale3 ? scale3[a] : 1.0; p.shift[a] = shift3 ? shift3[a] : 0.0;
    p.bound[a] = bound3[a]; p.order[a] = order3[a];
    B200_REQUIRE(order3[a] >= 0 && order3[a] <= 7, "grid_pull: interpolation order must be 0..7 (got %d)", order3[a]);
    B200_REQUIRE((bound3[a] >= 0 && bound3[a] <= 5) || bound3[a] == 7, "grid_pull: bound must be 0 replicate, 1 dct1, 2 dct2, 3 dst1, 4 dst2, 5 dft or 7 zero (got %d)", bound3[a]);
    max_order = std::max(max_order, order3[a]);
  }
  p.extrapolate = extrapolate; p.half_even = nearest_half_even;
  const long long Vo = (long long)Xo * Yo * Zo;
  dim3 grid_dim((unsigned)((Vo + 127) / 128), B);
  cudaStream_t st = (cudaStream_t)stream;
  // dtype dispatch: source f16/f32, grid f32/f64 (0 = f32, 2 = f64), output f16/f32
  B200_REQUIRE(src_dtype == B200_DT_F32 || src_dtype == B200_DT_F16, "grid_pull: source must be float32 or float16");
  B200_REQUIRE(out_dtype == B200_DT_F32 || out_dtype == B200_DT_F16, "grid_pull: output must be float32 or float16");
  B200_REQUIRE(grid_dtype == B200_DT_F32 || grid_dtype == 2, "grid_pull: grid must be float32 (0) or float64 (2)");
#define GP3(TS, TG, TO) do { if (max_order <= 1) grid_pull_kernel<TS, TG, TO, 2><<<grid_dim, 128, 0, st>>>(p); \
                             else grid_pull_kernel<TS, TG, TO, 8><<<grid_dim, 128, 0, st>>>(p); } while (0)
#define GP2(TS, TG) do { if (out_dtype == B200_DT_F16) GP3(TS, TG, __half); else GP3(TS, TG, float); } while (0)
#define GP1(TS) do { if (grid_dtype == 2) GP2(TS, double); else GP2(TS, float); } while (0)
  if (src_dtype == B200_DT_F16) GP1(__half); else GP1(float);
#undef GP1
#undef GP2
#undef GP3
  B200_LAUNCH_CHECK("grid_pull_kernel");
  return B200_OK;
}

// grid_push (input != null) and grid_count (input == null, C = 1): out [B, C, X, Y, Z] fp32 is zeroed here, then splatted into.
extern "C" int b200_grid_push(const void* input, int in_dtype, int B, int C, int Xi, int Yi, int Zi, const void* grid, int grid_dtype,
                              long long grid_stride_b, long long grid_stride_c, long long grid_stride_v, int X, int Y, int Z,
                              const double* scale3, const double* shift3, const int* bound3, const int* order3, int extrapolate,
                              void* out, void* stream) {
  B200_REQUIRE(grid && out && bound3 && order3, "grid_push: null pointer");
  B200_REQUIRE(B > 0 && C > 0 && X > 0 && Y > 0 && Z > 0 && Xi > 0 && Yi > 0 && Zi > 0, "grid_push: empty problem");
  B200_REQUIRE(input || C == 1, "grid_count: one channel");
  B200_REQUIRE(B <= 65535, "grid_push: batch too large for one launch");
  B200_REQUIRE((long long)X * Y * Z < (1LL << 31), "grid_push: output volume too large");
  GridPullP p;
  p.src = input; p.grid = grid; p.out = out;
  p.B = B; p.C = C; p.X = X; p.Y = Y; p.Z = Z; p.Xo = Xi; p.Yo = Yi; p.Zo = Zi;
  p.g_sb = grid_stride_b; p.g_sc = grid_stride_c; p.g_sv = grid_stride_v;
  int max_order = 0;
  for (int a = 0; a < 3; ++a) {
    p.scale[a] = scale3 ? scale3[a] : 1.0; p.shift[a] = shift3 ? shift3[a] : 0.0;
    p.bound[a] = bound3[a]; p.order[a] = order3[a];
    B200_REQUIRE(order3[a] >= 0 && order3[a] <= 7, "grid_push: interpolation order must be 0..7 (got %d)", order3[a]);
    B200_REQUIRE((bound3[a] >= 0 && bound3[a] <= 5) || bound3[a] == 7, "grid_push: bound must be 0 replicate, 1 dct1, 2 dct2, 3 dst1, 4 dst2, 5 dft or 7 zero (got %d)", bound3[a]);
    max_order = std::max(max_order, order3[a]);
  }
  p.extrapolate = extrapolate; p.half_even = 0;
  B200_REQUIRE(in_dtype == B200_DT_F32 || in_dtype == B200_DT_F16, "grid_push: input must be float32 or float16");
  B200_REQUIRE(grid_dtype == B200_DT_F32 || grid_dtype == 2, "grid_push: grid must be float32 (0) or float64 (2)");
  cudaStream_t st = (cudaStream_t)stream;
  B200_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * C * X * Y * Z, st));
  const long long Vi = (long long)Xi * Yi * Zi;
  dim3 grid_dim((unsigned)((Vi + 127) / 128), B);
#define GS2(TS, TG) do { if (max_order <= 1) grid_push_kernel<TS, TG, 2><<<grid_dim, 128, 0, st>>>(p); \
                         else grid_push_kernel<TS, TG, 8><<<grid_dim, 128, 0, st>>>(p); } while (0)
#define GS1(TS) do { if (grid_dtype == 2) GS2(TS, double); else GS2(TS, float); } while (0)
  if (in_dtype == B200_DT_F16) GS1(__half); else GS1(float);
#undef GS1
#undef GS2
  B200_LAUNCH_CHECK("grid_push_kernel");
  return B200_OK;
}

// monai._C.grid_grad: out [B, C, Xo, Yo, Zo, 3]
extern "C" int b200_grid_grad(const void* src, int src_dtype, int B, int C, int X, int Y, int Z, const void* grid, int grid_dtype,
                              long long grid_stride_b, long long grid_stride_c, long long grid_stride_v, int Xo, int Yo, int Zo,
                              const double* scale3, const double* shift3, const int* bound3, const int* order3, int extrapolate,
                              void* out, int out_dtype, void* stream) {
  B200_REQUIRE(src && grid && out && bound3 && order3, "grid_grad: null pointer");
  B200_REQUIRE(B > 0 && C > 0 && X > 0 && Y > 0 && Z > 0 && Xo > 0 && Yo > 0 && Zo > 0, "grid_grad: empty problem");
  B200_REQUIRE(B <= 65535, "grid_grad: batch too large for one launch");
  B200_REQUIRE((long long)X * Y * Z < (1LL << 31), "grid_grad: source volume too large");
  GridPullP p;
  p.src = src; p.grid = grid; p.out = out;
  p.B = B; p.C = C; p.X = X; p.Y = Y; p.Z = Z; p.Xo = Xo; p.Yo = Yo; p.Zo = Zo;
  p.g_sb = grid_stride_b; p.g_sc = grid_stride_c; p.g_sv = grid_stride_v;
  int max_order = 0;
  for (int a = 0; a < 3; ++a) {
    p.scale[a] = scale3 ? scale3[a] : 1.0; p.shift[a] = shift3 ? shift3[a] : 0.0;
    p.bound[a] = bound3[a]; p.order[a] = order3[a];
    B200_REQUIRE(order3[a] >= 0 && order3[a] <= 7, "grid_grad: interpolation order must be 0..7 (got %d)", order3[a]);
    B200_REQUIRE((bound3[a] >= 0 && bound3[a] <= 5) || bound3[a] == 7, "grid_grad: bound must be 0 replicate, 1 dct1, 2 dct2, 3 dst1, 4 dst2, 5 dft or 7 zero (got %d)", bound3[a]);
    max_order = std::max(max_order, order3[a]);
  }
  p.extrapolate = extrapolate; p.half_even = 0;
  B200_REQUIRE(src_dtype == B200_DT_F32 || src_dtype == B200_DT_F16, "grid_grad: source must be float32 or float16");
  B200_REQUIRE(out_dtype == B200_DT_F32 || out_dtype == B200_DT_F16, "grid_grad: output must be float32 or float16");
  B200_REQUIRE(grid_dtype == B200_DT_F32 || grid_dtype == 2, "grid_grad: grid must be float32 (0) or float64 (2)");
  const long long Vo = (long long)Xo * Yo * Zo;
  dim3 grid_dim((unsigned)((Vo + 127) / 128), B);
  cudaStream_t st = (cudaStream_t)stream;
#define GG3(TS, TG, TO) do { if (max_order <= 1) grid_grad_kernel<TS, TG, TO, 2><<<grid_dim, 128, 0, st>>>(p); \
                             else grid_grad_kernel<TS, TG, TO, 8><<<grid_dim, 128, 0, st>>>(p); } while (0)
#define GG2(TS, TG) do { if (out_dtype == B200_DT_F16) GG3(TS, TG, __half); else GG3(TS, TG, float); } while (0)
#define GG1(TS) do { if (grid_dtype == 2) GG2(TS, double); else GG2(TS, float); } while (0)
  if (src_dtype == B200_DT_F16) GG1(__half); else GG1(float);
#undef GG1
#undef GG2
#undef GG3
  B200_LAUNCH_CHECK("grid_grad_kernel");
  return B200_OK;
}
