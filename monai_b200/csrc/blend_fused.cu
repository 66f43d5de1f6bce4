// Second-generation sliding-window blend kernels:
//
//  * sw_blend8_lean_kernel -- the C2 / C3 / C5 geometry (fp16 predictions, W / roi / W-starts multiples of 8, at most three
//    covering windows per axis, all predictions resident): the same gather-form arithmetic as sw_blend8_kernel (blend.cu)
//    without its software pipeline -- per-axis tables are recomputed from a few integers instead of living in indexed
//    register arrays (those spilled to local memory), and memory-level parallelism comes from 24 resident warps per SM,
//    each with up to six 16-byte loads in flight.  Bit-identical to the other blend kernels.
//
//  * sw_blend_resample_kernel -- the north star's fused kernel: out = trilinear_resample(blend(predictions), M) in ONE pass,
//    i.e. monai/inferers/utils.py:286-298,351-360 (weighted overlap-add and normalisation) composed with the affine
//    resampling that follows the inferer in a segmentation bundle (Invertd of Spacingd: monai/transforms/spatial/array.py:
//    545-546 -> monai/transforms/spatial/functional.py:68-184).  The blended volume is never written: for every output voxel
//    the (up to) eight source corners are blended on the fly from the resident window predictions (gather form, the same
//    ascending window order and fp32 operations as the plain blend) and interpolated.  With an identity matrix every output
//    voxel has one corner of weight 1, so the result is bit-identical to the plain blend.
#include "common.cuh"
#include "blend.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

__device__ __forceinline__ void cvt8h(const uint4& r, float (&v)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
}
template <typename TO> __device__ __forceinline__ void st8(TO* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__half>(__half* p, const float (&v)[8]) {
  uint4 r;
  __half2* h = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

constexpr int kLeanRows = 8, kLeanMaxRoiW = 512, kLeanK = 3;
constexpr int kLeanDZ = 16;   // most D planes per block: the per-block prologue (weight table, covering ranges, barrier) is paid once for
                              // `dz` planes -- with one plane per block (65 536 blocks of 16 loads per thread on C3: 1.06 ms) it cost 15 %
                              // of the run time (8 planes: 0.91 ms); small volumes keep fewer planes per block so the grid still fills the GPU

// block = 8 warps; a warp covers a compact 8 (h) x 32 (w) patch: lane = row * 4 + octet (as sw_blend8_kernel)
template <typename TO>
__global__ void __launch_bounds__(32 * kLeanRows, 3) sw_blend8_lean_kernel(BlendParams p, int dzn) {
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int w8 = (blockIdx.x * 32 + wrp * 4 + (lane & 3)) * 8;
  const int h = p.h0 + blockIdx.y * kLeanRows + (lane >> 2);
  const int nd_box = p.d1 - p.d0;
  const int nzb = (nd_box + dzn - 1) / dzn;
  const int dbase = (blockIdx.z % nzb) * dzn + p.d0, b = blockIdx.z / nzb;
  __shared__ __align__(16) float s_gw[kLeanMaxRoiW];
  __shared__ int s_cov[32 + kLeanRows + kLeanDZ];      // covering window ranges (lo | count << 16): 32 octets, 8 rows, the planes
  for (int i = threadIdx.x; i < p.rw; i += blockDim.x) s_gw[i] = __ldg(p.gw + i);
  if (threadIdx.x < 32 + kLeanRows + kLeanDZ) {
    const int t = threadIdx.x;
    const int* st = t < 32 ? p.starts_w : (t < 32 + kLeanRows ? p.starts_h : p.starts_d);
    const int ns = t < 32 ? p.nw : (t < 32 + kLeanRows ? p.nh : p.nd);
    const int r = t < 32 ? p.rw : (t < 32 + kLeanRows ? p.rh : p.rd);
    const int xq = t < 32 ? (blockIdx.x * 32 + t) * 8 : (t < 32 + kLeanRows ? p.h0 + blockIdx.y * kLeanRows + (t - 32) : dbase + (t - 32 - kLeanRows));
    int lo, cn;
    blend_cover(st, ns, r, xq, lo, cn);
    s_cov[t] = lo | (cn << 16);
  }
  __syncthreads();
  if (h >= p.h1 || w8 >= p.W) return;
  const int cw = s_cov[wrp * 4 + (lane & 3)], ch = s_cov[32 + (lane >> 2)];
  const int ih_lo = ch & 0xffff, nhc = ch >> 16, iw_lo = cw & 0xffff, nwc = cw >> 16;
  const int num_win = p.nd * p.nh * p.nw;
  const long long vol = (long long)p.D * p.H * p.W;
  const __half* __restrict__ preds = (const __half*)p.preds;
  // local W coordinate of the octet inside each covering W window (kLeanK >= nwc is guaranteed by the dispatcher)
  int lwk[kLeanK];
#pragma unroll
  for (int k = 0; k < kLeanK; ++k) lwk[k] = w8 - __ldg(p.starts_w + iw_lo + (k < nwc ? k : 0));
  for (int dz = 0; dz < dzn; ++dz) {
  const int d = dbase + dz;
  if (d >= p.d1) break;
  const int cd = s_cov[32 + kLeanRows + dz];
  const int id_lo = cd & 0xffff, ndc = cd >> 16;
  const long long voff = ((long long)d * p.H + h) * p.W + w8;
  for (int c0 = 0; c0 < p.C; c0 += 2) {
    const bool two = c0 + 1 < p.C;
    float cnt[8], a0[8], a1[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) { cnt[v] = 0.f; a0[v] = 0.f; a1[v] = 0.f; }
    for (int a = 0; a < ndc; ++a) {
      const int ia = id_lo + a;
      const int ld = d - __ldg(p.starts_d + ia);
      const float gdv = __ldg(p.gd + ld);
      for (int e = 0; e < nhc; ++e) {
        const int ie = ih_lo + e;
        const int lh = h - __ldg(p.starts_h + ie);
        const float gdh = __fmul_rn(gdv, __ldg(p.gh + lh));
        const __half* pp = preds + ((long long)b * num_win + ((long long)ia * p.nh + ie) * p.nw + iw_lo) * p.ps_n + (long long)ld * p.ps_d +
                           (long long)lh * p.ps_h + (long long)c0 * p.ps_c;
        uint4 r0[kLeanK], r1[kLeanK];
#pragma unroll
        for (int k = 0; k < kLeanK; ++k) {
          if (k < nwc) {
            const __half* q = pp + (long long)k * p.ps_n + lwk[k];
            r0[k] = __ldg(reinterpret_cast<const uint4*>(q));
            if (two) r1[k] = __ldg(reinterpret_cast<const uint4*>(q + p.ps_c));
          }
        }
#pragma unroll
        for (int k = 0; k < kLeanK; ++k) {
          if (k < nwc) {
            float t[8], xv[8];
            const float4 g0 = *reinterpret_cast<const float4*>(s_gw + lwk[k]), g1 = *reinterpret_cast<const float4*>(s_gw + lwk[k] + 4);
            t[0] = g0.x; t[1] = g0.y; t[2] = g0.z; t[3] = g0.w; t[4] = g1.x; t[5] = g1.y; t[6] = g1.z; t[7] = g1.w;
#pragma unroll
            for (int v = 0; v < 8; ++v) { t[v] = fmaxf(__fmul_rn(gdh, t[v]), p.clamp_min); cnt[v] = __fadd_rn(cnt[v], t[v]); }
            cvt8h(r0[k], xv);
#pragma unroll
            for (int v = 0; v < 8; ++v) a0[v] = blend_acc<__half>(a0[v], xv[v], t[v]);
            if (two) {
              cvt8h(r1[k], xv);
#pragma unroll
              for (int v = 0; v < 8; ++v) a1[v] = blend_acc<__half>(a1[v], xv[v], t[v]);
            }
          }
        }
      }
    }
    const long long o0 = ((long long)b * p.C + c0) * vol + voff;
#pragma unroll
    for (int v = 0; v < 8; ++v) { const float cf = BlendFin<TO>::prep(cnt[v]); a0[v] = BlendFin<TO>::apply(a0[v], cf); a1[v] = BlendFin<TO>::apply(a1[v], cf); }
    st8<TO>((TO*)p.out + o0, a0);
    if (two) st8<TO>((TO*)p.out + o0 + vol, a1);
  }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// fused blend + affine resample
// ---------------------------------------------------------------------------------------------------------------------
struct BlendRsParams {
  BlendParams b;
  double m[12];            // output voxel index (d, h, w, 1) -> coordinate in the blended volume
  int oD, oH, oW;
  int interp;              // 0 nearest (round half to even), 1 trilinear
  int pad;                 // 0 zeros, 1 border
};

constexpr int kRsCMax = 4;   // channels blended per pass over the covering windows

template <typename TP> __device__ __forceinline__ float ld_pred(const TP* p);
template <> __device__ __forceinline__ float ld_pred<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_pred<__half>(const __half* p) { return __half2float(__ldg(p)); }

// blended values of channels [c0, c0 + nc) at voxel (d, h, w) of batch item b: sum(w p) / sum(w) over the covering windows
// in ascending window index (MODE 0), or acc / sum(w) (MODE 2) -- the arithmetic of sw_blend_kernel, normalised as TO
template <typename TP, typename TO, int MODE>
__device__ __forceinline__ void blend_point(const BlendParams& p, int b, int c0, int nc, int d, int h, int w, float (&val)[kRsCMax]) {
  int id_lo, ndc, ih_lo, nhc, iw_lo, nwc;
  blend_cover(p.starts_d, p.nd, p.rd, d, id_lo, ndc);
  blend_cover(p.starts_h, p.nh, p.rh, h, ih_lo, nhc);
  blend_cover(p.starts_w, p.nw, p.rw, w, iw_lo, nwc);
  const int num_win = p.nd * p.nh * p.nw;
  float cnt = 0.f, num[kRsCMax];
#pragma unroll
  for (int c = 0; c < kRsCMax; ++c) num[c] = 0.f;
  const TP* preds = (const TP*)p.preds;
  for (int a = 0; a < ndc; ++a) {
    const int id = id_lo + a, ld = d - __ldg(p.starts_d + id);
    for (int e = 0; e < nhc; ++e) {
      const int ih = ih_lo + e, lh = h - __ldg(p.starts_h + ih);
      const float gdh = p.wmap ? 0.f : __fmul_rn(__ldg(p.gd + ld), __ldg(p.gh + lh));
      for (int k = 0; k < nwc; ++k) {
        const int iw = iw_lo + k, lw = w - __ldg(p.starts_w + iw);
        const float wt = p.wmap ? __ldg(p.wmap + ((long long)ld * p.rh + lh) * p.rw + lw) : fmaxf(__fmul_rn(gdh, __ldg(p.gw + lw)), p.clamp_min);
        cnt = __fadd_rn(cnt, wt);
        if (MODE == 0) {
          const long long widx = (long long)b * num_win + ((long long)id * p.nh + ih) * p.nw + iw;
          const TP* pp = preds + (widx - p.win_begin) * p.ps_n + (long long)ld * p.ps_d + (long long)lh * p.ps_h + (long long)lw * p.ps_w + (long long)c0 * p.ps_c;
#pragma unroll
          for (int c = 0; c < kRsCMax; ++c)
            if (c < nc) num[c] = blend_acc<TP>(num[c], ld_pred<TP>(pp + (long long)c * p.ps_c), wt);
        }
      }
    }
  }
  const float cf = BlendFin<TO>::prep(cnt);
  const long long vol = (long long)p.D * p.H * p.W;
#pragma unroll
  for (int c = 0; c < kRsCMax; ++c) {
    if (c < nc) {
      if (MODE == 2) num[c] = p.acc[((long long)b * p.C + c0 + c) * vol + ((long long)d * p.H + h) * p.W + w];
      val[c] = BlendFin<TO>::apply(num[c], cf);
    }
  }
}

template <typename TP, typename TO, int MODE>
__global__ void __launch_bounds__(128) sw_blend_resample_kernel(BlendRsParams q) {
  const BlendParams& p = q.b;
  const int ow = blockIdx.x * blockDim.x + threadIdx.x;
  const int oh = blockIdx.y;
  const int od = blockIdx.z % q.oD, b = blockIdx.z / q.oD;
  if (ow >= q.oW) return;
  // source coordinate in fp64, the reference's coordinate dtype (and the same fma chain as resample_affine_kernel)
  double ca = fma(q.m[0], (double)od, fma(q.m[1], (double)oh, fma(q.m[2], (double)ow, q.m[3])));
  double cb = fma(q.m[4], (double)od, fma(q.m[5], (double)oh, fma(q.m[6], (double)ow, q.m[7])));
  double cc = fma(q.m[8], (double)od, fma(q.m[9], (double)oh, fma(q.m[10], (double)ow, q.m[11])));
  const double lim = 1.0e9;
  ca = fmin(lim, fmax(-lim, ca)); cb = fmin(lim, fmax(-lim, cb)); cc = fmin(lim, fmax(-lim, cc));
  if (q.pad == 1) {
    ca = fmin((double)(p.D - 1), fmax(ca, 0.0)); cb = fmin((double)(p.H - 1), fmax(cb, 0.0)); cc = fmin((double)(p.W - 1), fmax(cc, 0.0));
  }
  int a0, b0, c0i;
  float wa[2], wb[2], wc[2];
  if (q.interp == 0) {
    a0 = (int)nearbyint(ca); b0 = (int)nearbyint(cb); c0i = (int)nearbyint(cc);
    wa[0] = wb[0] = wc[0] = 1.f; wa[1] = wb[1] = wc[1] = 0.f;
  } else {
    const double fa = floor(ca), fb = floor(cb), fc = floor(cc);
    a0 = (int)fa; b0 = (int)fb; c0i = (int)fc;
    const float ta = (float)(ca - fa), tb = (float)(cb - fb), tc = (float)(cc - fc);
    wa[0] = 1.f - ta; wa[1] = ta; wb[0] = 1.f - tb; wb[1] = tb; wc[0] = 1.f - tc; wc[1] = tc;
  }
  // corners outside the blended volume carry weight 0 (zeros padding; with border padding only the +1 corner of a
  // coordinate clamped to size-1 can be outside, and its weight is already 0)
  if (a0 < 0 || a0 >= p.D) wa[0] = 0.f;
  if (a0 + 1 < 0 || a0 + 1 >= p.D) wa[1] = 0.f;
  if (b0 < 0 || b0 >= p.H) wb[0] = 0.f;
  if (b0 + 1 < 0 || b0 + 1 >= p.H) wb[1] = 0.f;
  if (c0i < 0 || c0i >= p.W) wc[0] = 0.f;
  if (c0i + 1 < 0 || c0i + 1 >= p.W) wc[1] = 0.f;
  const long long ovol = (long long)q.oD * q.oH * q.oW;
  const long long ooff = ((long long)od * q.oH + oh) * q.oW + ow;
  for (int ch0 = 0; ch0 < p.C; ch0 += kRsCMax) {
    const int nc = min(kRsCMax, p.C - ch0);
    float out[kRsCMax];
#pragma unroll
    for (int c = 0; c < kRsCMax; ++c) out[c] = 0.f;
#pragma unroll 1
    for (int cr = 0; cr < 8; ++cr) {
      const int da = cr >> 2, db = (cr >> 1) & 1, dc = cr & 1;
      const float wgt = wa[da] * wb[db] * wc[dc];
      if (wgt == 0.f) continue;
      float val[kRsCMax];
      blend_point<TP, TO, MODE>(p, b, ch0, nc, a0 + da, b0 + db, c0i + dc, val);
#pragma unroll
      for (int c = 0; c < kRsCMax; ++c)
        if (c < nc) out[c] = fmaf(val[c], wgt, out[c]);
    }
#pragma unroll
    for (int c = 0; c < kRsCMax; ++c)
      if (c < nc) io<TO>::st((TO*)p.out + ((long long)b * p.C + ch0 + c) * ovol + ooff, out[c]);
  }
}

int launch_blend8_lean(const BlendParams& p, int out_dtype, cudaStream_t st) {
  dim3 block(32 * kLeanRows);
  // planes per block: as many as keep >= 12 waves of blocks (3 resident blocks per SM)
  const long long per_plane = (long long)ceil_div(p.W / 8, 32) * ceil_div(p.h1 - p.h0, kLeanRows) * p.B;
  int dzn = kLeanDZ;
  while (dzn > 1 && per_plane * ceil_div(p.d1 - p.d0, dzn) < 12LL * 3 * num_sms()) dzn >>= 1;
  dim3 grid(ceil_div(p.W / 8, 32), ceil_div(p.h1 - p.h0, kLeanRows), ceil_div(p.d1 - p.d0, dzn) * p.B);
  if (grid.y == 0 || grid.z == 0) return B200_OK;
  B200_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "sw_blend: volume too large for the launch grid");
  if (out_dtype == B200_DT_F16) sw_blend8_lean_kernel<__half><<<grid, block, 0, st>>>(p, dzn);
  else sw_blend8_lean_kernel<float><<<grid, block, 0, st>>>(p, dzn);
  B200_LAUNCH_CHECK("sw_blend8_lean_kernel");
  return B200_OK;
}

int launch_blend_resample(const BlendParams& p, const double* m, int oD, int oH, int oW, int interp, int pad, int mode, int pred_dtype,
                          int out_dtype, cudaStream_t st) {
  B200_REQUIRE(mode == 0 || mode == 2, "sw_blend: the fused resample needs all predictions resident (mode 0) or the accumulators (mode 2)");
  B200_REQUIRE(oD > 0 && oH > 0 && oW > 0 && oH <= 65535 && (long long)oD * p.B <= 65535, "sw_blend: bad output grid for the fused resample");
  B200_REQUIRE(interp == 0 || interp == 1, "sw_blend: fused resample interpolation must be 0 (nearest) or 1 (trilinear)");
  B200_REQUIRE(pad == 0 || pad == 1, "sw_blend: fused resample padding must be 0 (zeros) or 1 (border)");
  B200_REQUIRE(mode == 2 || (!p.slot_map && p.win_begin == 0 && p.win_end == p.B * p.nd * p.nh * p.nw), "sw_blend: the fused resample (mode 0) needs every window resident");
  BlendRsParams q;
  q.b = p;
  for (int i = 0; i < 12; ++i) q.m[i] = m[i];
  q.oD = oD; q.oH = oH; q.oW = oW; q.interp = interp; q.pad = pad;
  dim3 block(oW >= 128 ? 128 : (oW >= 64 ? 64 : 32));
  dim3 grid(ceil_div(oW, block.x), oH, oD * p.B);
#define LR(TP, TO, M) sw_blend_resample_kernel<TP, TO, M><<<grid, block, 0, st>>>(q)
  if (mode == 2) {
    if (out_dtype == B200_DT_F16) LR(float, __half, 2); else LR(float, float, 2);
  } else if (pred_dtype == B200_DT_F16) {
    if (out_dtype == B200_DT_F16) LR(__half, __half, 0); else LR(__half, float, 0);
  } else {
    if (out_dtype == B200_DT_F16) LR(float, __half, 0); else LR(float, float, 0);
  }
#undef LR
  B200_LAUNCH_CHECK("sw_blend_resample_kernel");
  return B200_OK;
}

}  // namespace b200
