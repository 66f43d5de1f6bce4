// Parameters and small helpers shared by the sliding-window blend kernels (blend.cu, blend_fused.cu).
#pragma once
#include "common.cuh"

namespace b200 {

struct BlendParams {
  const void* preds;          // windows [win_begin, win_end) resident, element strides below
  long long ps_n, ps_c, ps_d, ps_h, ps_w;
  int win_begin, win_end;     // flat window indices (batch-major, then d,h,w "ij" order)
  int B, C, D, H, W;          // blended volume (already padded to >= roi)
  int rd, rh, rw;             // roi
  const int* starts_d; int nd;
  const int* starts_h; int nh;
  const int* starts_w; int nw;
  const float* gd; const float* gh; const float* gw;   // 1-D importance factors
  float clamp_min;
  const float* wmap;          // optional dense roi weight map [rd,rh,rw] (overrides gd/gh/gw)
  void* out;                  // MODE 0: final [B,C,D,H,W] (out dtype); MODE 1: fp32 accumulators (+=)
  const float* acc;           // MODE 2: fp32 accumulators to normalise
  int d0, d1, h0, h1;         // box of output rows to visit (d in [d0,d1), h in [h0,h1))
  int offsets_fit_i32;        // one (d) layer of windows spans < 2^31 prediction elements (8-voxel pipelined path)
  const int* slot_map;        // optional: flat window id -> slot of its prediction in `preds` (-1 = not resident); overrides win_begin
  int n_slots;                // number of resident slots when slot_map is given
};

constexpr int kMaxStarts = 512;

// numerator update.  fp32 predictions: separate multiply and add, the reference's operation order, so results are
// bit-identical to it.  fp16 predictions (where the reference itself accumulates in fp16 and parity is a tolerance): one
// fused multiply-add in fp32 -- fewer instructions and one rounding less.
template <typename TP> __device__ __forceinline__ float blend_acc(float acc, float x, float w);
template <> __device__ __forceinline__ float blend_acc<float>(float acc, float x, float w) { return __fadd_rn(acc, __fmul_rn(x, w)); }
template <> __device__ __forceinline__ float blend_acc<__half>(float acc, float x, float w) { return fmaf(x, w, acc); }

// normalisation sum(w p) / sum(w).  fp32 results: one IEEE division, the reference's `out /= count`.  fp16 results (the
// value is rounded to 11 bits right after): one reciprocal per voxel shared by the channels, a multiply per channel --
// EVERY blend kernel uses this same form, so the one-shot, streaming and fused paths stay bit-identical to each other.
template <typename TO> struct BlendFin;
template <> struct BlendFin<float> {
  static __device__ __forceinline__ float prep(float cnt) { return cnt; }
  static __device__ __forceinline__ float apply(float num, float c) { return __fdiv_rn(num, c); }
};
template <> struct BlendFin<__half> {
  static __device__ __forceinline__ float prep(float cnt) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(cnt)); return r; }
  static __device__ __forceinline__ float apply(float num, float c) { return num * c; }
};

// index range [lo, lo + n) of the (sorted) window starts that cover coordinate x along one axis
__device__ __forceinline__ void blend_cover(const int* __restrict__ st, int ns, int r, int x, int& lo, int& n) {
  lo = 0; n = 0;
  for (int i = 0; i < ns; ++i) { const int s = __ldg(st + i); if (s <= x && x < s + r) { if (!n) lo = i; ++n; } }
}

}  // namespace b200
