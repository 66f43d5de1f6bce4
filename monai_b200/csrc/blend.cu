// Sliding-window gather + importance-weighted overlap blend.
//
// Replaces monai/inferers/utils.py:217-224 (window gather), :264-275 (count map), :286-288 + :351-360
// (weighted scatter-add) and :297-298 (normalise) of the reference.  The blend is written in *gather form*:
// one thread owns one output voxel and walks the (Cartesian) window table in ascending window index, so
//   out[b,c,v] = ( sum_w imp(v - s_w) * pred_w[c, v - s_w] ) / ( sum_w imp(v - s_w) )
// is produced with the same fp32 operation order as the reference loop (mul, then sequential adds, then one
// IEEE divide) and the count map is never materialised.  imp() is evaluated on the fly from the three 1-D
// vectors of compute_importance_map (monai/data/utils.py:1084-1134): ((g_d*g_h)*g_w) clamped from below.
#include "common.cuh"
#include "../../include/monai_b200.h"

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
};

constexpr int kMaxStarts = 512;

// MODE 0: all windows resident -> write normalised result.  MODE 1: accumulate numerators (+=) for the
// resident window range.  MODE 2: divide accumulators by the analytic count (all windows).
template <typename TP, typename TO, int MODE>
__global__ void __launch_bounds__(256) sw_blend_kernel(BlendParams p) {
  __shared__ int s_w[kMaxStarts];
  __shared__ int s_did[32], s_hid[32];
  __shared__ int s_ndc, s_nhc;
  const int h = blockIdx.y + p.h0;
  const int d = blockIdx.z % (p.d1 - p.d0) + p.d0;
  const int b = blockIdx.z / (p.d1 - p.d0);
  for (int i = threadIdx.x; i < p.nw; i += blockDim.x) s_w[i] = p.starts_w[i];
  if (threadIdx.x == 0) {
    int n = 0;
    for (int i = 0; i < p.nd && n < 32; ++i) { int s = p.starts_d[i]; if (s <= d && d < s + p.rd) s_did[n++] = i; }
    s_ndc = n; n = 0;
    for (int i = 0; i < p.nh && n < 32; ++i) { int s = p.starts_h[i]; if (s <= h && h < s + p.rh) s_hid[n++] = i; }
    s_nhc = n;
  }
  __syncthreads();
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= p.W) return;
  const int num_win = p.nd * p.nh * p.nw;
  const long long vol = (long long)p.D * p.H * p.W;
  const long long voff = ((long long)d * p.H + h) * p.W + w;

  float cnt = 0.f;
  float accv[8];
  constexpr int CMAX = 8;
  // channels are processed in groups of CMAX so the weight walk is shared between channels
  for (int c0 = 0; c0 < (MODE == 2 ? 1 : p.C); c0 += CMAX) {
    const int cn = min(CMAX, p.C - c0);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) accv[c] = 0.f;
    if (MODE == 1) {  // continue the running sums so the addition order stays "one window after the other"
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < cn) accv[c] = *((const float*)p.out + ((long long)b * p.C + c0 + c) * vol + voff);
    }
    cnt = 0.f;
    for (int a = 0; a < s_ndc; ++a) {
      const int id = s_did[a];
      const int ld = d - p.starts_d[id];
      for (int e = 0; e < s_nhc; ++e) {
        const int ih = s_hid[e];
        const int lh = h - p.starts_h[ih];
        const float gdh = p.wmap ? 0.f : __fmul_rn(p.gd[ld], p.gh[lh]);
        for (int iw = 0; iw < p.nw; ++iw) {
          const int lw = w - s_w[iw];
          if (lw < 0 || lw >= p.rw) continue;
          float wt;
          if (p.wmap) wt = p.wmap[((long long)ld * p.rh + lh) * p.rw + lw];
          else wt = fmaxf(__fmul_rn(gdh, p.gw[lw]), p.clamp_min);
          cnt = __fadd_rn(cnt, wt);
          if (MODE != 2) {
            const int widx = b * num_win + (id * p.nh + ih) * p.nw + iw;
            if (widx < p.win_begin || widx >= p.win_end) continue;
            const TP* pp = (const TP*)p.preds + (long long)(widx - p.win_begin) * p.ps_n +
                           (long long)ld * p.ps_d + (long long)lh * p.ps_h + (long long)lw * p.ps_w +
                           (long long)c0 * p.ps_c;
#pragma unroll
            for (int c = 0; c < CMAX; ++c)
              if (c < cn) accv[c] = __fadd_rn(accv[c], __fmul_rn(io<TP>::ld(pp + c * p.ps_c), wt));
          }
        }
      }
    }
    if (MODE == 0) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < cn) io<TO>::st((TO*)p.out + ((long long)b * p.C + c0 + c) * vol + voff, __fdiv_rn(accv[c], cnt));
    } else if (MODE == 1) {
#pragma unroll
      for (int c = 0; c < CMAX; ++c)
        if (c < cn) *((float*)p.out + ((long long)b * p.C + c0 + c) * vol + voff) = accv[c];
    }
  }
  if (MODE == 2) {
    for (int c = 0; c < p.C; ++c) {
      const long long o = ((long long)b * p.C + c) * vol + voff;
      io<TO>::st((TO*)p.out + o, __fdiv_rn(p.acc[o], cnt));
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) sw_gather_kernel(const TI* __restrict__ vol, TO* __restrict__ out,
                                                        const int* __restrict__ tab, int C, int D, int H, int W,
                                                        int rd, int rh, int rw) {
  // grid: x over w, y over (d*rh + h), z over (win*C + c)
  const int win = blockIdx.z / C, c = blockIdx.z % C;
  const int ld = blockIdx.y / rh, lh = blockIdx.y % rh;
  const int b = tab[win * 4 + 0], sd = tab[win * 4 + 1], sh = tab[win * 4 + 2], sw = tab[win * 4 + 3];
  const TI* src = vol + ((((long long)b * C + c) * D + sd + ld) * H + sh + lh) * W + sw;
  TO* dst = out + ((((long long)win * C + c) * rd + ld) * rh + lh) * (long long)rw;
  for (int lw = blockIdx.x * blockDim.x + threadIdx.x; lw < rw; lw += gridDim.x * blockDim.x)
    io<TO>::st(dst + lw, io<TI>::ld(src + lw));
}

}  // namespace b200

using namespace b200;

template <int MODE>
static int launch_blend(const BlendParams& p, int pred_dtype, int out_dtype, cudaStream_t st) {
  dim3 block(p.W >= 192 ? 256 : (p.W >= 96 ? 128 : 64));
  dim3 grid(ceil_div(p.W, block.x), p.h1 - p.h0, (p.d1 - p.d0) * p.B);
  if (grid.y == 0 || grid.z == 0) return B200_OK;
  B200_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "sw_blend: volume too large for the launch grid");
#define LB(TP, TO) sw_blend_kernel<TP, TO, MODE><<<grid, block, 0, st>>>(p)
  if (MODE == 1) {
    if (pred_dtype == B200_DT_F16) LB(__half, float); else LB(float, float);
  } else if (MODE == 2) {
    if (out_dtype == B200_DT_F16) LB(float, __half); else LB(float, float);
  } else {
    if (pred_dtype == B200_DT_F16 && out_dtype == B200_DT_F16) LB(__half, __half);
    else if (pred_dtype == B200_DT_F16) LB(__half, float);
    else if (out_dtype == B200_DT_F16) LB(float, __half);
    else LB(float, float);
  }
#undef LB
  B200_LAUNCH_CHECK("sw_blend_kernel");
  return B200_OK;
}

extern "C" int b200_sw_blend(const b200_blend_desc* dsc, int mode, void* stream) {
  B200_REQUIRE(dsc != nullptr, "sw_blend: null descriptor");
  B200_REQUIRE(mode >= 0 && mode <= 2, "sw_blend: mode must be 0 (final), 1 (accumulate) or 2 (finalize)");
  B200_REQUIRE(dsc->nw <= kMaxStarts, "sw_blend: more than %d window starts along the last axis", kMaxStarts);
  B200_REQUIRE(dsc->B > 0 && dsc->C > 0 && dsc->D > 0 && dsc->H > 0 && dsc->W > 0, "sw_blend: empty volume");
  B200_REQUIRE(dsc->rd <= dsc->D && dsc->rh <= dsc->H && dsc->rw <= dsc->W, "sw_blend: roi larger than the padded volume");
  B200_REQUIRE(dsc->pred_dtype == B200_DT_F32 || dsc->pred_dtype == B200_DT_F16, "sw_blend: bad pred dtype");
  B200_REQUIRE(dsc->out_dtype == B200_DT_F32 || dsc->out_dtype == B200_DT_F16, "sw_blend: bad out dtype");
  BlendParams p;
  p.preds = dsc->preds;
  p.ps_n = dsc->pred_stride[0]; p.ps_c = dsc->pred_stride[1]; p.ps_d = dsc->pred_stride[2];
  p.ps_h = dsc->pred_stride[3]; p.ps_w = dsc->pred_stride[4];
  p.win_begin = dsc->win_begin; p.win_end = dsc->win_end;
  p.B = dsc->B; p.C = dsc->C; p.D = dsc->D; p.H = dsc->H; p.W = dsc->W;
  p.rd = dsc->rd; p.rh = dsc->rh; p.rw = dsc->rw;
  p.starts_d = dsc->starts_d; p.nd = dsc->nd; p.starts_h = dsc->starts_h; p.nh = dsc->nh;
  p.starts_w = dsc->starts_w; p.nw = dsc->nw;
  p.gd = dsc->gd; p.gh = dsc->gh; p.gw = dsc->gw; p.clamp_min = dsc->clamp_min; p.wmap = dsc->wmap;
  p.out = dsc->out; p.acc = dsc->acc;
  p.d0 = dsc->box[0]; p.d1 = dsc->box[1]; p.h0 = dsc->box[2]; p.h1 = dsc->box[3];
  if (p.d1 <= 0) { p.d0 = 0; p.d1 = p.D; }
  if (p.h1 <= 0) { p.h0 = 0; p.h1 = p.H; }
  B200_REQUIRE(p.d0 >= 0 && p.d1 <= p.D && p.h0 >= 0 && p.h1 <= p.H, "sw_blend: box outside the volume");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 0) return launch_blend<0>(p, dsc->pred_dtype, dsc->out_dtype, st);
  if (mode == 1) return launch_blend<1>(p, dsc->pred_dtype, dsc->out_dtype, st);
  return launch_blend<2>(p, dsc->pred_dtype, dsc->out_dtype, st);
}

extern "C" int b200_sw_gather(const void* vol, int in_dtype, void* out, int out_dtype, const int32_t* win_tab,
                              int n_win, int C, int D, int H, int W, int rd, int rh, int rw, void* stream) {
  if (n_win == 0) return B200_OK;
  B200_REQUIRE(vol && out && win_tab, "sw_gather: null pointer");
  B200_REQUIRE(rd <= D && rh <= H && rw <= W, "sw_gather: roi larger than the volume");
  B200_REQUIRE((long long)n_win * C <= 65535 && (long long)rd * rh <= 65535, "sw_gather: batch too large for one launch");
  dim3 block(rw >= 128 ? 128 : 64), grid(ceil_div(rw, block.x), rd * rh, n_win * C);
  cudaStream_t st = (cudaStream_t)stream;
#define LG(TI, TO) sw_gather_kernel<TI, TO><<<grid, block, 0, st>>>((const TI*)vol, (TO*)out, win_tab, C, D, H, W, rd, rh, rw)
  if (in_dtype == B200_DT_F16 && out_dtype == B200_DT_F16) LG(__half, __half);
  else if (in_dtype == B200_DT_F16 && out_dtype == B200_DT_F32) LG(__half, float);
  else if (in_dtype == B200_DT_F32 && out_dtype == B200_DT_F16) LG(float, __half);
  else if (in_dtype == B200_DT_F32 && out_dtype == B200_DT_F32) LG(float, float);
  else return set_err(B200_ERR_INVALID, "sw_gather: bad dtype");
#undef LG
  B200_LAUNCH_CHECK("sw_gather_kernel");
  return B200_OK;
}
