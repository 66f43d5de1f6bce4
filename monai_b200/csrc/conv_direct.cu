// Direct 3-D convolution / transposed convolution on CUDA cores, fp32 accumulation.
//
// This is the exact-parity path: it serves fp32 models (config C1, BasicUNet fp32) and the layers whose
// channel counts cannot feed a tensor-core tile (Cin = 1 stems, Cout = 2 heads, stride-2 UNet layers).
// Semantics follow nn.Conv3d / nn.ConvTranspose3d as instantiated by
// monai/networks/blocks/convolutions.py:131-152 (padding from same_padding, layers/convutils.py:22-43,
// output_padding = stride - 1, convutils.py:46-53): zero padding, dilation 1, groups 1.
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

struct ConvP {
  b200_conv_desc d;
  const void* x; const float* w; const float* bias; void* y;
  int ntaps;
  int cls;          // number of parity classes (transposed: sd*sh*sw, else 1)
  int Dc, Hc, Wc;   // coarse output extent per parity class
};

constexpr int kCiT = 8;

// Each thread owns one output voxel and CO_T output channels.  Weights for (CO_T couts x kCiT cins x taps)
// are staged in shared memory as [tap][ci][co] so the inner product reads them as broadcast vectors.
// A per-block tap table (valid flag + input offset per axis) removes every division from the inner loops: for a
// forward conv  i = o*stride + (k - pad); for a transposed conv all voxels of a block share one parity class, so
// i = o_coarse + (parity + pad - k)/stride for the taps whose numerator is divisible (the others contribute nothing).
template <typename TI, typename TO, int CO_T>
__global__ void __launch_bounds__(128) conv3d_direct_kernel(ConvP p) {
  extern __shared__ float s_wt[];  // [ntaps][kCiT][CO_T]
  __shared__ int s_tap[128][4];    // tap id, dz, dy, dx of the taps that are live for this block
  __shared__ int s_ntap;
  const b200_conv_desc& d = p.d;
  const int co_blocks = (d.Cout + CO_T - 1) / CO_T;
  const int cob = blockIdx.y % co_blocks;
  const int cls = blockIdx.y / co_blocks;
  const int n = blockIdx.z;
  const int co0 = cob * CO_T;
  int pz = 0, py = 0, px = 0;
  if (d.transposed) { px = cls % d.sw; py = (cls / d.sw) % d.sh; pz = cls / (d.sw * d.sh); }
  if (threadIdx.x == 0) {
    int nt = 0;
    for (int tap = 0; tap < p.ntaps && nt < 128; ++tap) {
      const int kz = tap / (d.kh * d.kw), ky = (tap / d.kw) % d.kh, kx = tap % d.kw;
      int dz, dy, dx;
      if (d.transposed) {
        const int tz = pz + d.pd - kz, ty = py + d.ph - ky, tx = px + d.pw - kx;
        // floor-division-safe divisibility test (numerators may be negative)
        if (((tz % d.sd) + d.sd) % d.sd || ((ty % d.sh) + d.sh) % d.sh || ((tx % d.sw) + d.sw) % d.sw) continue;
        dz = (tz - (((tz % d.sd) + d.sd) % d.sd)) / d.sd; dy = (ty - (((ty % d.sh) + d.sh) % d.sh)) / d.sh; dx = (tx - (((tx % d.sw) + d.sw) % d.sw)) / d.sw;
      } else {
        dz = kz - d.pd; dy = ky - d.ph; dx = kx - d.pw;
      }
      s_tap[nt][0] = tap; s_tap[nt][1] = dz; s_tap[nt][2] = dy; s_tap[nt][3] = dx;
      ++nt;
    }
    s_ntap = nt;
  }
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncoarse = (long long)p.Dc * p.Hc * p.Wc;
  const bool active = t < ncoarse;
  int cz = 0, cy = 0, cx = 0;
  if (active) { cx = (int)(t % p.Wc); cy = (int)((t / p.Wc) % p.Hc); cz = (int)(t / ((long long)p.Wc * p.Hc)); }
  const int oz = d.transposed ? cz * d.sd + pz : cz, oy = d.transposed ? cy * d.sh + py : cy, ox = d.transposed ? cx * d.sw + px : cx;
  const int bz = d.transposed ? cz : cz * d.sd, by = d.transposed ? cy : cy * d.sh, bx = d.transposed ? cx : cx * d.sw;
  const bool inb = active && oz < d.Do && oy < d.Ho && ox < d.Wo;

  float acc[CO_T];
#pragma unroll
  for (int c = 0; c < CO_T; ++c) acc[c] = 0.f;

  const long long in_cs = (long long)d.Di * d.Hi * d.Wi;
  const TI* xin = (const TI*)p.x + (long long)n * d.in_stride_n;

  for (int ci0 = 0; ci0 < d.Cin; ci0 += kCiT) {
    const int cin = min(kCiT, d.Cin - ci0);
    __syncthreads();
    for (int i = threadIdx.x; i < p.ntaps * kCiT * CO_T; i += blockDim.x) {
      const int co = i % CO_T, ci = (i / CO_T) % kCiT, tap = i / (CO_T * kCiT);
      float v = 0.f;
      if (ci < cin && co0 + co < d.Cout) {
        const long long widx = d.transposed
            ? ((long long)(ci0 + ci) * d.Cout + (co0 + co)) * p.ntaps + tap
            : ((long long)(co0 + co) * d.Cin + (ci0 + ci)) * p.ntaps + tap;
        v = p.w[widx];
      }
      s_wt[i] = v;
    }
    __syncthreads();
    if (!inb) continue;
    const int nt = s_ntap;
    for (int q = 0; q < nt; ++q) {
      const int iz = bz + s_tap[q][1], iy = by + s_tap[q][2], ix = bx + s_tap[q][3];
      if (iz < 0 || iz >= d.Di || iy < 0 || iy >= d.Hi || ix < 0 || ix >= d.Wi) continue;
      const TI* xp = xin + (long long)ci0 * in_cs + ((long long)iz * d.Hi + iy) * d.Wi + ix;
      const float* wp = s_wt + s_tap[q][0] * kCiT * CO_T;
      if (cin == kCiT) {
        float xv[kCiT];
#pragma unroll
        for (int ci = 0; ci < kCiT; ++ci) xv[ci] = io<TI>::ld(xp + (long long)ci * in_cs);
#pragma unroll
        for (int ci = 0; ci < kCiT; ++ci) {
#pragma unroll
          for (int c = 0; c < CO_T; ++c) acc[c] = fmaf(xv[ci], wp[ci * CO_T + c], acc[c]);
        }
      } else {
        for (int ci = 0; ci < cin; ++ci) {
          const float xv = io<TI>::ld(xp + (long long)ci * in_cs);
#pragma unroll
          for (int c = 0; c < CO_T; ++c) acc[c] = fmaf(xv, wp[ci * CO_T + c], acc[c]);
        }
      }
    }
  }
  if (!inb) return;
  const long long out_cs = (long long)d.Do * d.Ho * d.Wo;
  TO* yo = (TO*)p.y + (long long)n * d.out_stride_n + ((long long)oz * d.Ho + oy) * d.Wo + ox;
#pragma unroll
  for (int c = 0; c < CO_T; ++c)
    if (co0 + c < d.Cout) io<TO>::st(yo + (long long)(co0 + c) * out_cs, acc[c] + (p.bias ? p.bias[co0 + c] : 0.f));
}

template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int NC, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const long long total = (long long)NC * Do * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), oz = (int)((i / ((long long)Wo * Ho)) % Do);
    const long long nc = i / ((long long)Wo * Ho * Do);
    const T* p = x + ((nc * D + 2 * oz) * H + 2 * oy) * (long long)W + 2 * ox;
    float m = -INFINITY;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c) m = fmaxf(m, io<T>::ld(p + ((long long)a * H + b) * W + c));
    io<T>::st(y + i, m);
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int C, int Di, int Hi, int Wi,
                                     int Ctot, int c_off, int Do, int Ho, int Wo) {
  const long long total = (long long)N * C * Do * Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho), oz = (int)((i / ((long long)Wo * Ho)) % Do);
    const int c = (int)((i / ((long long)Wo * Ho * Do)) % C);
    const int n = (int)(i / ((long long)Wo * Ho * Do * C));
    const int iz = min(oz, Di - 1), iy = min(oy, Hi - 1), ix = min(ox, Wi - 1);
    const T v = x[((((long long)n * C + c) * Di + iz) * Hi + iy) * Wi + ix];
    y[((((long long)n * Ctot + c_off + c) * Do + oz) * Ho + oy) * Wo + ox] = v;
  }
}

}  // namespace b200

using namespace b200;

template <int CO_T>
static int launch_conv(const ConvP& p, cudaStream_t st) {
  const b200_conv_desc& d = p.d;
  const long long ncoarse = (long long)p.Dc * p.Hc * p.Wc;
  const int co_blocks = (d.Cout + CO_T - 1) / CO_T;
  dim3 block(128), grid(ceil_div(ncoarse, 128), co_blocks * p.cls, d.N);
  B200_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "conv3d_direct: grid too large");
  const size_t smem = (size_t)p.ntaps * kCiT * CO_T * sizeof(float);
  B200_REQUIRE(smem <= 40 * 1024 && p.ntaps <= 128, "conv3d_direct: kernel volume too large (%d taps)", p.ntaps);
#define LC(TI, TO) conv3d_direct_kernel<TI, TO, CO_T><<<grid, block, smem, st>>>(p)
  if (d.in_dtype == B200_DT_F16 && d.out_dtype == B200_DT_F16) LC(__half, __half);
  else if (d.in_dtype == B200_DT_F16 && d.out_dtype == B200_DT_F32) LC(__half, float);
  else if (d.in_dtype == B200_DT_F32 && d.out_dtype == B200_DT_F16) LC(float, __half);
  else if (d.in_dtype == B200_DT_F32 && d.out_dtype == B200_DT_F32) LC(float, float);
  else return set_err(B200_ERR_INVALID, "conv3d_direct: bad dtype");
#undef LC
  B200_LAUNCH_CHECK("conv3d_direct_kernel");
  return B200_OK;
}

extern "C" int b200_conv3d_direct(const b200_conv_desc* desc, const void* x, const float* weight, const float* bias,
                                  void* y, void* stream) {
  B200_REQUIRE(desc && x && weight && y, "conv3d_direct: null pointer");
  const b200_conv_desc& d = *desc;
  B200_REQUIRE(d.N > 0 && d.Cin > 0 && d.Cout > 0, "conv3d_direct: empty problem");
  B200_REQUIRE(d.sd > 0 && d.sh > 0 && d.sw > 0 && d.kd > 0 && d.kh > 0 && d.kw > 0, "conv3d_direct: bad kernel/stride");
  // shape consistency (same formulas as torch)
  if (!d.transposed) {
    B200_REQUIRE(d.Do == (d.Di + 2 * d.pd - d.kd) / d.sd + 1 && d.Ho == (d.Hi + 2 * d.ph - d.kh) / d.sh + 1 &&
                 d.Wo == (d.Wi + 2 * d.pw - d.kw) / d.sw + 1, "conv3d_direct: output shape does not match conv arithmetic");
  } else {
    const int lo_d = (d.Di - 1) * d.sd - 2 * d.pd + d.kd, lo_h = (d.Hi - 1) * d.sh - 2 * d.ph + d.kh,
              lo_w = (d.Wi - 1) * d.sw - 2 * d.pw + d.kw;
    B200_REQUIRE(d.Do >= lo_d && d.Do < lo_d + d.sd && d.Ho >= lo_h && d.Ho < lo_h + d.sh && d.Wo >= lo_w && d.Wo < lo_w + d.sw,
                 "conv3d_direct: output shape does not match transposed-conv arithmetic");
  }
  ConvP p;
  p.d = d; p.x = x; p.w = weight; p.bias = bias; p.y = y;
  p.ntaps = d.kd * d.kh * d.kw;
  if (d.transposed) {
    p.cls = d.sd * d.sh * d.sw;
    p.Dc = ceil_div(d.Do, d.sd); p.Hc = ceil_div(d.Ho, d.sh); p.Wc = ceil_div(d.Wo, d.sw);
  } else {
    p.cls = 1; p.Dc = d.Do; p.Hc = d.Ho; p.Wc = d.Wo;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (d.Cout <= 2) return launch_conv<2>(p, st);
  if (d.Cout <= 4) return launch_conv<4>(p, st);
  if (d.Cout <= 8 || d.Cout % 16) return launch_conv<8>(p, st);
  return launch_conv<16>(p, st);
}

extern "C" int b200_maxpool3d_2(const void* x, int dtype, int NC, int D, int H, int W, void* y, void* stream) {
  B200_REQUIRE(x && y, "maxpool3d_2: null pointer");
  const long long total = (long long)NC * (D / 2) * (H / 2) * (W / 2);
  if (total == 0) return B200_OK;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F16) maxpool2_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)x, (__half*)y, NC, D, H, W);
  else if (dtype == B200_DT_F32) maxpool2_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, NC, D, H, W);
  else return set_err(B200_ERR_INVALID, "maxpool3d_2: bad dtype");
  B200_LAUNCH_CHECK("maxpool2_kernel");
  return B200_OK;
}

extern "C" int b200_copy_channels(const void* x, int dtype, int N, int C, int Di, int Hi, int Wi, void* y, int Ctot,
                                  int c_off, int Do, int Ho, int Wo, void* stream) {
  B200_REQUIRE(x && y, "copy_channels: null pointer");
  B200_REQUIRE(c_off >= 0 && c_off + C <= Ctot, "copy_channels: channel slice outside the destination");
  const long long total = (long long)N * C * Do * Ho * Wo;
  if (total == 0) return B200_OK;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == B200_DT_F16)
    copy_channels_kernel<__half><<<blocks, 256, 0, st>>>((const __half*)x, (__half*)y, N, C, Di, Hi, Wi, Ctot, c_off, Do, Ho, Wo);
  else if (dtype == B200_DT_F32)
    copy_channels_kernel<float><<<blocks, 256, 0, st>>>((const float*)x, (float*)y, N, C, Di, Hi, Wi, Ctot, c_off, Do, Ho, Wo);
  else return set_err(B200_ERR_INVALID, "copy_channels: bad dtype");
  B200_LAUNCH_CHECK("copy_channels_kernel");
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Thin segmentation head: ConvTranspose3d(kernel 3, stride 2, padding 1, output_padding 1) from NC8 fp16 features to
// <= 4 NCDHW logit channels (the top layer of UNet, monai/networks/nets/unet.py:_get_up_layer with is_top).  With two
// output channels the GEMM is 16x padded on tensor cores and bound by the im2col gather, so this layer runs on CUDA
// cores: one thread per output voxel of ONE parity class (block-uniform live taps), 16-byte channel vectors in,
// weights broadcast from shared memory, fp32 accumulation.
namespace b200 {

struct HeadP {
  const __half* x; void* y; const float* w; const float* bias;
  int N, Cin, Cout, Di, Hi, Wi, in_ctot, in_coff, out_dtype;
};

// Thin ConvTranspose3d(k3, s2, p1, output_padding 1) head (UNet's last up layer, convolutions.py:131-152 with
// is_transposed=True): a handful of output channels, so CUDA cores.  One thread owns one INPUT cell (cz,cy,cx) and
// produces the 2x2x2 output voxels (2c + p) of every output channel from the 2x2x2 input neighbourhood {c, c+1}^3:
// along an axis, parity 0 takes (input c, tap 1) and parity 1 takes (input c, tap 2) + (input c+1, tap 0) -- 27
// (neighbour, tap) pairs in total, all resolved at compile time.  Weights sit in shared memory as [tap][cin][CO].
template <typename TO, int CO>
__global__ void __launch_bounds__(256) convt3s2_head_nc8_kernel(HeadP p) {
  extern __shared__ float s_hw[];  // [27][Cin][CO]
  for (int i = threadIdx.x; i < 27 * p.Cin * p.Cout; i += blockDim.x) {
    // ConvTranspose3d weight [Cin][Cout][3][3][3], read linearly
    const int tap = i % 27, co = (i / 27) % p.Cout, ci = i / (27 * p.Cout);
    s_hw[(tap * p.Cin + ci) * CO + co] = p.w[i];
  }
  if (p.Cout < CO)
    for (int i = threadIdx.x; i < 27 * p.Cin; i += blockDim.x)
      for (int co = p.Cout; co < CO; ++co) s_hw[i * CO + co] = 0.f;
  __syncthreads();
  const long long Si = (long long)p.Di * p.Hi * p.Wi;
  const int Ho = 2 * p.Hi, Wo = 2 * p.Wi;
  const long long So = 8 * Si;
  const int C8 = p.Cin / 8;
  float bias[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) bias[co] = (p.bias && co < p.Cout) ? p.bias[co] : 0.f;
  for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < (long long)p.N * Si; cell += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(cell / Si);
    const long long t = cell - (long long)n * Si;
    const int cx = (int)(t % p.Wi), cy = (int)((t / p.Wi) % p.Hi), cz = (int)(t / ((long long)p.Wi * p.Hi));
    const __half* xn = p.x + ((long long)n * (p.in_ctot / 8) + p.in_coff / 8) * Si * 8;
    float acc[8][CO];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int co = 0; co < CO; ++co) acc[q][co] = bias[co];
    for (int c = 0; c < C8; ++c) {
      const float* wc = s_hw + c * 8 * CO;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int iz = cz + dz, iy = cy + dy, ix = cx + dx;
            uint4 raw = make_uint4(0, 0, 0, 0);
            if (iz < p.Di && iy < p.Hi && ix < p.Wi)
              raw = __ldg(reinterpret_cast<const uint4*>(xn + ((long long)c * Si + ((long long)iz * p.Hi + iy) * p.Wi + ix) * 8));
            const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
            float v[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const float2 f2 = __half22float2(h2[j]); v[2 * j] = f2.x; v[2 * j + 1] = f2.y; }
#pragma unroll
            for (int oz = 0; oz < (dz ? 1 : 2); ++oz)
#pragma unroll
              for (int oy = 0; oy < (dy ? 1 : 2); ++oy)
#pragma unroll
                for (int ox = 0; ox < (dx ? 1 : 2); ++ox) {
                  const int pz = dz ? 1 : oz, py = dy ? 1 : oy, px = dx ? 1 : ox;
                  const int kz = dz ? 0 : (oz ? 2 : 1), ky = dy ? 0 : (oy ? 2 : 1), kx = dx ? 0 : (ox ? 2 : 1);
                  const int tap = (kz * 3 + ky) * 3 + kx, par = pz * 4 + py * 2 + px;
                  const float* wt = wc + (long long)tap * p.Cin * CO;
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (CO == 2) {
                      const float2 w2 = *reinterpret_cast<const float2*>(wt + j * 2);
                      acc[par][0] = fmaf(v[j], w2.x, acc[par][0]); acc[par][1] = fmaf(v[j], w2.y, acc[par][1]);
                    } else {
                      const float4 w4 = *reinterpret_cast<const float4*>(wt + j * 4);
                      acc[par][0] = fmaf(v[j], w4.x, acc[par][0]); acc[par][1] = fmaf(v[j], w4.y, acc[par][1]);
                      acc[par][2 % CO] = fmaf(v[j], w4.z, acc[par][2 % CO]); acc[par][3 % CO] = fmaf(v[j], w4.w, acc[par][3 % CO]);
                    }
                  }
                }
          }
    }
    TO* yn = (TO*)p.y + (long long)n * p.Cout * So;
#pragma unroll
    for (int co = 0; co < CO; ++co) {
      if (co < p.Cout) {
#pragma unroll
        for (int pz = 0; pz < 2; ++pz)
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            TO* yp = yn + (long long)co * So + ((long long)(2 * cz + pz) * Ho + (2 * cy + py)) * Wo + 2 * cx;
            io<TO>::st2(yp, acc[pz * 4 + py * 2][co], acc[pz * 4 + py * 2 + 1][co]);
          }
      }
    }
  }
}

}  // namespace b200

extern "C" int b200_convt3s2_head_nc8(const void* x, int N, int Cin, int Di, int Hi, int Wi, int in_ctot, int in_coff,
                                      const float* weight, const float* bias, int Cout, void* y, int out_dtype, void* stream) {
  B200_REQUIRE(x && weight && y, "convt3s2_head_nc8: null pointer");
  B200_REQUIRE(Cout >= 1 && Cout <= 4, "convt3s2_head_nc8: 1..4 output channels (got %d)", Cout);
  B200_REQUIRE(Cin % 8 == 0 && in_ctot % 8 == 0 && in_coff % 8 == 0 && in_coff + Cin <= in_ctot, "convt3s2_head_nc8: bad channel slice");
  B200_REQUIRE(N <= 65535, "convt3s2_head_nc8: batch too large");
  HeadP p{(const __half*)x, y, weight, bias, N, Cin, Cout, Di, Hi, Wi, in_ctot, in_coff, out_dtype};
  const long long Si = (long long)Di * Hi * Wi;
  const int CO = Cout <= 2 ? 2 : 4;
  const size_t smem = (size_t)27 * Cin * CO * sizeof(float);
  B200_REQUIRE(smem <= 96 * 1024, "convt3s2_head_nc8: Cin too large (%d)", Cin);
  const long long cells = (long long)N * Si;
  dim3 grid((unsigned)std::min<long long>(ceil_div(cells, 256), (long long)num_sms() * 8));
  cudaStream_t st = (cudaStream_t)stream;
#define LH(TO, CO_) do { \
    B200_CUDA(cudaFuncSetAttribute(convt3s2_head_nc8_kernel<TO, CO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); \
    convt3s2_head_nc8_kernel<TO, CO_><<<grid, 256, smem, st>>>(p); } while (0)
  if (out_dtype == B200_DT_F16) { if (CO == 2) LH(__half, 2); else LH(__half, 4); }
  else if (out_dtype == B200_DT_F32) { if (CO == 2) LH(float, 2); else LH(float, 4); }
  else return set_err(B200_ERR_INVALID, "convt3s2_head_nc8: bad dtype");
#undef LH
  B200_LAUNCH_CHECK("convt3s2_head_nc8_kernel");
  return B200_OK;
}
