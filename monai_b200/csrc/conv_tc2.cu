// 3x3x3 stride-1 convolution on tcgen05, second generation.
//
// Same data layout, halo tiling and epilogue as conv_tc.cu, with the three changes that attack its bound (shared-
// memory operand bandwidth at small N, ncu: sm__mem_tensor_cycles_active 51 %):
//   1. the A operand is fed from TENSOR MEMORY: every shifted halo view (plane p, kh, kw) is copied shared->TMEM once
//      with tcgen05.cp (128 x 256 b) and then used by up to three MMAs (the (sub, kd) pairs with sub + kd = p), so the
//      A bytes read from shared memory per MMA drop from 4 KB to ~2 KB and the MMA itself only streams B;
//   2. the kernel is persistent (one CTA per SM loops over output tiles) with TWO accumulator sets in TMEM, so the
//      epilogue of tile i overlaps the MMAs of tile i+1;
//   3. the weight slab of a 16-channel slice (all 27 taps) arrives as one bulk copy.
// Used for N tiles <= 64 columns (the SwinUNETR 48-channel layers); wider layers are tensor-bound in conv_tc.cu.
#include "common.cuh"
#include "tc05.cuh"
#include "../../include/monai_b200.h"

namespace b200 {

constexpr int k2TH = 16, k2TW = 8, k2HH = 18, k2HW = 10;
constexpr int k2SA = 2, k2SB = 2, k2Slots = 8;

template <int NT, int BD>
struct Conv2Cfg {
  static constexpr int kPlanes = BD + 2;
  static constexpr int kChunkBytes = kPlanes * k2HH * k2HW * 16;
  static constexpr int kABytes = 2 * kChunkBytes;
  static constexpr int kBTapBytes = NT * 32;
  static constexpr int kBBytes = 27 * kBTapBytes;
  static constexpr int kAccCols = BD * NT;
  static constexpr int kSlotCol0 = 2 * kAccCols;
  static constexpr int kTmemCols = (kSlotCol0 + 8 * k2Slots <= 128) ? 128 : (kSlotCol0 + 8 * k2Slots <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = k2SA * kABytes + k2SB * kBBytes + 256 + 2 * NT * 4 + 128;
  static_assert(kSlotCol0 + 8 * k2Slots <= 512, "TMEM overflow");
};

struct Conv2Params {
  b200_conv_tc_desc d;
  const __half* w; const float* bias; __half* y; float* stats;
  int tiles_w, tiles_h, tiles_d;
};

template <int NT, int BD>
__global__ void __launch_bounds__(192, 1) conv3x3x3_tc2_kernel(const __grid_constant__ CUtensorMap tmap, Conv2Params p) {
  using Cfg = Conv2Cfg<NT, BD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + k2SA * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + k2SB * Cfg::kBBytes);
  uint64_t* full_a = bars;
  uint64_t* empty_a = bars + k2SA;
  uint64_t* full_b = bars + 2 * k2SA;
  uint64_t* empty_b = full_b + k2SB;
  uint64_t* acc_full = empty_b + k2SB;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_stats = reinterpret_cast<float*>(bars + 32);

  const b200_conv_tc_desc& d = p.d;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kc = d.Cin / 16;
  const int n_nt = d.Cout / NT;
  const long long tiles_sp = (long long)p.tiles_w * p.tiles_h * p.tiles_d;
  const long long total_tiles = tiles_sp * n_nt * d.N;

  if (threadIdx.x == 0) {
    for (int i = 0; i < k2SA; ++i) { tc::mbar_init(&full_a[i], 1); tc::mbar_init(&empty_a[i], 1); }
    for (int i = 0; i < k2SB; ++i) { tc::mbar_init(&full_b[i], 1); tc::mbar_init(&empty_b[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&acc_full[i], 1); tc::mbar_init(&acc_empty[i], 128); }
    tc::fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 2 * NT; i += blockDim.x) s_stats[i] = 0.f;
  if (warp == 1) tc::tmem_alloc(tmem_slot, Cfg::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](long long tile, int& n, int& nt, int& d0, int& h0, int& w0) {
    nt = (int)(tile % n_nt); tile /= n_nt;
    const int tw = (int)(tile % p.tiles_w); tile /= p.tiles_w;
    const int th = (int)(tile % p.tiles_h); tile /= p.tiles_h;
    const int td = (int)(tile % p.tiles_d); tile /= p.tiles_d;
    n = (int)tile; w0 = tw * k2TW; h0 = th * k2TH; d0 = td * BD;
  };

  if (warp == 0) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&tmap);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n, nt, d0, h0, w0; decode(tile, n, nt, d0, h0, w0);
        const __half* wbase = p.w + (long long)nt * num_kc * (Cfg::kBBytes / 2);
        for (int kc = 0; kc < num_kc; ++kc) {
          tc::mbar_wait(&empty_a[sa], pa ^ 1);
          tc::mbar_arrive_expect_tx(&full_a[sa], Cfg::kABytes);
          tc::tma_load_5d(smem_a + sa * Cfg::kABytes, &tmap, &full_a[sa], (w0 - 1) * 8, h0 - 1, d0 - 1, (d.in_coff + kc * 16) / 8, n);
          if (++sa == k2SA) { sa = 0; pa ^= 1; }
          tc::mbar_wait(&empty_b[sb], pb ^ 1);
          tc::mbar_arrive_expect_tx(&full_b[sb], Cfg::kBBytes);
          tc::bulk_load(smem_b + sb * Cfg::kBBytes, wbase + (long long)kc * (Cfg::kBBytes / 2), Cfg::kBBytes, &full_b[sb]);
          if (++sb == k2SB) { sb = 0; pb ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc_f16(128, NT);
      int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
      int it = 0; uint32_t slot = 0;
      for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t aph = (uint32_t)((it >> 1) & 1);
        tc::mbar_wait(&acc_empty[buf], aph ^ 1);
        tc::fence_after_sync();
        const uint32_t acc0 = tmem_base + buf * Cfg::kAccCols;
        for (int kc = 0; kc < num_kc; ++kc) {
          tc::mbar_wait(&full_a[sa], pa);
          tc::mbar_wait(&full_b[sb], pb);
          tc::fence_after_sync();
          const uint32_t a_base = tc::smem_u32(smem_a + sa * Cfg::kABytes), b_base = tc::smem_u32(smem_b + sb * Cfg::kBBytes);
#pragma unroll 1
          for (int t9 = 0; t9 < 9; ++t9) {
            const int kh = t9 / 3, kw = t9 % 3;
#pragma unroll
            for (int pl = 0; pl < BD + 2; ++pl) {
              // stage the shifted halo view (plane pl, kh, kw) in TMEM once ...
              const uint32_t a_tm = tmem_base + Cfg::kSlotCol0 + 8 * (slot & (k2Slots - 1));
              ++slot;
              tc::tmem_cp_128x256b(a_tm, tc::make_desc_kmajor_noswz(a_base + ((pl * k2HH + kh) * k2HW + kw) * 16, Cfg::kChunkBytes, k2HW * 16));
              // ... and use it for every (sub, kd) with sub + kd == pl
#pragma unroll
              for (int kd = 0; kd < 3; ++kd) {
                const int sub = pl - kd;
                if (sub < 0 || sub >= BD) continue;
                const uint64_t bdesc = tc::make_desc_kmajor_noswz(b_base + (kd * 9 + t9) * Cfg::kBTapBytes, NT * 16, 128);
                tc::mma_f16_ts(acc0 + sub * NT, a_tm, bdesc, idesc, (kc | t9 | kd) != 0 ? 1u : 0u);
              }
            }
          }
          tc::mma_commit(&empty_a[sa]);
          tc::mma_commit(&empty_b[sb]);
          if (++sa == k2SA) { sa = 0; pa ^= 1; }
          if (++sb == k2SB) { sb = 0; pb ^= 1; }
        }
        tc::mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const long long S = (long long)d.D * d.H * d.W;
    int it = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      int n, nt, d0, h0, w0; decode(tile, n, nt, d0, h0, w0);
      const int buf = it & 1;
      const uint32_t aph = (uint32_t)((it >> 1) & 1);
      const int h = h0 + (row >> 3), w = w0 + (row & 7);
      const bool hw_ok = h < d.H && w < d.W;
      const int co0 = nt * NT;
      __half* ybase = p.y + (((long long)n * (d.out_ctot / 8) + (d.out_coff + co0) / 8) * S) * 8;
      tc::mbar_wait(&acc_full[buf], aph);
      tc::fence_after_sync();
      const uint32_t tacc = tmem_base + buf * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);
      uint32_t vn[8];
      tc::tmem_ld8(tacc, vn);
#pragma unroll 1
      for (int cc = 0; cc < NT / 8; ++cc) {
        float bsum[8], bsq[8], bias8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { bsum[j] = 0.f; bsq[j] = 0.f; bias8[j] = p.bias ? p.bias[co0 + cc * 8 + j] : 0.f; }
#pragma unroll
        for (int sub = 0; sub < BD; ++sub) {
          uint32_t v[8];
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = vn[j];
          {
            const int nsub = sub + 1 < BD ? sub + 1 : 0, ncc = sub + 1 < BD ? cc : cc + 1;
            if (ncc < NT / 8) tc::tmem_ld8(tacc + nsub * NT + ncc * 8, vn);
          }
          const int dz = d0 + sub;
          const bool ok = hw_ok && dz < d.D;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = __uint_as_float(v[j]) + bias8[j];
            if (ok) { bsum[j] += f[j]; bsq[j] = fmaf(f[j], f[j], bsq[j]); }
          }
          if (ok) {
            uint4 hv;
            __half2* hp = reinterpret_cast<__half2*>(&hv);
#pragma unroll
            for (int j = 0; j < 4; ++j) hp[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
            *reinterpret_cast<uint4*>(ybase + ((long long)cc * S + ((long long)dz * d.H + h) * d.W + w) * 8) = hv;
          }
        }
        if (p.stats) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a1 = warp_sum(bsum[j]), b1 = warp_sum(bsq[j]);
            if (lane == 0) { atomicAdd(&s_stats[2 * (cc * 8 + j)], a1); atomicAdd(&s_stats[2 * (cc * 8 + j) + 1], b1); }
          }
        }
      }
      tc::fence_before_sync();
      tc::mbar_arrive(&acc_empty[buf]);
      if (p.stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int t = threadIdx.x - 64;
        for (int i = t; i < 2 * NT; i += 128) { atomicAdd(&p.stats[((long long)n * d.Cout + co0) * 2 + i], s_stats[i]); s_stats[i] = 0.f; }
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int NT, int BD>
int launch_conv_tc2(const b200_conv_tc_desc& d, const void* x, const void* w, const float* bias, void* y, float* stats, cudaStream_t st) {
  using Cfg = Conv2Cfg<NT, BD>;
  EncodeTiledFn enc = get_encode_tiled();
  B200_REQUIRE(enc != nullptr, "conv3x3x3_tc: cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tmap;
  const cuuint64_t S = (cuuint64_t)d.D * d.H * d.W;
  cuuint64_t gdim[5] = {(cuuint64_t)d.W * 8, (cuuint64_t)d.H, (cuuint64_t)d.D, (cuuint64_t)(d.in_ctot / 8), (cuuint64_t)d.N};
  cuuint64_t gstr[4] = {(cuuint64_t)d.W * 16, (cuuint64_t)d.H * d.W * 16, S * 16, S * 16 * (cuuint64_t)(d.in_ctot / 8)};
  cuuint32_t box[5] = {(cuuint32_t)k2HW * 8, (cuuint32_t)k2HH, (cuuint32_t)(BD + 2), 2, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, "conv3x3x3_tc: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  Conv2Params p;
  p.d = d; p.w = (const __half*)w; p.bias = bias; p.y = (__half*)y; p.stats = stats;
  p.tiles_w = ceil_div(d.W, k2TW); p.tiles_h = ceil_div(d.H, k2TH); p.tiles_d = ceil_div(d.D, BD);
  const long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_d * (d.Cout / NT) * d.N;
  auto kern = conv3x3x3_tc2_kernel<NT, BD>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  dim3 grid((unsigned)std::min<long long>(total, num_sms()));
  kern<<<grid, 192, Cfg::kSmemBytes, st>>>(tmap, p);
  B200_LAUNCH_CHECK("conv3x3x3_tc2_kernel");
  return B200_OK;
}

// explicit instantiations used by the dispatcher in conv_tc.cu
template int launch_conv_tc2<48, 4>(const b200_conv_tc_desc&, const void*, const void*, const float*, void*, float*, cudaStream_t);
template int launch_conv_tc2<48, 2>(const b200_conv_tc_desc&, const void*, const void*, const float*, void*, float*, cudaStream_t);
template int launch_conv_tc2<32, 4>(const b200_conv_tc_desc&, const void*, const void*, const float*, void*, float*, cudaStream_t);
template int launch_conv_tc2<16, 4>(const b200_conv_tc_desc&, const void*, const void*, const float*, void*, float*, cudaStream_t);

}  // namespace b200
