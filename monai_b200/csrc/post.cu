// Channel-wise post-processing of the stitched logits: the step on the other side of the inferer
// in every segmentation bundle.  Replaces, for channel-first tensors [C][S] (S = flattened spatial size):
//   Activations  (monai/transforms/post/array.py:63-128):  torch.sigmoid / torch.softmax(dim=0)
//   AsDiscrete   (monai/transforms/post/array.py:131-251): torch.argmax(dim=0, keepdim=True), one_hot, `>= threshold`, torch.round
// One thread owns one voxel: the C values are read once (stride S between channels, coalesced across the warp), the
// result is written once -- HBM-bound passes.  Softmax is max-subtracted like ATen's; argmax keeps the FIRST maximum
// (torch.argmax semantics for ties); round is half-to-even.
#include "common.cuh"
#include "../../include/monai_b200.h"

namespace b200 {


template <typename TI, typename TO>
__global__ void __launch_bounds__(256) channel_softmax_kernel(const TI* __restrict__ x, TO* __restrict__ y, int C, long long S) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmaxf(m, io<TI>::ld(x + (long long)c * S + s));
  float sum = 0.f;
  for (int c = 0; c < C; ++c) sum += expf(io<TI>::ld(x + (long long)c * S + s) - m);
  const float inv = 1.f / sum;
  for (int c = 0; c < C; ++c) io<TO>::st(y + (long long)c * S + s, expf(io<TI>::ld(x + (long long)c * S + s) - m) * inv);
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) channel_sigmoid_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const float v = io<TI>::ld(x + i);
    io<TO>::st(y + i, 1.f / (1.f + expf(-v)));
  }
}

// argmax over channels; onehot == 0: y[0][s] = index as float, else y[k][s] = (k == index) for k < onehot
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) channel_argmax_kernel(const TI* __restrict__ x, TO* __restrict__ y, int C, long long S, int onehot) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float best = io<TI>::ld(x + s);
  int arg = 0;
  for (int c = 1; c < C; ++c) {
    const float v = io<TI>::ld(x + (long long)c * S + s);
    if (v > best || (v != v && best == best)) { best = v; arg = c; }   // first maximum wins; NaN is a maximum, as in torch
  }
  if (onehot <= 0) {
    io<TO>::st(y + s, (float)arg);
  } else {
    for (int k = 0; k < onehot; ++k) io<TO>::st(y + (long long)k * S + s, k == arg ? 1.f : 0.f);
  }
}

// mode 0: y = (x >= param), mode 1: y = round-half-even(x), mode 2: one-hot of an index map x[1][S] -> y[param][S]
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) elementwise_post_kernel(const TI* __restrict__ x, TO* __restrict__ y, long long total, int mode,
                                                               float param, long long S) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (mode == 0) io<TO>::st(y + i, io<TI>::ld(x + i) >= param ? 1.f : 0.f);
    else if (mode == 1) io<TO>::st(y + i, rintf(io<TI>::ld(x + i)));
    else {
      const long long s = i % S;
      const int k = (int)(i / S);
      io<TO>::st(y + i, (int)io<TI>::ld(x + s) == k ? 1.f : 0.f);
    }
  }
}

// AvgMerger (monai/inferers/merger.py:103-205): values[.., loc : loc + patch] += patch, counts[..] += 1; finalize: values /= counts.
// One thread per patch element; a patch covers each merged element at most once per call, so no atomics are needed.
template <typename TI, typename TC>
__global__ void __launch_bounds__(256) patch_accumulate_kernel(const TI* __restrict__ patch, float* __restrict__ values, TC* __restrict__ counts,
                                                               long long NC, int pd, int ph, int pw, int md, int mh, int mw, int ld, int lh,
                                                               int lw) {
  const long long pvol = (long long)pd * ph * pw, mvol = (long long)md * mh * mw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < NC * pvol; i += (long long)gridDim.x * blockDim.x) {
    const long long nc = i / pvol, r = i - nc * pvol;
    const int w = (int)(r % pw), h = (int)((r / pw) % ph), d = (int)(r / ((long long)pw * ph));
    const long long o = nc * mvol + ((long long)(ld + d) * mh + (lh + h)) * mw + (lw + w);
    values[o] += io<TI>::ld(patch + i);
    counts[o] = (TC)(counts[o] + 1);
  }
}

template <typename TC>
__global__ void __launch_bounds__(256) patch_finalize_kernel(float* __restrict__ values, const TC* __restrict__ counts, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    values[i] = __fdiv_rn(values[i], (float)counts[i]);   // 0 / 0 -> NaN for never-covered elements, as values.div_(counts) gives
}

}  // namespace b200

using namespace b200;

extern "C" int b200_patch_accumulate(const void* patch, int dtype, long long NC, int pd, int ph, int pw, float* values, void* counts,
                                     int count_bytes, int md, int mh, int mw, int ld, int lh, int lw, void* stream) {
  B200_REQUIRE(patch && values && counts, "patch_accumulate: null pointer");
  B200_REQUIRE(count_bytes == 1 || count_bytes == 4, "patch_accumulate: counts must be uint8 or int32");
  B200_REQUIRE(NC >= 0 && pd > 0 && ph > 0 && pw > 0, "patch_accumulate: bad patch shape");
  B200_REQUIRE(ld >= 0 && lh >= 0 && lw >= 0 && ld + pd <= md && lh + ph <= mh && lw + pw <= mw,
               "patch_accumulate: patch at (%d,%d,%d) of size (%d,%d,%d) leaves the merged volume (%d,%d,%d)", ld, lh, lw, pd, ph, pw, md, mh, mw);
  const long long total = NC * pd * ph * pw;
  if (total == 0) return B200_OK;
  const unsigned blocks = (unsigned)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 32);
  cudaStream_t st = (cudaStream_t)stream;
#define LPA(TI, TC) patch_accumulate_kernel<TI, TC><<<blocks, 256, 0, st>>>((const TI*)patch, values, (TC*)counts, NC, pd, ph, pw, md, mh, mw, ld, lh, lw)
  if (dtype == B200_DT_F32) { if (count_bytes == 1) LPA(float, unsigned char); else LPA(float, int); }
  else if (dtype == B200_DT_F16) { if (count_bytes == 1) LPA(__half, unsigned char); else LPA(__half, int); }
  else return set_err(B200_ERR_INVALID, "patch_accumulate: bad dtype");
#undef LPA
  B200_LAUNCH_CHECK("patch_accumulate_kernel");
  return B200_OK;
}

extern "C" int b200_patch_finalize(float* values, const void* counts, int count_bytes, long long total, void* stream) {
  B200_REQUIRE(values && counts, "patch_finalize: null pointer");
  B200_REQUIRE(count_bytes == 1 || count_bytes == 4, "patch_finalize: counts must be uint8 or int32");
  if (total <= 0) return B200_OK;
  const unsigned blocks = (unsigned)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (count_bytes == 1) patch_finalize_kernel<unsigned char><<<blocks, 256, 0, st>>>(values, (const unsigned char*)counts, total);
  else patch_finalize_kernel<int><<<blocks, 256, 0, st>>>(values, (const int*)counts, total);
  B200_LAUNCH_CHECK("patch_finalize_kernel");
  return B200_OK;
}

#define B200_POST_DISPATCH(KERNEL_CALL)                                                            \
  if (in_dtype == B200_DT_F32 && out_dtype == B200_DT_F32) { using TI = float; using TO = float; KERNEL_CALL; }   \
  else if (in_dtype == B200_DT_F16 && out_dtype == B200_DT_F32) { using TI = __half; using TO = float; KERNEL_CALL; } \
  else if (in_dtype == B200_DT_F32 && out_dtype == B200_DT_F16) { using TI = float; using TO = __half; KERNEL_CALL; } \
  else if (in_dtype == B200_DT_F16 && out_dtype == B200_DT_F16) { using TI = __half; using TO = __half; KERNEL_CALL; } \
  else return set_err(B200_ERR_INVALID, "channel_post: bad dtype");

extern "C" int b200_channel_post(const void* x, int in_dtype, int C, long long S, int op, float param, int onehot, void* y,
                                 int out_dtype, void* stream) {
  B200_REQUIRE(x && y, "channel_post: null pointer");
  B200_REQUIRE(C > 0 && S >= 0, "channel_post: bad sizes");
  B200_REQUIRE(op >= 0 && op <= 5, "channel_post: op must be 0 (softmax), 1 (sigmoid), 2 (argmax), 3 (threshold), 4 (round) or 5 (one-hot)");
  if (S == 0) return B200_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned vox_blocks = (unsigned)ceil_div(S, 256);
  const long long total = (long long)C * S;
  const unsigned flat_blocks = (unsigned)std::min<long long>(ceil_div(total, 256), (long long)num_sms() * 32);
  if (op == 0) {
    B200_POST_DISPATCH((channel_softmax_kernel<TI, TO><<<vox_blocks, 256, 0, st>>>((const TI*)x, (TO*)y, C, S)));
  } else if (op == 1) {
    B200_POST_DISPATCH((channel_sigmoid_kernel<TI, TO><<<flat_blocks, 256, 0, st>>>((const TI*)x, (TO*)y, total)));
  } else if (op == 2) {
    B200_REQUIRE(onehot >= 0, "channel_post: negative class count");
    B200_POST_DISPATCH((channel_argmax_kernel<TI, TO><<<vox_blocks, 256, 0, st>>>((const TI*)x, (TO*)y, C, S, onehot)));
  } else if (op == 3 || op == 4) {
    B200_POST_DISPATCH((elementwise_post_kernel<TI, TO><<<flat_blocks, 256, 0, st>>>((const TI*)x, (TO*)y, total, op == 3 ? 0 : 1, param, S)));
  } else {
    B200_REQUIRE(C == 1 && onehot > 0, "channel_post: one-hot takes a single-channel index map and a positive class count");
    const long long tot = (long long)onehot * S;
    const unsigned blocks = (unsigned)std::min<long long>(ceil_div(tot, 256), (long long)num_sms() * 32);
    B200_POST_DISPATCH((elementwise_post_kernel<TI, TO><<<blocks, 256, 0, st>>>((const TI*)x, (TO*)y, tot, 2, 0.f, S)));
  }
  B200_LAUNCH_CHECK("channel_post_kernel");
  return B200_OK;
}


// dst[i] += src[i] (fp32): the add of the partial numerators received from a peer rank in the depth-sharded sliding-window job
// (monai_b200/parallel/sharded.py; the reference has no multi-GPU form of sliding_window_inference to cite).
namespace b200 {
__global__ void __launch_bounds__(256) add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long n4 = n / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<float4*>(dst)[i];
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + i);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    reinterpret_cast<float4*>(dst)[i] = a;
  }
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] += src[i];
}
}  // namespace b200

extern "C" int b200_add_f32(float* dst, const float* src, long long n, void* stream) {
  B200_REQUIRE(dst && src, "add_f32: null pointer");
  B200_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "add_f32: pointers must be 16-byte aligned");
  if (n <= 0) return B200_OK;
  const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)b200::num_sms() * 8);
  b200::add_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
  B200_LAUNCH_CHECK("add_f32_kernel");
  return B200_OK;
}
