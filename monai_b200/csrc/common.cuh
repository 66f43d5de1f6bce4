// Shared helpers for the monai_b200 CUDA kernels (sm_100a only).
// Nothing here depends on torch; the C-ABI in include/monai_b200.h is plain pointers + sizes.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>

#define B200_OK 0
#define B200_ERR_INVALID 1
#define B200_ERR_CUDA 2
#define B200_ERR_UNSUPPORTED 3

#define B200_DT_F32 0
#define B200_DT_F16 1

namespace b200 {

// thread-local last error message (returned by b200_last_error()).
char* err_buf();
int set_err(int code, const char* fmt, ...);

inline int cuda_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) return set_err(B200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return B200_OK;
}

#define B200_CUDA(expr)                                            \
  do {                                                             \
    int _rc = ::b200::cuda_check((expr), #expr);                   \
    if (_rc) return _rc;                                           \
  } while (0)

#define B200_LAUNCH_CHECK(name)                                    \
  do {                                                             \
    ::b200::note_launch();                                         \
    int _rc = ::b200::cuda_check(cudaGetLastError(), name);        \
    if (_rc) return _rc;                                           \
  } while (0)

#define B200_REQUIRE(cond, ...)                                    \
  do {                                                             \
    if (!(cond)) return ::b200::set_err(B200_ERR_INVALID, __VA_ARGS__); \
  } while (0)

int num_sms();
void note_launch();

template <typename T> struct io;
template <> struct io<float> {
  __device__ static __forceinline__ float ld(const float* p) { return __ldg(p); }
  __device__ static __forceinline__ void st(float* p, float v) { *p = v; }
  __device__ static __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }  // 8-byte aligned
};
template <> struct io<__half> {
  __device__ static __forceinline__ float ld(const __half* p) { return __half2float(__ldg(p)); }
  __device__ static __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
  __device__ static __forceinline__ void st2(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }  // 4-byte aligned
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace b200
