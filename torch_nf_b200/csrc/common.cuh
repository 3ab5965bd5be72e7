// Shared helpers for the tnf CUDA sources (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tnf.h"

namespace tnf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int num_sms();
// sums[2D+1] = column-wise sum of partial[nblocks][2][D] (fixed order) and the row count
int colstats_reduce_launch(const double* partial, int nblocks, int D, double* sums, double rows, cudaStream_t st);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  return 0;
}

#define TNF_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      tnf::set_error(__VA_ARGS__);     \
      return (code);                   \
    }                                  \
  } while (0)

// dtype dispatch: body sees `T`
#define TNF_DISPATCH(dtype, ...)                                           \
  do {                                                                     \
    if ((dtype) == TNF_F32) { using T = float; __VA_ARGS__; }              \
    else if ((dtype) == TNF_F64) { using T = double; __VA_ARGS__; }        \
    else { tnf::set_error("bad dtype %d", (int)(dtype)); return TNF_ERR_ARG; } \
  } while (0)

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <typename T> __device__ __forceinline__ T t_tanh(T x);
template <> __device__ __forceinline__ float t_tanh<float>(float x) { return tanhf(x); }
template <> __device__ __forceinline__ double t_tanh<double>(double x) { return tanh(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T t_log1p(T x);
template <> __device__ __forceinline__ float t_log1p<float>(float x) { return log1pf(x); }
template <> __device__ __forceinline__ double t_log1p<double>(double x) { return log1p(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace tnf
