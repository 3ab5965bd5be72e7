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

// fused fold kernels of the chain executor (elementwise.cu)
struct FoldStep {
  int kind;                 // 0 BatchNorm (a = mean, b = alpha, ld = log-det), 1 Affine (a = alpha(D), b = shift(D)), 2 emit
  const float* a; const float* b; const float* ld;
  float* ps_out; float* pb_out;
};
constexpr int kMaxFoldSteps = 96;
struct FoldPlan { int n, D; FoldStep s[kMaxFoldSteps]; };
int chain_fold_inv_launch(const FoldPlan& plan, float* scal, cudaStream_t st);
// peer != NULL (world > 1, finalize): the statistics are first exchanged over NVLink peer memory inside the kernel
// (tnf_peer_t, include/tnf.h) with sequence number `seq`; `sums` then holds the world's totals
int bn_fold_fwd_launch(double* sums, int D, double eps, float* mean, float* alpha, float* log_det, const float* ps_in,
                       const float* pb_in, const float* aff, float* ps_out, float* pb_out, float* scal, int finalize,
                       const tnf_peer_t* peer, unsigned long long seq, cudaStream_t st);

// tnf_coupling_tc with the fused base density of the chain executor: out_lp != NULL (inverse direction, TNF_LD_ADD) makes
// the layer emit log N(z_out) - log_det[row] - sum s - lp_scal[0] instead of z_out / log_det (coupling_tc.cu)
bool tc_lp_fusable(int D, int U, int L, int precision);
// the layer's kernel can emit the column statistics of its output (the next BatchNorm's batch statistics)
bool tc_stats_fusable(int D, int U, int L, int precision);
int coupling_tc_impl(const float* z_in, float* z_out, float* log_det, const void* packed, int64_t rows, int D, int U, int L,
                     int transform_upper, int direction, int accum, const float* pre_scale, const float* pre_shift,
                     double* col_stats, void* stats_workspace, int precision, int variant, void* debug, float* out_lp,
                     const float* lp_scal, tnf_stream_t stream, void* ev_after_kernel = nullptr);

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
