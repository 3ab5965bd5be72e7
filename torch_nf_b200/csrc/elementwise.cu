// Elementwise / row-reduction bijectors of the torch_nf chain (sm_100a).
// HBM-bound kernels: coalesced accesses, grid sized in multiples of the SM count.
//   Affine      reference torch_nf/bijectors.py:277-315
//   BatchNorm   reference torch_nf/bijectors.py:389-426
//   ToInterval  reference torch_nf/bijectors.py:509-557
//   ToSimplex   reference torch_nf/bijectors.py:574-591
//   base density / sampling  reference torch_nf/density_estimator.py:366-372,413-416
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace tnf {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static inline int grid_for(int64_t work_items, int per_block, int waves = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)num_sms() * waves;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// rounding-explicit helpers: the reference evaluates a*b and +c as two torch
// ops (two roundings); keep that so closed-form tests hold to the last ulp.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
// exp rounded from a double evaluation: matches a correctly-rounded host exp
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ double exp_cr(double x) { return exp(x); }

// ---------------------------------------------------------------- Affine
template <typename T>
__global__ void affine_kernel(const T* __restrict__ z_in, T* __restrict__ z_out, T* __restrict__ log_det,
                              const T* __restrict__ params, int64_t pstride, int64_t M, int64_t N, int D,
                              int inverse, int chunks_per_m, int64_t rows_per_chunk) {
  extern __shared__ unsigned char smem_raw[];
  T* scale = reinterpret_cast<T*>(smem_raw);
  T* shift = scale + D;
  for (int64_t blk = blockIdx.x; blk < M * chunks_per_m; blk += gridDim.x) {
    int64_t m = blk / chunks_per_m;
    int chunk = (int)(blk % chunks_per_m);
    const T* p = params + m * pstride;
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      scale[d] = exp_cr(p[d]);
      shift[d] = p[D + d];
    }
    __syncthreads();
    if (chunk == 0 && log_det != nullptr && threadIdx.x == 0) {
      T s = T(0);
      for (int d = 0; d < D; ++d) s += p[d];
      log_det[m] = s;
    }
    int64_t r0 = (int64_t)chunk * rows_per_chunk;
    int64_t r1 = r0 + rows_per_chunk < N ? r0 + rows_per_chunk : N;
    const T* zi = z_in + (m * N + r0) * D;
    T* zo = z_out + (m * N + r0) * D;
    int64_t n_el = (r1 - r0) * D;
    for (int64_t e = threadIdx.x; e < n_el; e += blockDim.x) {
      int d = (int)(e % D);
      T v = zi[e];
      zo[e] = inverse ? (v - shift[d]) / scale[d] : add_rn(mul_rn(scale[d], v), shift[d]);
    }
  }
}

// regime B (few samples per parameter row): one thread per element
template <typename T>
__global__ void affine_flat_kernel(const T* __restrict__ z_in, T* __restrict__ z_out, T* __restrict__ log_det,
                                   const T* __restrict__ params, int64_t pstride, int64_t M, int64_t N, int D,
                                   int inverse) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n_el = M * N * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride) {
    int d = (int)(e % D);
    int64_t m = e / ((int64_t)N * D);
    const T* p = params + m * pstride;
    T sc = exp_cr(p[d]), sh = p[D + d];
    T v = z_in[e];
    z_out[e] = inverse ? (v - sh) / sc : add_rn(mul_rn(sc, v), sh);
    if (log_det != nullptr && d == 0 && (e / D) % N == 0) {
      T s = T(0);
      for (int k = 0; k < D; ++k) s += p[k];
      log_det[m] = s;
    }
  }
}

// grad: forward  y = e^a z + b : g_z = g_y e^a ; g_a = sum_n g_y z e^a + g_ld ; g_b = sum_n g_y
//       inverse  y = (z-b)/e^a : g_z = g_y/e^a ; g_a = -sum_n g_y y   + g_ld ; g_b = -sum_n g_y/e^a
// one block per m; threads own columns (deterministic per-column sums over n).
template <typename T>
__global__ void affine_bwd_kernel(const T* __restrict__ z_in, const T* __restrict__ params, int64_t pstride,
                                  const T* __restrict__ g_y, const T* __restrict__ g_ld, T* __restrict__ g_z,
                                  T* __restrict__ g_params, int64_t gstride, int64_t M, int64_t N, int D,
                                  int inverse) {
  extern __shared__ unsigned char smem_raw[];
  double* red = reinterpret_cast<double*>(smem_raw);  // [2][blockDim.x]
  const int lanes_per_col = blockDim.x / D > 0 ? blockDim.x / D : 1;
  for (int64_t m = blockIdx.x; m < M; m += gridDim.x) {
    const T* p = params + m * pstride;
    T gl = g_ld ? g_ld[m] : T(0);
    for (int d0 = 0; d0 < D; d0 += blockDim.x) {
      // columns d0 .. d0+blockDim.x-1 ; when D < blockDim.x several lanes share a column
      int t = threadIdx.x;
      int d = d0 + (D < (int)blockDim.x ? t % D : t);
      int sub = D < (int)blockDim.x ? t / D : 0;
      bool active = d < D && sub < lanes_per_col;
      double ga = 0.0, gb = 0.0;
      if (active) {
        T sc = exp_cr(p[d]);
        T sh = p[D + d];
#pragma unroll 4
        for (int64_t n = sub; n < N; n += lanes_per_col) {
          int64_t e = (m * N + n) * D + d;
          T gy = g_y ? g_y[e] : T(0);
          T z = z_in[e];
          if (!inverse) {
            g_z[e] = gy * sc;
            ga += (double)(gy * z * sc);
            gb += (double)gy;
          } else {
            T y = (z - sh) / sc;
            T gz = gy / sc;
            g_z[e] = gz;
            ga -= (double)(gy * y);
            gb -= (double)gz;
          }
        }
      }
      __syncthreads();
      red[t] = ga;
      red[blockDim.x + t] = gb;
      __syncthreads();
      if (active && sub == 0) {
        double sa = 0.0, sb = 0.0;
        for (int k = 0; k < lanes_per_col; ++k) {
          int src = D < (int)blockDim.x ? k * D + (d - d0) : t;
          sa += red[src];
          sb += red[blockDim.x + src];
        }
        T* gp = g_params + (gstride ? m * gstride : 0);
        if (gstride) {
          gp[d] += (T)(sa + (double)gl);
          gp[D + d] += (T)sb;
        } else {  // one shared row: every m accumulates into it
          atomicAdd(&gp[d], (T)(sa + (double)gl));
          atomicAdd(&gp[D + d], (T)sb);
        }
      }
      __syncthreads();
    }
  }
}

// ONE shared parameter row and a large batch (regime A training), float32, D % 4 == 0: 128-bit accesses, a thread owns
// four columns and walks rows_per_iter-strided rows of its block's chunk; per-block column sums through shared memory,
// one atomicAdd per column and block into the shared gradient row.  HBM-bound (the generic kernel above: one 4-byte
// access per thread and iteration, 0.48 ms per 2^20 x 64 pass; this one: see profiles/r02_lines).
__global__ void __launch_bounds__(256) affine_bwd_shared_f32_kernel(const float* __restrict__ z_in, const float* __restrict__ params,
                                                                    const float* __restrict__ g_y, const float* __restrict__ g_ld,
                                                                    float* __restrict__ g_z, float* __restrict__ g_params,
                                                                    int64_t N, int D, int inverse, int64_t rows_per_block) {
  extern __shared__ unsigned char smem_raw[];
  float* red = reinterpret_cast<float*>(smem_raw);            // [2][rpi][D]
  const int lanes = D / 4, rpi = blockDim.x / lanes;          // threads per row, rows per iteration
  const int t = threadIdx.x, c4 = (t % lanes) * 4, rsub = t / lanes;
  const bool active = rsub < rpi;
  const int64_t n0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t n1 = n0 + rows_per_block < N ? n0 + rows_per_block : N;
  float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga;
  if (active) {
    const float4 al = *reinterpret_cast<const float4*>(params + c4), sh = *reinterpret_cast<const float4*>(params + D + c4);
    const float4 sc = make_float4(exp_cr(al.x), exp_cr(al.y), exp_cr(al.z), exp_cr(al.w));   // as the forward kernel
#pragma unroll 4
    for (int64_t n = n0 + rsub; n < n1; n += rpi) {
      const int64_t e = n * D + c4;
      const float4 gy = g_y ? *reinterpret_cast<const float4*>(g_y + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 z = *reinterpret_cast<const float4*>(z_in + e);
      float4 gz;
      if (!inverse) {
        gz = make_float4(gy.x * sc.x, gy.y * sc.y, gy.z * sc.z, gy.w * sc.w);
        ga.x = fmaf(gz.x, z.x, ga.x); ga.y = fmaf(gz.y, z.y, ga.y); ga.z = fmaf(gz.z, z.z, ga.z); ga.w = fmaf(gz.w, z.w, ga.w);
        gb.x += gy.x; gb.y += gy.y; gb.z += gy.z; gb.w += gy.w;
      } else {
        gz = make_float4(gy.x / sc.x, gy.y / sc.y, gy.z / sc.z, gy.w / sc.w);
        const float4 y = make_float4((z.x - sh.x) / sc.x, (z.y - sh.y) / sc.y, (z.z - sh.z) / sc.z, (z.w - sh.w) / sc.w);
        ga.x = fmaf(-gy.x, y.x, ga.x); ga.y = fmaf(-gy.y, y.y, ga.y); ga.z = fmaf(-gy.z, y.z, ga.z); ga.w = fmaf(-gy.w, y.w, ga.w);
        gb.x -= gz.x; gb.y -= gz.y; gb.z -= gz.z; gb.w -= gz.w;
      }
      *reinterpret_cast<float4*>(g_z + e) = gz;
    }
    *reinterpret_cast<float4*>(red + (size_t)rsub * D + c4) = ga;
    *reinterpret_cast<float4*>(red + (size_t)(rpi + rsub) * D + c4) = gb;
  }
  __syncthreads();
  for (int d = t; d < D; d += blockDim.x) {
    float sa = 0.f, sb = 0.f;
    for (int k = 0; k < rpi; ++k) { sa += red[(size_t)k * D + d]; sb += red[(size_t)(rpi + k) * D + d]; }
    if (blockIdx.x == 0 && g_ld) sa += g_ld[0];               // the Affine log-det gradient enters once
    atomicAdd(&g_params[d], sa);
    atomicAdd(&g_params[D + d], sb);
  }
}

// one sample (or a few) per parameter row - the conditional regime (N = 1): a block per m with two barriers per row is
// ~1 ms at M = 2^18; here a thread owns (m, d) and walks its N samples, no shared memory, no barriers
template <typename T>
__global__ void affine_bwd_flat_kernel(const T* __restrict__ z_in, const T* __restrict__ params, int64_t pstride,
                                       const T* __restrict__ g_y, const T* __restrict__ g_ld, T* __restrict__ g_z,
                                       T* __restrict__ g_params, int64_t gstride, int64_t M, int64_t N, int D, int inverse) {
  const int64_t total = M * D, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t m = e / D;
    const int d = (int)(e - m * D);
    const T* p = params + m * pstride;
    const T sc = exp_cr(p[d]), sh = p[D + d];
    const T gl = g_ld ? g_ld[m] : T(0);
    double ga = 0.0, gb = 0.0;
    for (int64_t n = 0; n < N; ++n) {
      const int64_t i = (m * N + n) * D + d;
      const T gy = g_y ? g_y[i] : T(0);
      const T z = z_in[i];
      if (!inverse) {
        g_z[i] = gy * sc;
        ga += (double)(gy * z * sc);
        gb += (double)gy;
      } else {
        const T y = (z - sh) / sc;
        const T gz = gy / sc;
        g_z[i] = gz;
        ga -= (double)(gy * y);
        gb -= (double)gz;
      }
    }
    T* gp = g_params + m * gstride;
    gp[d] += (T)(ga + (double)gl);
    gp[D + d] += (T)gb;
  }
}

// ---------------------------------------------------------------- column statistics
constexpr int kStatThreads = 256;
constexpr int kStatMaxBlocks = 1184;  // 8 x 148

template <typename T>
__global__ void colstats_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t rows, int D,
                                double* __restrict__ partial /* [grid][2][D] */) {
  // a: values whose column sums are wanted; if b != nullptr the second sum is sum(a*b)
  // instead of sum(a*a)  (BatchNorm backward needs sum g, sum g*y).
  __shared__ double red[2][kStatThreads];
  const int nt = blockDim.x;
  double* out = partial + (int64_t)blockIdx.x * 2 * D;
  int64_t rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
  int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  if (D <= nt) {
    const int lanes = nt / D;  // threads sharing one column
    const int t = threadIdx.x;
    const int d = t % D, sub = t / D;
    double s1 = 0.0, s2 = 0.0;
    if (sub < lanes) {
      // four independent row streams per thread: more loads in flight, shorter dependent add chains
      double p1[4] = {0.0, 0.0, 0.0, 0.0}, p2[4] = {0.0, 0.0, 0.0, 0.0};
      int64_t r = r0 + sub;
      for (; r + 3 * (int64_t)lanes < r1; r += 4 * (int64_t)lanes) {
        T va[4], vb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          va[u] = a[(r + u * (int64_t)lanes) * D + d];
          vb[u] = b ? b[(r + u * (int64_t)lanes) * D + d] : va[u];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          p1[u] += (double)va[u];
          p2[u] += (double)va[u] * (double)vb[u];
        }
      }
      for (; r < r1; r += lanes) {
        double v = (double)a[r * D + d];
        double w = b ? (double)b[r * D + d] : v;
        p1[0] += v;
        p2[0] += v * w;
      }
      s1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
      s2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
    }
    red[0][t] = s1;
    red[1][t] = s2;
    __syncthreads();
    if (t < D) {
      double t1 = 0.0, t2 = 0.0;
      for (int k = 0; k < lanes; ++k) {
        t1 += red[0][k * D + t];
        t2 += red[1][k * D + t];
      }
      out[t] = t1;
      out[D + t] = t2;
    }
  } else {
    for (int d = threadIdx.x; d < D; d += nt) {
      double s1 = 0.0, s2 = 0.0;
      for (int64_t r = r0; r < r1; ++r) {
        double v = (double)a[r * D + d];
        double w = b ? (double)b[r * D + d] : v;
        s1 += v;
        s2 += v * w;
      }
      out[d] = s1;
      out[D + d] = s2;
    }
  }
}

// one warp per output element: lanes stride over the per-block partials, fixed-order shuffle tree
__global__ void colstats_reduce_kernel(const double* __restrict__ partial, int nblocks, int D,
                                       double* __restrict__ sums, double rows) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i > 2 * D) return;
  if (i == 2 * D) {
    if (lane == 0) sums[i] = rows;
    return;
  }
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partial[(int64_t)b * 2 * D + i];
  s = warp_sum(s);
  if (lane == 0) sums[i] = s;
}

template <typename T>
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int D, double eps,
                                   T* __restrict__ mean, T* __restrict__ alpha, T* __restrict__ log_det) {
  __shared__ double red[256];
  const double n = sums[2 * D];
  double acc = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double mu = sums[d] / n;
    double var = sums[D + d] / n - mu * mu;
    if (var < 0.0) var = 0.0;
    T al = (T)sqrt(var + eps);
    mean[d] = (T)mu;
    alpha[d] = al;
    acc += (double)t_log<T>(al);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) log_det[0] = (T)(-red[0]);
}

template <typename T>
__global__ void bn_apply_kernel(const T* __restrict__ z_in, T* __restrict__ z_out, const T* __restrict__ mean,
                                const T* __restrict__ alpha, int64_t n_el, int D, int inverse) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride) {
    int d = (int)(e % D);
    T v = z_in[e];
    z_out[e] = inverse ? add_rn(mul_rn(v, alpha[d]), mean[d]) : (v - mean[d]) / alpha[d];
  }
}

template <typename T>
__global__ void bn_bwd_apply_kernel(const T* __restrict__ g_y, const T* __restrict__ y, const T* __restrict__ alpha,
                                    const double* __restrict__ gsums, const T* __restrict__ g_ld,
                                    const double* __restrict__ count, T* __restrict__ g_z, int64_t n_el, int D) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const double n = count[0];
  T gl = g_ld ? g_ld[0] : T(0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride) {
    int d = (int)(e % D);
    T mg = (T)(gsums[d] / n), mgy = (T)(gsums[D + d] / n);
    T yy = y[e];
    T gy = g_y ? g_y[e] : T(0);
    // log_det = -sum_d log alpha_d  =>  d log_det / d z = -(y/alpha)/n
    g_z[e] = (gy - mg - yy * mgy) / alpha[d] - gl * (yy / alpha[d]) / (T)n;
  }
}

// ---------------------------------------------------------------- row-group kernels
// G lanes cooperate on one row (G = 1, 8 or 32); rows are distributed grid-stride.
template <typename T> __device__ __forceinline__ T softplus_t(T x) {  // F.softplus(beta=1, threshold=20)
  return x > T(20) ? x : t_log1p<T>(t_exp<T>(x));
}
template <typename T> __device__ __forceinline__ T logsigmoid_t(T x) {
  T mn = x < T(0) ? x : T(0);
  return mn - t_log1p<T>(t_exp<T>(-t_abs<T>(x)));
}
template <typename T> __device__ __forceinline__ T sigmoid_t(T x) { return T(1) / (T(1) + t_exp<T>(-x)); }

template <typename T, int G> __device__ __forceinline__ T group_sum(T v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T> struct TiEps { static __device__ __forceinline__ T v() { return (T)1e-12; } };

template <typename T, int G>
__global__ void tointerval_kernel(const T* __restrict__ z_in, T* __restrict__ z_out, T* __restrict__ log_det,
                                  const float* __restrict__ c, int64_t rows, int D, int inverse, int accum) {
  const float *tanh_flg = c, *sp_flg = c + D, *tanh_m = c + 2 * D, *tanh_c = c + 3 * D, *sp_m = c + 4 * D,
              *sp_c = c + 5 * D, *log_m = c + 6 * D;
  const T eps = TiEps<T>::v();
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int lane = threadIdx.x % G;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;  // keep shuffles convergent
  for (int64_t r = gid; r < rows_pad; r += ngroups) {
    T ld = T(0);
    if (r < rows) {
      for (int d = lane; d < D; d += G) {
        T z = z_in[r * D + d];
        if (!inverse) {
          if (tanh_flg[d] != 0.f) {
            T th = t_tanh<T>(z);
            ld += (T)log_m[d] + t_log<T>(T(1) - th * th + eps);
            z = add_rn(mul_rn((T)tanh_m[d], th), (T)tanh_c[d]);
          } else if (sp_flg[d] != 0.f) {
            ld += logsigmoid_t<T>(z);
            z = add_rn(mul_rn((T)sp_m[d], softplus_t<T>(z)), (T)sp_c[d]);
          }
        } else {
          if (sp_flg[d] != 0.f) {
            z = t_log<T>(t_exp<T>((z - (T)sp_c[d]) / (T)sp_m[d]) - T(1) + eps);
            ld += logsigmoid_t<T>(z);
          } else if (tanh_flg[d] != 0.f) {
            T x = (z - (T)tanh_c[d]) / (T)tanh_m[d];
            z = T(0.5) * (t_log<T>(T(1) + x + eps) - t_log<T>(T(1) - x + eps));
            T th = t_tanh<T>(z);
            ld += (T)log_m[d] + t_log<T>(T(1) - th * th + eps);
          }
        }
        z_out[r * D + d] = z;
      }
    }
    ld = group_sum<T, G>(ld);
    if (r < rows && lane == 0) {
      if (accum == TNF_LD_WRITE) log_det[r] = ld;
      else if (accum == TNF_LD_ADD) log_det[r] += ld;
      else log_det[r] -= ld;
    }
  }
}

// backward w.r.t. the INPUT of the direction that was run.
//  forward : y = f(z),  ld = l(z)         g_z = g_y f'(z) + g_ld l'(z)
//  inverse : x = f^-1(z), ld = l(x)       g_z = (g_x + g_ld l'(x)) / f'(x)
template <typename T, int G>
__global__ void tointerval_bwd_kernel(const T* __restrict__ z_in, const float* __restrict__ c,
                                      const T* __restrict__ g_out, const T* __restrict__ g_ld, T* __restrict__ g_in,
                                      int64_t rows, int D, int inverse) {
  const float *tanh_flg = c, *sp_flg = c + D, *tanh_m = c + 2 * D, *tanh_c = c + 3 * D, *sp_m = c + 4 * D,
              *sp_c = c + 5 * D;
  const T eps = TiEps<T>::v();
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n_el = rows * D;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride) {
    int d = (int)(e % D);
    int64_t r = e / D;
    T z = z_in[e];
    T go = g_out ? g_out[e] : T(0);
    T gl = g_ld ? g_ld[r] : T(0);
    T g = go;
    if (tanh_flg[d] != 0.f) {
      T x = z;
      if (inverse) {
        T u = (z - (T)tanh_c[d]) / (T)tanh_m[d];
        x = T(0.5) * (t_log<T>(T(1) + u + eps) - t_log<T>(T(1) - u + eps));
      }
      T th = t_tanh<T>(x);
      T sech2 = T(1) - th * th;
      T fp = (T)tanh_m[d] * sech2;                       // f'(x)
      T lp = -T(2) * th * sech2 / (sech2 + eps);         // l'(x)
      g = inverse ? (go + gl * lp) / fp : go * fp + gl * lp;
    } else if (sp_flg[d] != 0.f) {
      T x = z;
      if (inverse) x = t_log<T>(t_exp<T>((z - (T)sp_c[d]) / (T)sp_m[d]) - T(1) + eps);
      T sg = sigmoid_t<T>(x);
      T fp = (T)sp_m[d] * sg;
      T lp = T(1) - sg;
      g = inverse ? (go + gl * lp) / fp : go * fp + gl * lp;
    }
    g_in[e] = g;
  }
}

template <typename T, int G>
__global__ void tosimplex_kernel(const T* __restrict__ z_in, T* __restrict__ z_out, T* __restrict__ log_det,
                                 int64_t rows, int Din, int Dattr, int accum) {
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int lane = threadIdx.x % G;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;
  for (int64_t r = gid; r < rows_pad; r += ngroups) {
    T se = T(0), sz = T(0);
    if (r < rows)
      for (int d = lane; d < Din; d += G) {
        T z = z_in[r * Din + d];
        se += t_exp<T>(z);
        sz += z;
      }
    se = group_sum<T, G>(se);
    sz = group_sum<T, G>(sz);
    if (r < rows) {
      T den = se + T(1);
      for (int d = lane; d < Din; d += G) z_out[r * (Din + 1) + d] = t_exp<T>(z_in[r * Din + d]) / den;
      if (lane == 0) {
        z_out[r * (Din + 1) + Din] = T(1) / den;
        T ld = t_log<T>(T(1) - (se / den) + (T)1e-10) - (T)Dattr * t_log<T>(den) + sz;
        if (accum == TNF_LD_WRITE) log_det[r] = ld;
        else if (accum == TNF_LD_ADD) log_det[r] += ld;
        else log_det[r] -= ld;
      }
    }
  }
}

// y_d = e_d/den, y_last = 1/den, den = 1 + S, S = sum e.   With G_d = g_y[d]:
//  g_z[k] = e_k/den * (G_k - sum_d G_d y_d - G_last/den) + g_ld * (1 - Dattr e_k/den - e_k/(den^2 (1-S/den+eps)))
template <typename T, int G>
__global__ void tosimplex_bwd_kernel(const T* __restrict__ z_in, const T* __restrict__ g_out,
                                     const T* __restrict__ g_ld, T* __restrict__ g_in, int64_t rows, int Din,
                                     int Dattr) {
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int lane = threadIdx.x % G;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;
  for (int64_t r = gid; r < rows_pad; r += ngroups) {
    T se = T(0), sge = T(0);
    if (r < rows)
      for (int d = lane; d < Din; d += G) {
        T e = t_exp<T>(z_in[r * Din + d]);
        se += e;
        if (g_out) sge += g_out[r * (Din + 1) + d] * e;
      }
    se = group_sum<T, G>(se);
    sge = group_sum<T, G>(sge);
    if (r < rows) {
      T den = se + T(1);
      T glast = g_out ? g_out[r * (Din + 1) + Din] : T(0);
      T gl = g_ld ? g_ld[r] : T(0);
      T dot = (sge + glast) / den;  // sum_d G_d y_d incl. the last coordinate
      T q = T(1) - (se / den) + (T)1e-10;
      for (int d = lane; d < Din; d += G) {
        T e = t_exp<T>(z_in[r * Din + d]);
        T gk = g_out ? g_out[r * (Din + 1) + d] : T(0);
        T gz = e / den * (gk - dot);
        gz += gl * (T(1) - (T)Dattr * e / den - e / (den * den * q));
        g_in[r * Din + d] = gz;
      }
    }
  }
}

template <typename T, int G>
__global__ void base_logprob_kernel(const T* __restrict__ z, const T* __restrict__ sub, const T* __restrict__ scal,
                                    int64_t scal_div, T* __restrict__ out, int64_t rows, int D) {
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int lane = threadIdx.x % G;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;
  const T cst = (T)((double)D * 0.91893853320467274178);  // D*log(sqrt(2*pi))
  for (int64_t r = gid; r < rows_pad; r += ngroups) {
    T s = T(0);
    if (r < rows)
      for (int d = lane; d < D; d += G) {
        T v = z[r * D + d];
        s -= v * v;
      }
    s = group_sum<T, G>(s);
    if (r < rows && lane == 0) {
      T lp = s / T(2) - cst;
      T ld = sub ? sub[r] : T(0);
      if (scal) ld += scal[r / scal_div];
      out[r] = (sub || scal) ? lp - ld : lp;
    }
  }
}

template <typename T>
__global__ void base_logprob_bwd_kernel(const T* __restrict__ z, const T* __restrict__ g_out, T* __restrict__ g_z,
                                        int64_t n_el, int D) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride)
    g_z[e] = -z[e] * g_out[e / D];
}

// ---------------------------------------------------------------- base sampling (Philox4x32-10)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  // u1 in (0,1], u2 in [0,1)
  float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
  float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
  // hardware lg2 / sqrt / sin / cos (relative error ~2^-22, absolute ~2^-21 for sin / cos on [-pi, pi)): the libm
  // versions cost ~100 instructions per pair and made the sampler instruction-bound at 4x its HBM time
  float r = __fsqrt_rn(-1.3862943611198906f * __log2f(u1));       // -2 ln(u1) = -2 ln2 log2(u1)
  float s, c;
  __sincosf(6.283185307179586f * (u2 - 0.5f), &s, &c);          // angle in [-pi, pi): a uniform angle either way
  return make_float2(r * c, r * s);
}

// one thread per row-quad-of-4 elements; log-density reduced per row in a second pass
__global__ void base_sample_kernel(float* __restrict__ z, int64_t n_el, uint64_t seed, uint64_t offset) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t nq = (n_el + 3) / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    uint64_t cidx = (uint64_t)q + offset;
    uint4 ctr = make_uint4((uint32_t)cidx, (uint32_t)(cidx >> 32), 0u, 0u);
    uint4 rnd = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    float2 p0 = box_muller(rnd.x, rnd.y), p1 = box_muller(rnd.z, rnd.w);
    int64_t e = q * 4;
    if (e + 3 < n_el) {
      reinterpret_cast<float4*>(z)[q] = make_float4(p0.x, p0.y, p1.x, p1.y);
    } else {
      float v[4] = {p0.x, p0.y, p1.x, p1.y};
      for (int k = 0; k < 4 && e + k < n_el; ++k) z[e + k] = v[k];
    }
  }
}

// D = 4*G with G a power of two <= 32: the G threads that draw one row also reduce its -1/2 sum z^2, so the base
// log-density costs no second pass over z.  Same Philox counters as base_sample_kernel -> identical samples.
template <int G>
__global__ void base_sample_logq_kernel(float* __restrict__ z, double* __restrict__ log_q, int64_t rows, uint64_t seed,
                                        uint64_t offset) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;    // multiple of 32, so row groups never straddle iterations
  const int64_t nq = rows * G;
  const int64_t nq_pad = (nq + stride - 1) / stride * stride;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq_pad; q += stride) {
    double s = 0.0;
    if (q < nq) {
      uint64_t cidx = (uint64_t)q + offset;
      uint4 ctr = make_uint4((uint32_t)cidx, (uint32_t)(cidx >> 32), 0u, 0u);
      uint4 rnd = philox4x32_10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      float2 p0 = box_muller(rnd.x, rnd.y), p1 = box_muller(rnd.z, rnd.w);
      reinterpret_cast<float4*>(z)[q] = make_float4(p0.x, p0.y, p1.x, p1.y);
      s = -((double)p0.x * p0.x + (double)p0.y * p0.y + (double)p1.x * p1.x + (double)p1.y * p1.y);
    }
    s = group_sum<double, G>(s);
    if (q < nq && (threadIdx.x % G) == 0) log_q[q / G] = 0.5 * s - (double)(4 * G) * 0.91893853320467274178;
  }
}

template <int G>
__global__ void base_logq_kernel(const float* __restrict__ omega, double* __restrict__ log_q, int64_t rows, int D) {
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int lane = threadIdx.x % G;
  int64_t ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t rows_pad = (rows + ngroups - 1) / ngroups * ngroups;
  for (int64_t r = gid; r < rows_pad; r += ngroups) {
    double s = 0.0;
    if (r < rows)
      for (int d = lane; d < D; d += G) {
        double v = (double)omega[r * D + d];
        s -= v * v;
      }
    s = group_sum<double, G>(s);
    if (r < rows && lane == 0) log_q[r] = 0.5 * s - (double)D * 0.91893853320467274178;
  }
}

template <typename T>
__global__ void finish_logq_kernel(double* __restrict__ log_q, const T* __restrict__ ld, const T* __restrict__ scal,
                                   int64_t scal_div, int64_t rows) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    T v = ld ? ld[r] : T(0);
    if (scal) v += scal[r / scal_div];
    log_q[r] -= (double)v;
  }
}

template <typename T>
__global__ void accum_bcast_kernel(T* __restrict__ dst, const T* __restrict__ src, int64_t n, int64_t div) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i / div];
}

__global__ void fold_colaffine_kernel(const float* __restrict__ ps_in, const float* __restrict__ pb_in, int kind,
                                      const float* __restrict__ a, const float* __restrict__ b,
                                      float* __restrict__ ps_out, float* __restrict__ pb_out,
                                      float* __restrict__ ld_accum, int D) {
  int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d == 0 && ld_accum != nullptr && kind >= TNF_FOLD_AFF_FWD) {
    float s = 0.f;
    for (int k = 0; k < D; ++k) s += a[k];
    ld_accum[0] += s;
  }
  if (d >= D) return;
  float ps = ps_in ? ps_in[d] : 1.0f, pb = pb_in ? pb_in[d] : 0.0f;
  float s, t;   // the new map z -> z*s + t
  if (kind == TNF_FOLD_BN_FWD) { s = 1.0f / b[d]; t = -a[d] / b[d]; }
  else if (kind == TNF_FOLD_BN_INV) { s = b[d]; t = a[d]; }
  else if (kind == TNF_FOLD_AFF_FWD) { s = exp_cr(a[d]); t = b[d]; }
  else { s = 1.0f / exp_cr(a[d]); t = -b[d] * s; }
  ps_out[d] = ps * s;
  pb_out[d] = fmaf(pb, s, t);
}

__global__ void colaffine_kernel(const float* __restrict__ z_in, float* __restrict__ z_out,
                                 const float* __restrict__ scale, const float* __restrict__ shift, int64_t n_el,
                                 int D) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += stride) {
    int d = (int)(e % D);
    z_out[e] = fmaf(z_in[e], scale[d], shift[d]);
  }
}
// D % 4 == 0, 16-byte aligned: one float4 per thread and step; a thread's column is fixed when the stride is a
// multiple of D/4, so its scale / shift stay in registers
__global__ void colaffine4_kernel(const float4* __restrict__ z_in, float4* __restrict__ z_out,
                                  const float* __restrict__ scale, const float* __restrict__ shift, int64_t n4, int D4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;       // multiple of D4 (launcher)
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int d = (int)(e % D4) * 4;
  const float4 sc = *reinterpret_cast<const float4*>(scale + d), sh = *reinterpret_cast<const float4*>(shift + d);
  for (; e < n4; e += stride) {
    const float4 v = __ldcs(z_in + e);
    __stcs(z_out + e, make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)));
  }
}

// launch helper for the row-group kernels
#define TNF_ROWGROUP(D, ...)                       \
  do {                                             \
    if ((D) < 16) { constexpr int G = 1; __VA_ARGS__; }      \
    else if ((D) < 64) { constexpr int G = 8; __VA_ARGS__; } \
    else { constexpr int G = 32; __VA_ARGS__; }              \
  } while (0)

static inline int rowgroup_grid(int64_t rows, int G) { return grid_for(rows * G, 256, 16); }

// ---- fused fold kernels of the chain executor (chain.cu): the same arithmetic, in the same order, as the sequences of
// bn_finalize / fold_colaffine / accum_bcast launches they replace (results are bit-identical to the unfused plan)
// Inverse direction: every BatchNorm (remembered statistics) and Affine of the chain folded into per-coupling-layer
// pre-affines with ONE launch.  Steps are in inverse execution order; kind 2 = emit the pending map for the next
// coupling layer (or the chain's end) and reset it.
__global__ void chain_fold_inv_kernel(FoldPlan plan, float* __restrict__ scal) {
  const int D = plan.D, d = threadIdx.x;
  if (d == 0) {   // scalar log-det: BatchNorm -sum log alpha (remembered), Affine sum alpha (sequential, as fold_colaffine)
    float acc = scal[0];
    for (int i = 0; i < plan.n; ++i) {
      if (plan.s[i].kind == 0) acc += plan.s[i].ld[0];
      else if (plan.s[i].kind == 1) {
        float sa = 0.f;
        for (int k = 0; k < D; ++k) sa += plan.s[i].a[k];
        acc += sa;
      }
    }
    scal[0] = acc;
  }
  if (d >= D) return;
  float ps = 1.0f, pb = 0.0f;
  for (int i = 0; i < plan.n; ++i) {
    const FoldStep& st = plan.s[i];
    if (st.kind == 2) {
      st.ps_out[d] = ps; st.pb_out[d] = pb;
      ps = 1.0f; pb = 0.0f;
    } else {
      float s_, t_;
      if (st.kind == 0) { s_ = st.b[d]; t_ = st.a[d]; }                       // BatchNorm inverse: z * alpha + mean
      else { s_ = 1.0f / exp_cr(st.a[d]); t_ = -st.b[d] * s_; }               // Affine inverse: (z - shift) / exp(alpha)
      ps = ps * s_;
      pb = fmaf(pb, s_, t_);
    }
  }
}

int chain_fold_inv_launch(const FoldPlan& plan, float* scal, cudaStream_t st) {
  chain_fold_inv_kernel<<<1, plan.D < 32 ? 32 : plan.D, 0, st>>>(plan, scal);
  return check_launch("chain_fold_inv");
}

// Sample direction, one launch per BatchNorm: statistics -> mean / alpha / log-det (bn_finalize), fold (z - mean) / alpha
// into the pending map, scal += log-det, and - when an Affine follows - fold exp(alpha) z + shift and scal += sum alpha.
// The cross-rank exchange of the statistics (tnf_peer_t): peer stores of this rank's sums into every rank's symmetric
// buffer, a system-scope release of the sequence number, an acquire spin on this rank's own counters, and the sum of
// the world's contributions in rank order.  One CTA; ~2 NVLink round trips instead of a collective launch.
struct PeerX {
  int rank, world;
  double* stats[TNF_PEER_MAX];
  unsigned long long* flags[TNF_PEER_MAX];
  unsigned long long seq;
};
__device__ __forceinline__ void peer_exchange(const PeerX& px, double* sums, int n) {
  const int par = (int)(px.seq & 1ull);
  for (int r = 0; r < px.world; ++r) {
    double* dst = px.stats[r] + ((size_t)par * px.world + px.rank) * TNF_PEER_SLOT;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sums[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < px.world) {
    unsigned long long* f = px.flags[threadIdx.x] + px.rank;            // my counter in rank threadIdx.x's flag array
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(px.seq) : "memory");
    const unsigned long long* mine = px.flags[px.rank] + threadIdx.x;   // rank threadIdx.x's counter in my array
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
    } while (v < px.seq);
  }
  __syncthreads();
  const double* src = px.stats[px.rank] + (size_t)par * px.world * TNF_PEER_SLOT;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double t = 0.0;
    for (int r = 0; r < px.world; ++r) t += src[(size_t)r * TNF_PEER_SLOT + i];
    sums[i] = t;
  }
  __syncthreads();
}

__global__ void bn_fold_fwd_kernel(double* __restrict__ sums, int D, double eps, float* __restrict__ mean,
                                   float* __restrict__ alpha, float* __restrict__ log_det, const float* __restrict__ ps_in,
                                   const float* __restrict__ pb_in, const float* __restrict__ aff, float* __restrict__ ps_out,
                                   float* __restrict__ pb_out, float* __restrict__ scal, int finalize, PeerX px) {
  __shared__ double red[256];
  if (finalize && px.world > 1) peer_exchange(px, sums, 2 * D + 1);
  if (finalize) {
    const double n = sums[2 * D];
    double acc = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      double mu = sums[d] / n;
      double var = sums[D + d] / n - mu * mu;
      if (var < 0.0) var = 0.0;
      float al = (float)sqrt(var + eps);
      mean[d] = (float)mu;
      alpha[d] = al;
      acc += (double)logf(al);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s_ = blockDim.x / 2; s_ > 0; s_ >>= 1) {
      if ((int)threadIdx.x < s_) red[threadIdx.x] += red[threadIdx.x + s_];
      __syncthreads();
    }
    if (threadIdx.x == 0) log_det[0] = (float)(-red[0]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float acc = scal[0] + log_det[0];
    if (aff != nullptr) {
      float sa = 0.f;
      for (int k = 0; k < D; ++k) sa += aff[k];
      acc += sa;
    }
    scal[0] = acc;
  }
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float ps = ps_in ? ps_in[d] : 1.0f, pb = pb_in ? pb_in[d] : 0.0f;
    const float a_ = mean[d], b_ = alpha[d];
    float s_ = 1.0f / b_, t_ = -a_ / b_;
    ps = ps * s_;
    pb = fmaf(pb, s_, t_);
    if (aff != nullptr) {
      s_ = exp_cr(aff[d]); t_ = aff[D + d];
      ps = ps * s_;
      pb = fmaf(pb, s_, t_);
    }
    ps_out[d] = ps;
    pb_out[d] = pb;
  }
}

int bn_fold_fwd_launch(double* sums, int D, double eps, float* mean, float* alpha, float* log_det, const float* ps_in,
                       const float* pb_in, const float* aff, float* ps_out, float* pb_out, float* scal, int finalize,
                       const tnf_peer_t* peer, unsigned long long seq, cudaStream_t st) {
  PeerX px{};
  px.world = 1;
  if (peer != nullptr && peer->world > 1 && finalize) {
    TNF_REQUIRE(peer->world <= TNF_PEER_MAX && peer->rank >= 0 && peer->rank < peer->world && 2 * D + 1 <= TNF_PEER_SLOT && seq >= 1,
                TNF_ERR_ARG, "bn_fold_fwd: bad peer description (world %d, rank %d, D %d)", peer->world, peer->rank, D);
    px.rank = peer->rank; px.world = peer->world; px.seq = seq;
    for (int r = 0; r < peer->world; ++r) {
      TNF_REQUIRE(peer->stats[r] && peer->flags[r], TNF_ERR_ARG, "bn_fold_fwd: peer buffer of rank %d missing", r);
      px.stats[r] = peer->stats[r]; px.flags[r] = peer->flags[r];
    }
  }
  bn_fold_fwd_kernel<<<1, 256, 0, st>>>(sums, D, eps, mean, alpha, log_det, ps_in, pb_in, aff, ps_out, pb_out, scal, finalize, px);
  return check_launch("bn_fold_fwd");
}

int colstats_reduce_launch(const double* partial, int nblocks, int D, double* sums, double rows, cudaStream_t st) {
  colstats_reduce_kernel<<<(2 * D + 1 + 7) / 8, 256, 0, st>>>(partial, nblocks, D, sums, rows);
  return check_launch("colstats_reduce");
}

}  // namespace tnf

using namespace tnf;

extern "C" {

int tnf_abi_version(void) { return TNF_ABI_VERSION; }
const char* tnf_last_error(void) { return tnf::g_err; }
int64_t tnf_launch_count(void) { return tnf::g_launches.load(std::memory_order_relaxed); }

int tnf_affine(const void* z_in, void* z_out, void* log_det, const void* params, int64_t pstride, int64_t M,
               int64_t N, int D, int direction, int dtype, tnf_stream_t stream) {
  TNF_REQUIRE(M >= 0 && N >= 0 && D >= 1, TNF_ERR_ARG, "tnf_affine: bad shape M=%lld N=%lld D=%d", (long long)M,
              (long long)N, D);
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && z_out && params, TNF_ERR_ARG, "tnf_affine: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (N <= 16) {
    TNF_DISPATCH(dtype, {
      affine_flat_kernel<T><<<grid_for(M * N * D, 256), 256, 0, st>>>((const T*)z_in, (T*)z_out, (T*)log_det,
                                                                      (const T*)params, pstride, M, N, D,
                                                                      direction == TNF_INVERSE);
    });
    return check_launch("tnf_affine");
  }
  // split each m's N rows into chunks so that M*chunks fills the machine
  int64_t want = (int64_t)num_sms() * 8;
  int64_t chunks = (want + M - 1) / M;
  int64_t max_chunks = (N * D + 4095) / 4096;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  int64_t rpc = (N + chunks - 1) / chunks;
  chunks = (N + rpc - 1) / rpc;
  int64_t nblk = M * chunks;
  int grid = (int)(nblk < want ? nblk : want);
  TNF_DISPATCH(dtype, {
    size_t smem = 2 * (size_t)D * sizeof(T);
    affine_kernel<T><<<grid, 256, smem, st>>>((const T*)z_in, (T*)z_out, (T*)log_det, (const T*)params, pstride, M,
                                               N, D, direction == TNF_INVERSE, (int)chunks, rpc);
  });
  return check_launch("tnf_affine");
}

int tnf_affine_bwd(const void* z_in, const void* params, int64_t pstride, const void* g_z_out, const void* g_log_det,
                   void* g_z_in, void* g_params, int64_t gstride, int64_t M, int64_t N, int D, int direction,
                   int dtype, tnf_stream_t stream) {
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && params && g_z_in && g_params, TNF_ERR_ARG, "tnf_affine_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (gstride != 0 && N <= 32 && M >= 4096) {   // many parameter rows with few samples each
    TNF_DISPATCH(dtype, {
      affine_bwd_flat_kernel<T><<<grid_for(M * D, 256), 256, 0, st>>>(
          (const T*)z_in, (const T*)params, pstride, (const T*)g_z_out, (const T*)g_log_det, (T*)g_z_in, (T*)g_params,
          gstride, M, N, D, direction == TNF_INVERSE);
    });
    return check_launch("tnf_affine_bwd");
  }
  int nt = 256;
  if (M == 1 && gstride == 0 && N >= 8192 && dtype == TNF_F32 && D % 4 == 0 && D / 4 <= 256 && 256 % (D / 4) == 0 &&
      ((((uintptr_t)z_in | (uintptr_t)params | (uintptr_t)g_z_out | (uintptr_t)g_z_in) & 15) == 0)) {
    const int rpi = 256 / (D / 4);
    int64_t blocks = (int64_t)num_sms() * 8;
    int64_t rpb = (N + blocks - 1) / blocks;
    rpb = (rpb + rpi - 1) / rpi * rpi;
    blocks = (N + rpb - 1) / rpb;
    affine_bwd_shared_f32_kernel<<<(int)blocks, 256, (size_t)2 * rpi * D * sizeof(float), st>>>(
        (const float*)z_in, (const float*)params, (const float*)g_z_out, (const float*)g_log_det, (float*)g_z_in,
        (float*)g_params, N, D, direction == TNF_INVERSE, rpb);
    return check_launch("tnf_affine_bwd");
  }
  if (M == 1 && gstride == 0 && N >= 8192) {
    // ONE shared parameter row and a large batch (regime A training): a single block would walk all N rows.  The rows
    // are cut into pseudo parameter rows of kRows samples that all accumulate into the shared gradient row (atomics,
    // as several m do); the remainder call adds the log-det gradient once.
    constexpr int64_t kRows = 1024;
    const int64_t Mf = N / kRows, tail = N - Mf * kRows;
    const int grid_f = (int)(Mf < (int64_t)num_sms() * 8 ? Mf : (int64_t)num_sms() * 8);
    TNF_DISPATCH(dtype, {
      const T* gy = (const T*)g_z_out;
      affine_bwd_kernel<T><<<grid_f, nt, 2 * nt * sizeof(double), st>>>(
          (const T*)z_in, (const T*)params, 0, gy, (const T*)nullptr, (T*)g_z_in, (T*)g_params, 0, Mf, kRows, D,
          direction == TNF_INVERSE);
      const int64_t off = Mf * kRows * D;
      affine_bwd_kernel<T><<<1, nt, 2 * nt * sizeof(double), st>>>(
          (const T*)z_in + off, (const T*)params, 0, gy ? gy + off : gy, (const T*)g_log_det, (T*)g_z_in + off,
          (T*)g_params, 0, 1, tail, D, direction == TNF_INVERSE);
    });
    count_launch();
    return check_launch("tnf_affine_bwd");
  }
  int grid = (int)(M < (int64_t)num_sms() * 8 ? M : (int64_t)num_sms() * 8);
  TNF_DISPATCH(dtype, {
    affine_bwd_kernel<T><<<grid, nt, 2 * nt * sizeof(double), st>>>(
        (const T*)z_in, (const T*)params, pstride, (const T*)g_z_out, (const T*)g_log_det, (T*)g_z_in, (T*)g_params,
        gstride, M, N, D, direction == TNF_INVERSE);
  });
  return check_launch("tnf_affine_bwd");
}

size_t tnf_colstats_workspace_bytes(int D) { return (size_t)kStatMaxBlocks * 2 * (size_t)D * sizeof(double); }

static int colstats_launch(const void* a, const void* b, int64_t rows, int D, double* sums, void* workspace,
                           int dtype, cudaStream_t st, const char* what) {
  TNF_REQUIRE(a && sums && workspace, TNF_ERR_ARG, "%s: null pointer", what);
  TNF_REQUIRE(rows >= 1 && D >= 1, TNF_ERR_ARG, "%s: bad shape rows=%lld D=%d", what, (long long)rows, D);
  int64_t nb = (rows + 63) / 64;
  int64_t cap = (int64_t)num_sms() * 8;
  if (cap > kStatMaxBlocks) cap = kStatMaxBlocks;
  int grid = (int)(nb < cap ? nb : cap);
  int nt = D <= kStatThreads ? (kStatThreads / D) * D : kStatThreads;
  if (nt < 32) nt = 32;
  TNF_DISPATCH(dtype, {
    colstats_kernel<T><<<grid, nt, 0, st>>>((const T*)a, (const T*)b, rows, D, (double*)workspace);
  });
  int rc = check_launch(what);
  if (rc) return rc;
  return colstats_reduce_launch((const double*)workspace, grid, D, sums, (double)rows, st);
}

int tnf_colstats(const void* z, int64_t rows, int D, double* sums, void* workspace, int dtype, tnf_stream_t stream) {
  return colstats_launch(z, nullptr, rows, D, sums, workspace, dtype, (cudaStream_t)stream, "tnf_colstats");
}

int tnf_bn_bwd_sums(const void* g_y, const void* y, int64_t rows, int D, double* gsums, void* workspace, int dtype,
                    tnf_stream_t stream) {
  TNF_REQUIRE(y, TNF_ERR_ARG, "tnf_bn_bwd_sums: null pointer");
  return colstats_launch(g_y, y, rows, D, gsums, workspace, dtype, (cudaStream_t)stream, "tnf_bn_bwd_sums");
}

int tnf_bn_finalize(const double* sums, int D, double eps, void* mean, void* alpha, void* log_det, int dtype,
                    tnf_stream_t stream) {
  TNF_REQUIRE(sums && mean && alpha && log_det, TNF_ERR_ARG, "tnf_bn_finalize: null pointer");
  TNF_REQUIRE(D >= 1, TNF_ERR_ARG, "tnf_bn_finalize: bad shape");
  TNF_DISPATCH(dtype, {
    bn_finalize_kernel<T><<<1, 256, 0, (cudaStream_t)stream>>>(sums, D, eps, (T*)mean, (T*)alpha, (T*)log_det);
  });
  return check_launch("tnf_bn_finalize");
}

int tnf_bn_apply(const void* z_in, void* z_out, const void* mean, const void* alpha, int64_t rows, int D,
                 int direction, int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && z_out && mean && alpha, TNF_ERR_ARG, "tnf_bn_apply: null pointer");
  int64_t n_el = rows * D;
  TNF_DISPATCH(dtype, {
    bn_apply_kernel<T><<<grid_for(n_el, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z_in, (T*)z_out, (const T*)mean, (const T*)alpha, n_el, D, direction == TNF_INVERSE);
  });
  return check_launch("tnf_bn_apply");
}

int tnf_bn_bwd_apply(const void* g_y, const void* y, const void* alpha, const double* gsums, const void* g_log_det,
                     const double* count, void* g_z, int64_t rows, int D, int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(y && alpha && gsums && g_z && count, TNF_ERR_ARG, "tnf_bn_bwd_apply: null pointer");
  int64_t n_el = rows * D;
  TNF_DISPATCH(dtype, {
    bn_bwd_apply_kernel<T><<<grid_for(n_el, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        (const T*)g_y, (const T*)y, (const T*)alpha, gsums, (const T*)g_log_det, count, (T*)g_z, n_el, D);
  });
  return check_launch("tnf_bn_bwd_apply");
}

int tnf_fold_colaffine(const float* ps_in, const float* pb_in, int kind, const float* a, const float* b, float* ps_out,
                       float* pb_out, float* ld_accum, int D, tnf_stream_t stream) {
  TNF_REQUIRE(a && b && ps_out && pb_out && D >= 1, TNF_ERR_ARG, "tnf_fold_colaffine: bad argument");
  TNF_REQUIRE(kind >= 0 && kind <= 3, TNF_ERR_ARG, "tnf_fold_colaffine: bad kind %d", kind);
  fold_colaffine_kernel<<<(D + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ps_in, pb_in, kind, a, b, ps_out, pb_out,
                                                                       ld_accum, D);
  return check_launch("tnf_fold_colaffine");
}

int tnf_colaffine(const float* z_in, float* z_out, const float* scale, const float* shift, int64_t rows, int D,
                  tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && z_out && scale && shift, TNF_ERR_ARG, "tnf_colaffine: null pointer");
  if (D % 4 == 0 && (256 % (D / 4)) == 0 && ((((uintptr_t)z_in | (uintptr_t)z_out | (uintptr_t)scale | (uintptr_t)shift) & 15) == 0)) {
    const int64_t n4 = rows * (D / 4);
    int blocks = grid_for(n4, 256 * 4);     // 256 threads per block: a multiple of D/4, so every thread keeps its columns
    colaffine4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)z_in, (float4*)z_out, scale, shift, n4, D / 4);
    return check_launch("tnf_colaffine");
  }
  colaffine_kernel<<<grid_for(rows * D, 256 * 4), 256, 0, (cudaStream_t)stream>>>(z_in, z_out, scale, shift, rows * D,
                                                                                  D);
  return check_launch("tnf_colaffine");
}

int tnf_tointerval(const void* z_in, void* z_out, void* log_det, const float* consts, int64_t rows, int D,
                   int direction, int accum, int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && z_out && log_det && consts, TNF_ERR_ARG, "tnf_tointerval: null pointer");
  TNF_DISPATCH(dtype, TNF_ROWGROUP(D, {
    tointerval_kernel<T, G><<<rowgroup_grid(rows, G), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z_in, (T*)z_out, (T*)log_det, consts, rows, D, direction == TNF_INVERSE, accum);
  }));
  return check_launch("tnf_tointerval");
}

int tnf_tointerval_bwd(const void* z_in, const float* consts, const void* g_z_out, const void* g_log_det,
                       void* g_z_in, int64_t rows, int D, int direction, int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && consts && g_z_in, TNF_ERR_ARG, "tnf_tointerval_bwd: null pointer");
  TNF_DISPATCH(dtype, {
    tointerval_bwd_kernel<T, 1><<<grid_for(rows * D, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z_in, consts, (const T*)g_z_out, (const T*)g_log_det, (T*)g_z_in, rows, D,
        direction == TNF_INVERSE);
  });
  return check_launch("tnf_tointerval_bwd");
}

int tnf_tosimplex(const void* z_in, void* z_out, void* log_det, int64_t rows, int D_in, int D_attr, int accum,
                  int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && z_out && log_det, TNF_ERR_ARG, "tnf_tosimplex: null pointer");
  TNF_DISPATCH(dtype, TNF_ROWGROUP(D_in, {
    tosimplex_kernel<T, G><<<rowgroup_grid(rows, G), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z_in, (T*)z_out, (T*)log_det, rows, D_in, D_attr, accum);
  }));
  return check_launch("tnf_tosimplex");
}

int tnf_tosimplex_bwd(const void* z_in, const void* g_z_out, const void* g_log_det, void* g_z_in, int64_t rows,
                      int D_in, int D_attr, int dtype, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && g_z_in, TNF_ERR_ARG, "tnf_tosimplex_bwd: null pointer");
  TNF_DISPATCH(dtype, TNF_ROWGROUP(D_in, {
    tosimplex_bwd_kernel<T, G><<<rowgroup_grid(rows, G), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z_in, (const T*)g_z_out, (const T*)g_log_det, (T*)g_z_in, rows, D_in, D_attr);
  }));
  return check_launch("tnf_tosimplex_bwd");
}

int tnf_accum_bcast(void* dst, const void* src, int64_t n_dst, int64_t div, int dtype, tnf_stream_t stream) {
  TNF_REQUIRE(dst && src && div >= 1, TNF_ERR_ARG, "tnf_accum_bcast: bad argument");
  if (n_dst == 0) return 0;
  TNF_DISPATCH(dtype, {
    accum_bcast_kernel<T><<<grid_for(n_dst, 256 * 4), 256, 0, (cudaStream_t)stream>>>((T*)dst, (const T*)src, n_dst,
                                                                                      div);
  });
  return check_launch("tnf_accum_bcast");
}

int tnf_base_logprob(const void* z, const void* sub, const void* scal, int64_t scal_div, void* out, int64_t rows,
                     int D, int dtype, tnf_stream_t stream) {
  TNF_REQUIRE(!scal || scal_div >= 1, TNF_ERR_ARG, "tnf_base_logprob: scal_div must be >= 1");
  if (rows == 0) return 0;
  TNF_REQUIRE(z && out, TNF_ERR_ARG, "tnf_base_logprob: null pointer");
  TNF_DISPATCH(dtype, TNF_ROWGROUP(D, {
    base_logprob_kernel<T, G><<<rowgroup_grid(rows, G), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z, (const T*)sub, (const T*)scal, scal_div, (T*)out, rows, D);
  }));
  return check_launch("tnf_base_logprob");
}

int tnf_base_logprob_bwd(const void* z, const void* g_out, void* g_z, int64_t rows, int D, int dtype,
                         tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(z && g_out && g_z, TNF_ERR_ARG, "tnf_base_logprob_bwd: null pointer");
  TNF_DISPATCH(dtype, {
    base_logprob_bwd_kernel<T><<<grid_for(rows * D, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        (const T*)z, (const T*)g_out, (T*)g_z, rows * D, D);
  });
  return check_launch("tnf_base_logprob_bwd");
}

int tnf_base_sample(float* z, double* log_q, int64_t rows, int D, uint64_t seed, uint64_t offset,
                    tnf_stream_t stream) {
  TNF_REQUIRE(((uintptr_t)z & 15) == 0, TNF_ERR_ALIGN, "tnf_base_sample: z must be 16-byte aligned");
  if (rows == 0) return 0;
  TNF_REQUIRE(z && log_q, TNF_ERR_ARG, "tnf_base_sample: null pointer");
  int64_t n_el = rows * D;
#define TNF_SAMPLE_FUSED(GV)                                                                                         \
  case 4 * GV:                                                                                                       \
    base_sample_logq_kernel<GV><<<grid_for(rows * GV, 256 * 2), 256, 0, (cudaStream_t)stream>>>(z, log_q, rows, seed, \
                                                                                                offset);             \
    return check_launch("tnf_base_sample");
  switch (D) {   // row width a power of two <= 128: samples and their log-density in one pass
    TNF_SAMPLE_FUSED(1) TNF_SAMPLE_FUSED(2) TNF_SAMPLE_FUSED(4) TNF_SAMPLE_FUSED(8) TNF_SAMPLE_FUSED(16) TNF_SAMPLE_FUSED(32)
    default: break;
  }
#undef TNF_SAMPLE_FUSED
  base_sample_kernel<<<grid_for((n_el + 3) / 4, 256 * 2), 256, 0, (cudaStream_t)stream>>>(z, n_el, seed, offset);
  int rc = check_launch("tnf_base_sample");
  if (rc) return rc;
  return tnf_base_logq(z, log_q, rows, D, stream);
}

int tnf_base_logq(const float* omega, double* log_q, int64_t rows, int D, tnf_stream_t stream) {
  if (rows == 0) return 0;
  TNF_REQUIRE(omega && log_q, TNF_ERR_ARG, "tnf_base_logq: null pointer");
  TNF_ROWGROUP(D, {
    base_logq_kernel<G><<<rowgroup_grid(rows, G), 256, 0, (cudaStream_t)stream>>>(omega, log_q, rows, D);
  });
  return check_launch("tnf_base_logq");
}

int tnf_finish_logq(double* log_q, const void* ld_acc, const void* scal, int64_t scal_div, int64_t rows, int dtype,
                    tnf_stream_t stream) {
  TNF_REQUIRE(!scal || scal_div >= 1, TNF_ERR_ARG, "tnf_finish_logq: scal_div must be >= 1");
  if (rows == 0) return 0;
  TNF_REQUIRE(log_q, TNF_ERR_ARG, "tnf_finish_logq: null pointer");
  TNF_DISPATCH(dtype, {
    finish_logq_kernel<T><<<grid_for(rows, 256 * 4), 256, 0, (cudaStream_t)stream>>>(
        log_q, (const T*)ld_acc, (const T*)scal, scal_div, rows);
  });
  return check_launch("tnf_finish_logq");
}

}  // extern "C"
