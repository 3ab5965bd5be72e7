// Shared pieces of the hyper-network-fusion kernels (cde_fused.cu: CUDA-core producer; cde_fused_tc.cu: tcgen05
// producer): the chain description, the stream order of the flow parameters and the per-sample inverse chain, fully
// unrolled over a parameter stream `S` whose positions are template arguments (S.template get<POS>()).
#pragma once
#include "common.cuh"

namespace tnf {
namespace cde {

constexpr int kThreads = 128;
constexpr int kBlockP = 32;      // parameters per stream block
constexpr int kRing = 4;         // cp.async ring depth (prefetch distance 3)
constexpr int kMaxStages = 8;

struct ChainDesc {   // what the kernel needs of the chain besides its compile-time shape
  int64_t rnvp_off[2 * kMaxStages];   // parameter offset of each RealNVP, chain order
  int64_t aff_off[kMaxStages];
  const float* bn_mean[2 * kMaxStages];
  const float* bn_alpha[2 * kMaxStages];
  const float* bn_ld[2 * kMaxStages];
  const float* ti_consts;
};

template <int D, int U, int L, bool UPPER>
struct Cpl {
  static constexpr int h = D / 2;
  static constexpr int d_in = UPPER ? h : D - h;
  static constexpr int d_out = D - d_in;
  static constexpr int c_off = UPPER ? 0 : h;
  static constexpr int t_off = UPPER ? h : 0;
  static constexpr int n_params = 2 * (d_in * U + d_out * U + d_out + U + (L - 1) * (U + 1) * U);
};
template <int D, int U, int L, int STAGES>
__host__ __device__ constexpr int chain_params() { return STAGES * (Cpl<D, U, L, true>::n_params + Cpl<D, U, L, false>::n_params + 2 * D); }

// stream position -> index in the reference's parameter row.  Inverse chain: stages last to first; inside a stage
// Affine, (BatchNorm), RealNVP(lower), (BatchNorm), RealNVP(upper); inside a bijector the reference's own order.
template <int D, int U, int L, int STAGES>
__host__ __device__ inline int64_t stream_to_param(int s, const ChainDesc& c) {
  constexpr int n_up = Cpl<D, U, L, true>::n_params, n_lo = Cpl<D, U, L, false>::n_params, per = n_up + n_lo + 2 * D;
  const int st = STAGES - 1 - s / per;
  int r = s % per;
  if (r < 2 * D) return c.aff_off[st] + r;
  r -= 2 * D;
  if (r < n_lo) return c.rnvp_off[2 * st + 1] + r;
  return c.rnvp_off[2 * st] + (r - n_lo);
}

template <int D, int U, int L, int STAGES>
__global__ void pack_kernel(ChainDesc c, const float* __restrict__ weight, const float* __restrict__ bias, int H,
                            float* __restrict__ packed) {
  constexpr int P = chain_params<D, U, L, STAGES>();
  constexpr int nblk = (P + kBlockP - 1) / kBlockP;
  const int64_t total = (int64_t)nblk * (H + 1) * kBlockP;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(e % kBlockP), k = (int)((e / kBlockP) % (H + 1)), b = (int)(e / ((int64_t)kBlockP * (H + 1)));
    const int s = b * kBlockP + i;
    float v = 0.f;
    if (s < P) {
      const int64_t p = stream_to_param<D, U, L, STAGES>(s, c);
      v = k < H ? weight[p * H + k] : bias[p];
    }
    packed[e] = v;
  }
}

// compile-time loop: f(IC<I>) for I = 0 .. N-1 (IC<I> converts to the constant I in constant expressions)
template <int V> struct IC {
  static constexpr int value = V;
  __host__ __device__ constexpr operator int() const { return V; }
};
template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for<I + 1, N>(f);
  }
}

// tanh / exp of the small nets.  kFast (tensor-core variant): the MUFU forms of the fp32-parity tensor-core kernel
// (coupling_tc6.cu: ex2.approx, 2 ulp, and tanh = (e - 1) / (e + 1) with an approximate reciprocal, abs error ~1e-7);
// otherwise libm's tanhf / expf as in the exact CUDA-core kernels.
template <bool FAST> __device__ __forceinline__ float cde_exp(float x) {
  if (FAST) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f)); return y; }
  return expf(x);
}
template <bool FAST> __device__ __forceinline__ float cde_tanh(float x) {
  if (FAST) {
    const float ax = fminf(fabsf(x), 10.0f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf((e - 1.0f) * r, x);
  }
  return tanhf(x);
}

// one conditioner layer of both nets at stream position POS: parameter order t_weight (K x J), s_weight, t_bias, s_bias
// (bijectors.py:224-235); 2 K J + 2 J parameters
template <int K, int J, bool ACT, int POS, typename Stream>
__device__ __forceinline__ void mlp_layer(const float (&in_t)[K], const float (&in_s)[K], float (&out_t)[J], float (&out_s)[J],
                                          Stream& S) {
#pragma unroll
  for (int j = 0; j < J; ++j) { out_t[j] = 0.f; out_s[j] = 0.f; }
  static_for<0, K>([&](auto k) {
    static_for<0, J>([&](auto j) { out_t[j] = fmaf(in_t[k], S.template get<POS + k * J + j>(), out_t[j]); });
  });
  static_for<0, K>([&](auto k) {
    static_for<0, J>([&](auto j) { out_s[j] = fmaf(in_s[k], S.template get<POS + K * J + k * J + j>(), out_s[j]); });
  });
  static_for<0, J>([&](auto j) {
    out_t[j] += S.template get<POS + 2 * K * J + j>();
    if (ACT) out_t[j] = cde_tanh<Stream::kFast>(out_t[j]);
  });
  static_for<0, J>([&](auto j) {
    out_s[j] += S.template get<POS + 2 * K * J + J + j>();
    if (ACT) out_s[j] = cde_tanh<Stream::kFast>(out_s[j]);
  });
}

// RealNVP.inverse_and_log_det (bijectors.py:183-206) at stream position POS: z2 <- (z2 - t(z1)) / exp(s(z1)), returns sum s
template <int D, int U, int L, bool UPPER, int POS, typename Stream>
__device__ __forceinline__ float coupling_inverse(float (&z)[D], Stream& S) {
  using C = Cpl<D, U, L, UPPER>;
  static_assert(L == 2, "compiled for two-layer conditioners");
  float z1[C::d_in];
#pragma unroll
  for (int k = 0; k < C::d_in; ++k) z1[k] = z[C::c_off + k];
  float ht[U], hs[U], gt[U], gs[U];
  mlp_layer<C::d_in, U, true, POS>(z1, z1, ht, hs, S);
  constexpr int P1 = POS + 2 * C::d_in * U + 2 * U;
  mlp_layer<U, U, true, P1>(ht, hs, gt, gs, S);
  constexpr int P2 = P1 + 2 * U * U + 2 * U;
  float t[C::d_out], s[C::d_out];
  mlp_layer<U, C::d_out, false, P2>(gt, gs, t, s, S);
  float ld = 0.f;
#pragma unroll
  for (int j = 0; j < C::d_out; ++j) {
    z[C::t_off + j] = Stream::kFast ? (z[C::t_off + j] - t[j]) * cde_exp<true>(-s[j]) : (z[C::t_off + j] - t[j]) / expf(s[j]);
    ld += s[j];
  }
  return ld;
}

// The whole inverse chain of one sample (density_estimator.py:393-416): z -> z0 with the parameters taken from the
// stream S in consumption order; returns log N(z0; 0, I) - sum of log-dets.
template <int D, int U, int L, int STAGES, typename Stream>
__device__ __forceinline__ float chain_logprob(float (&z)[D], const ChainDesc& c, Stream& S) {
  float ld = 0.f;
  if (c.ti_consts != nullptr) {   // ToInterval.inverse_and_log_det (bijectors.py:529-557), the arithmetic of tointerval_kernel
    const float* cc = c.ti_consts;
    const float *tanh_flg = cc, *sp_flg = cc + D, *tanh_m = cc + 2 * D, *tanh_c = cc + 3 * D, *sp_m = cc + 4 * D,
                *sp_c = cc + 5 * D, *log_m = cc + 6 * D;
    const float eps = 1e-12f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      float v = z[d];
      if (sp_flg[d] != 0.f) {
        v = logf(expf((v - sp_c[d]) / sp_m[d]) - 1.0f + eps);
        const float mn = v < 0.f ? v : 0.f;
        ld += mn - log1pf(expf(-fabsf(v)));
      } else if (tanh_flg[d] != 0.f) {
        const float x = (v - tanh_c[d]) / tanh_m[d];
        v = 0.5f * (logf(1.0f + x + eps) - logf(1.0f - x + eps));
        const float th = tanhf(v);
        ld += log_m[d] + logf(1.0f - th * th + eps);
      }
      z[d] = v;
    }
  }
  static_for<0, STAGES>([&](auto si) {
    constexpr int st = STAGES - 1 - si;
    constexpr int n_lo = Cpl<D, U, L, false>::n_params, n_up = Cpl<D, U, L, true>::n_params;
    constexpr int P0 = si * (2 * D + n_lo + n_up);      // stream position of this stage's first parameter
    {   // Affine.inverse_and_log_det (bijectors.py:297-315): params = [alpha (D), shift (D)]
      float al[D];
      static_for<0, D>([&](auto d) { al[d] = S.template get<P0 + d>(); ld += al[d]; });
      static_for<0, D>([&](auto d) { z[d] = (z[d] - S.template get<P0 + D + d>()) / expf(al[d]); });
    }
    {   // BatchNorm.inverse_and_log_det with the remembered statistics (:420-426), then the lower RealNVP
      constexpr int bi = 2 * st + 1;
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = fmaf(z[d], __ldg(c.bn_alpha[bi] + d), __ldg(c.bn_mean[bi] + d));
      ld += __ldg(c.bn_ld[bi]);
      ld += coupling_inverse<D, U, L, false, P0 + 2 * D>(z, S);
    }
    {
      constexpr int bi = 2 * st;
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = fmaf(z[d], __ldg(c.bn_alpha[bi] + d), __ldg(c.bn_mean[bi] + d));
      ld += __ldg(c.bn_ld[bi]);
      ld += coupling_inverse<D, U, L, true, P0 + 2 * D + n_lo>(z, S);
    }
  });
  float ss = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) ss = fmaf(z[d], z[d], ss);
  return (-0.5f * ss - (float)((double)D * 0.91893853320467274178)) - ld;
}

// the chain must be STAGES x [RealNVP(upper), BatchNorm, RealNVP(lower), BatchNorm, Affine] (+ ToInterval)
bool describe(const tnf_bijector_t* ch, int n, int D, ChainDesc* out, int* U, int* L, int* stages, int* sup);
bool shape_ok(int D, int U, int L, int stages);
// instantiated shapes (D, U): U = 15 is NormFlow's minimum width (density_estimator.py:344-348), what every LFI
// configuration of the reference ends up with (num_units = 2 D clamped up to 15); (8, 16) = max(15, 2 D); L = 2, one stage
#define TNF_CDE_SHAPES(X) X(2, 15) X(4, 15) X(6, 15) X(8, 15) X(8, 16)
// cde_fused_tc.cu
size_t tc_packed_bytes(int64_t D_params, int H);
int tc_pack(const ChainDesc& c, int D, int U, const float* weight, const float* bias, int H, void* packed, cudaStream_t st);
int tc_logprob(const ChainDesc& c, int D, int U, const float* h, int H, const void* packed, const float* z, int64_t M,
               float* log_prob, int dbgbits, cudaStream_t st);

}  // namespace cde
}  // namespace tnf
