// Hyper-network fusion: log q(z | x) of a conditional flow with the hyper-network's LAST Linear
// evaluated inside the flow kernel, so the (M, D_params) parameter matrix never exists in HBM.
//
// Replaces ConditionalDensityEstimator.log_prob (torch_nf/conditional_density_estimator.py:101-104):
//     params = self.param_net(x)                     # ... Linear(H -> D_params) last (:34-37)
//     return self.density_estimator.log_prob(z, params)
// for the 'coupling' chains NormFlow builds (density_estimator.py:260-282): STAGES x [RealNVP(upper), BatchNorm,
// RealNVP(lower), BatchNorm, Affine] (+ ToInterval), one parameter row per context and one sample per context (N = 1:
// regime B, configurations C2b / C4).  Unfused, that path writes and re-reads D_params * 4 bytes per sample (5.6 KB at
// C4 against 24 B of z); fused, the kernel reads h (H floats) and z (D floats) per sample.
//
// Design.  A thread owns a sample: z, the hidden activations of both conditioner nets and the log-det stay in
// registers for the whole inverse chain.  The flow consumes its parameters in a fixed order (bijectors last to first,
// inside a bijector front to back, torch_nf/bijectors.py:224-242), so tnf_cde_pack re-lays the last Linear in that
// STREAM order, as blocks of 32 consecutive parameters: block b = [H + 1][32] floats (rows 0..H-1 the weight columns,
// row H the bias).  The CTA streams the blocks through a cp.async ring in shared memory; for each block every thread
// forms ITS 32 parameters x_i = bias_i + sum_k h[k] W[k][i] (h transposed in shared memory, W read as broadcasts, 32
// FMAs per 9 shared-memory instructions) and the fully unrolled chain code consumes them from registers - every
// parameter position is a compile-time constant.  fp32 FMA throughout: same arithmetic class as the reference.
#include "cde_common.cuh"

namespace tnf {
namespace cde {

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// The parameter stream of one thread.  Every position is a template argument: the refill of the 32-parameter register
// batch happens at the (compile-time) block boundaries only and xb is indexed statically.
struct Stream {
  static constexpr bool kFast = false;
  float xb[kBlockP];
  int H, nblk;
  const float* packed;   // global, [nblk][H + 1][32]
  float* ring;           // shared, kRing x [H + 1][32]
  const float* hs;       // shared, h transposed: [H][kThreads + 1]

  __device__ __forceinline__ void issue(int b) {
    if (b < nblk) {
      const int n16 = (H + 1) * (kBlockP / 4);
      const float4* src = reinterpret_cast<const float4*>(packed + (size_t)b * (H + 1) * kBlockP);
      float4* dst = reinterpret_cast<float4*>(ring + (size_t)(b % kRing) * (H + 1) * kBlockP);
      for (int i = threadIdx.x; i < n16; i += kThreads) cp_async16(dst + i, src + i);
    }
    cp_async_commit();   // always: the group count stays uniform
  }
  __device__ __forceinline__ void start() {
#pragma unroll
    for (int b = 0; b < kRing - 1; ++b) issue(b);
  }
  __device__ __forceinline__ void produce(int b) {
    cp_async_wait<kRing - 2>();   // block b has landed (this thread's part) ...
    __syncthreads();              // ... for every thread, and everyone is done with block b - 1
    issue(b + kRing - 1);         // into the slot of block b - 1
    const float4* w = reinterpret_cast<const float4*>(ring + (size_t)(b % kRing) * (H + 1) * kBlockP);
#pragma unroll
    for (int q = 0; q < kBlockP / 4; ++q) {
      const float4 v = w[H * (kBlockP / 4) + q];
      xb[4 * q] = v.x; xb[4 * q + 1] = v.y; xb[4 * q + 2] = v.z; xb[4 * q + 3] = v.w;
    }
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float hk = hs[k * (kThreads + 1) + threadIdx.x];
#pragma unroll
      for (int q = 0; q < kBlockP / 4; ++q) {
        const float4 v = w[k * (kBlockP / 4) + q];
        xb[4 * q] = fmaf(hk, v.x, xb[4 * q]); xb[4 * q + 1] = fmaf(hk, v.y, xb[4 * q + 1]);
        xb[4 * q + 2] = fmaf(hk, v.z, xb[4 * q + 2]); xb[4 * q + 3] = fmaf(hk, v.w, xb[4 * q + 3]);
      }
    }
  }
  template <int POS>
  __device__ __forceinline__ float get() {
    if constexpr (POS % kBlockP == 0) produce(POS / kBlockP);
    return xb[POS % kBlockP];
  }
};

template <int D, int U, int L, int STAGES>
__global__ void __launch_bounds__(kThreads) cde_logprob_kernel(ChainDesc c, const float* __restrict__ h, int H,
                                                               const float* __restrict__ packed, const float* __restrict__ z_in,
                                                               int64_t M, float* __restrict__ out_lp) {
  extern __shared__ __align__(16) float smem[];
  constexpr int P = chain_params<D, U, L, STAGES>();
  Stream S;
  S.H = H; S.nblk = (P + kBlockP - 1) / kBlockP; S.packed = packed;
  S.ring = smem;
  float* hs = smem + (size_t)kRing * (H + 1) * kBlockP;
  S.hs = hs;
  const int64_t n_tiles = (M + kThreads - 1) / kThreads;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t m0 = tile * kThreads, m = m0 + threadIdx.x;
    const bool valid = m < M;
    __syncthreads();   // the previous tile is done with hs and the ring
    S.start();
    {   // h tile, transposed (coalesced global reads, conflict-free shared stores)
      const int64_t rows = (M - m0) < kThreads ? (M - m0) : kThreads;
      const float* src = h + m0 * H;
      for (int64_t i = threadIdx.x; i < (int64_t)kThreads * H; i += kThreads) {
        const int r = (int)(i / H), k = (int)(i % H);
        hs[k * (kThreads + 1) + r] = r < rows ? src[i] : 0.f;
      }
    }
    float z[D];
#pragma unroll
    for (int d = 0; d < D; ++d) z[d] = valid ? z_in[m * D + d] : 0.f;
    const float lp = chain_logprob<D, U, L, STAGES>(z, c, S);
    if (valid) out_lp[m] = lp;
    cp_async_wait<0>();
  }
}

bool describe(const tnf_bijector_t* ch, int n, int D, ChainDesc* out, int* U, int* L, int* stages, int* sup) {
  if (n < 5) return false;
  *sup = ch[n - 1].kind == TNF_BIJ_TOINTERVAL ? 1 : 0;
  const int body = n - *sup;
  if (body % 5 != 0 || body / 5 > kMaxStages) return false;
  *stages = body / 5;
  *U = ch[0].num_units; *L = ch[0].num_layers;
  for (int s = 0; s < *stages; ++s) {
    const tnf_bijector_t* b = ch + 5 * s;
    if (b[0].kind != TNF_BIJ_REALNVP || !b[0].transform_upper || b[1].kind != TNF_BIJ_BATCHNORM ||
        b[2].kind != TNF_BIJ_REALNVP || b[2].transform_upper || b[3].kind != TNF_BIJ_BATCHNORM || b[4].kind != TNF_BIJ_AFFINE)
      return false;
    if (b[0].num_units != *U || b[2].num_units != *U || b[0].num_layers != *L || b[2].num_layers != *L) return false;
    if (out) {
      out->rnvp_off[2 * s] = b[0].param_offset; out->rnvp_off[2 * s + 1] = b[2].param_offset;
      out->aff_off[s] = b[4].param_offset;
      out->bn_mean[2 * s] = b[1].bn_mean; out->bn_alpha[2 * s] = b[1].bn_alpha; out->bn_ld[2 * s] = b[1].bn_log_det;
      out->bn_mean[2 * s + 1] = b[3].bn_mean; out->bn_alpha[2 * s + 1] = b[3].bn_alpha; out->bn_ld[2 * s + 1] = b[3].bn_log_det;
    }
  }
  if (out) out->ti_consts = *sup ? ch[n - 1].consts : nullptr;
  (void)D;
  return true;
}

bool shape_ok(int D, int U, int L, int stages) {
  if (L != 2 || stages != 1) return false;
#define X(DV, UV) if (D == DV && U == UV) return true;
  TNF_CDE_SHAPES(X)
#undef X
  return false;
}

}  // namespace cde
}  // namespace tnf

using namespace tnf;

extern "C" {

int tnf_cde_supported(const tnf_bijector_t* chain, int n_bij, int D, int H) {
  int U, L, stages, sup;
  if (!chain || H < 1 || H > 256) return 0;
  if (!cde::describe(chain, n_bij, D, nullptr, &U, &L, &stages, &sup)) return 0;
  return cde::shape_ok(D, U, L, stages) ? 1 : 0;
}

size_t tnf_cde_packed_bytes(int64_t D_params, int H, int variant) {
  if (variant == TNF_CDE_TC) return cde::tc_packed_bytes(D_params, H);
  const int64_t nblk = (D_params + cde::kBlockP - 1) / cde::kBlockP;
  return (size_t)nblk * (size_t)(H + 1) * cde::kBlockP * sizeof(float);
}

int tnf_cde_pack(const tnf_bijector_t* chain, int n_bij, int D, const float* weight, const float* bias, int H,
                 void* packed, int variant, tnf_stream_t stream) {
  TNF_REQUIRE(tnf_cde_supported(chain, n_bij, D, H), TNF_ERR_UNSUPPORTED, "tnf_cde_pack: chain / shape not supported");
  TNF_REQUIRE(weight && bias && packed, TNF_ERR_ARG, "tnf_cde_pack: null pointer");
  TNF_REQUIRE(((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN, "tnf_cde_pack: packed buffer must be 16-byte aligned");
  cde::ChainDesc c{};
  int U, L, stages, sup;
  cde::describe(chain, n_bij, D, &c, &U, &L, &stages, &sup);
  if (variant == TNF_CDE_TC) return cde::tc_pack(c, D, U, weight, bias, H, packed, (cudaStream_t)stream);
#define X(DV, UV) if (D == DV && U == UV) cde::pack_kernel<DV, UV, 2, 1><<<num_sms(), 256, 0, (cudaStream_t)stream>>>(c, weight, bias, H, (float*)packed);
  TNF_CDE_SHAPES(X)
#undef X
  return check_launch("tnf_cde_pack");
}

int tnf_cde_logprob(const tnf_bijector_t* chain, int n_bij, int D, const float* h, int H, const void* packed,
                    const float* z, int64_t M, float* log_prob, int variant, tnf_stream_t stream) {
  TNF_REQUIRE(tnf_cde_supported(chain, n_bij, D, H), TNF_ERR_UNSUPPORTED, "tnf_cde_logprob: chain / shape not supported");
  TNF_REQUIRE(M >= 0, TNF_ERR_ARG, "tnf_cde_logprob: M < 0");
  if (M == 0) return 0;
  TNF_REQUIRE(h && packed && z && log_prob, TNF_ERR_ARG, "tnf_cde_logprob: null pointer");
  TNF_REQUIRE(((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN, "tnf_cde_logprob: packed buffer must be 16-byte aligned");
  cde::ChainDesc c{};
  int U, L, stages, sup;
  cde::describe(chain, n_bij, D, &c, &U, &L, &stages, &sup);
  for (int i = 0; i < n_bij; ++i)
    TNF_REQUIRE(chain[i].kind != TNF_BIJ_BATCHNORM || (chain[i].bn_mean && chain[i].bn_alpha && chain[i].bn_log_det), TNF_ERR_ARG,
                "tnf_cde_logprob: BatchNorm state missing");
  TNF_REQUIRE(!sup || c.ti_consts, TNF_ERR_ARG, "tnf_cde_logprob: ToInterval constants missing");
  if ((variant & 15) == TNF_CDE_TC) return cde::tc_logprob(c, D, U, h, H, packed, z, M, log_prob, variant >> 8, (cudaStream_t)stream);
  const size_t smem = ((size_t)cde::kRing * (H + 1) * cde::kBlockP + (size_t)H * (cde::kThreads + 1)) * sizeof(float);
  TNF_REQUIRE(smem <= 227 * 1024, TNF_ERR_UNSUPPORTED, "tnf_cde_logprob: H = %d needs %zu B shared memory", H, smem);
  const int64_t n_tiles = (M + cde::kThreads - 1) / cde::kThreads;
  const int64_t cap = (int64_t)num_sms() * (227 * 1024 / (smem + 1024) > 8 ? 8 : 227 * 1024 / (smem + 1024));
  const int grid = (int)(n_tiles < cap ? n_tiles : cap);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define TNF_CDE_LAUNCH(DV, UV)                                                                                          \
  do {                                                                                                                  \
    e = cudaFuncSetAttribute(cde::cde_logprob_kernel<DV, UV, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) cde::cde_logprob_kernel<DV, UV, 2, 1><<<grid, cde::kThreads, smem, st>>>(c, h, H, (const float*)packed, z, M, log_prob); \
  } while (0)
#define X(DV, UV) if (D == DV && U == UV) TNF_CDE_LAUNCH(DV, UV);
  TNF_CDE_SHAPES(X)
#undef X
#undef TNF_CDE_LAUNCH
  if (e != cudaSuccess) { set_error("tnf_cde_logprob: %s", cudaGetErrorString(e)); return (int)e; }
  return check_launch("tnf_cde_logprob");
}

}  // extern "C"
