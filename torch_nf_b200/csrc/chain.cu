// Whole bijector chains behind one C-ABI call: tnf_chain_logprob / tnf_chain_sample.
//
// Replaces the dispatch loops of NormFlow.forward (torch_nf/density_estimator.py:374-388) and
// NormFlow.inverse_and_log_det + log_prob (:393-416): the caller describes the chain once as a POD array
// (tnf_bijector_t, chain order) and the library runs it.
//
//  * small shared-weight chains (RealNVP / BatchNorm with remembered statistics / Affine, D <= 32, U <= 64 - the
//    reference's own test configurations and the LFI toy flows, C1 / C2a): ONE kernel, chain_small_kernel.  A thread
//    owns a sample; z stays in registers for the whole chain, the chain's parameters (<= 128 KB) sit in shared memory
//    and are read as broadcasts; the only HBM traffic is z in, log_prob (or z) out.  Those configurations are
//    launch-bound (2^16 x D=8 samples are 3 us of HBM time), so the launch count is what matters.
//  * everything else: the per-bijector kernels in the order and with the folding the host-side plan used
//    (BatchNorm / Affine folded into the next tensor-core coupling kernel as a per-column pre-affine; the base
//    density fused at the end), without returning to Python between launches.
#include <math.h>

#include "common.cuh"

namespace tnf {

constexpr int kSmallMaxD = 32;
constexpr int kSmallMaxU = 64;
constexpr int kSmallMaxBij = 40;
constexpr int kSmallMaxParams = 32768;   // floats of shared memory for the chain's parameter row (128 KB)

struct SmallBij {
  int kind, L, U, upper;
  int poff;               // first parameter of this bijector inside the row
  const float* a;         // BatchNorm: remembered mean (D)
  const float* b;         // BatchNorm: remembered alpha (D)
  const float* ld;        // BatchNorm: remembered log-det (1)
  float eps;              // BatchNorm (live statistics)
};
struct SmallChain {
  int n, D, n_params, inverse;
  SmallBij b[kSmallMaxBij];
};

// conditioner of one RealNVP layer for ONE sample: in[d_in] -> t[d_out], s[d_out]; parameters in shared memory
__device__ __forceinline__ void small_conditioner(const float* __restrict__ p, const float* in, int d_in, int d_out, int U,
                                                  int L, float* ht, float* hs, float* gt, float* gs, float* t,
                                                  float* s) {
  int K = d_in;
  const float* src_t = in;
  const float* src_s = in;
  float* dst_t = ht;
  float* dst_s = hs;
  for (int l = 0; l <= L; ++l) {
    const int J = l == L ? d_out : U;
    const float* Wt = p;
    const float* Ws = Wt + K * J;
    const float* bt = Ws + K * J;
    const float* bs = bt + J;
    float* ot = l == L ? t : dst_t;
    float* os = l == L ? s : dst_s;
    for (int j = 0; j < J; ++j) {
      float at = 0.f, as = 0.f;
      for (int k = 0; k < K; ++k) {
        at = fmaf(src_t[k], Wt[k * J + j], at);
        as = fmaf(src_s[k], Ws[k * J + j], as);
      }
      at += bt[j];
      as += bs[j];
      ot[j] = l < L ? tanhf(at) : at;
      os[j] = l < L ? tanhf(as) : as;
    }
    p = bs + J;
    src_t = dst_t;
    src_s = dst_s;
    dst_t = dst_t == ht ? gt : ht;
    dst_s = dst_s == hs ? gs : hs;
    K = J;
  }
}

// inverse != 0: z -> z0 through the chain backwards, out_lp = log N(z0) - sum log-dets (density_estimator.py:393-416);
// inverse == 0 (remembered BatchNorm statistics only): omega -> z, log_q (f64) -= sum log-dets (:374-388)
__global__ void __launch_bounds__(128) chain_small_kernel(SmallChain c, const float* __restrict__ z_in,
                                                          const float* __restrict__ params, int64_t rows,
                                                          float* __restrict__ z_out, float* __restrict__ out_lp,
                                                          double* __restrict__ log_q) {
  extern __shared__ float sp[];
  for (int i = threadIdx.x; i < c.n_params; i += blockDim.x) sp[i] = params[i];
  __syncthreads();
  const int D = c.D;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    float z[kSmallMaxD];
    float ht[kSmallMaxU], hs[kSmallMaxU], gt[kSmallMaxU], gs[kSmallMaxU], t[kSmallMaxD], s[kSmallMaxD];
    for (int d = 0; d < D; ++d) z[d] = z_in[r * D + d];
    float ld = 0.f;
    for (int ii = 0; ii < c.n; ++ii) {
      const SmallBij& b = c.b[c.inverse ? c.n - 1 - ii : ii];
      if (b.kind == TNF_BIJ_REALNVP) {
        const int h = D / 2;
        const int d_in = b.upper ? h : D - h, d_out = D - d_in;
        const int c_off = b.upper ? 0 : h, t_off = b.upper ? h : 0;
        small_conditioner(sp + b.poff, z + c_off, d_in, d_out, b.U, b.L, ht, hs, gt, gs, t, s);
        for (int j = 0; j < d_out; ++j) {
          z[t_off + j] = c.inverse ? (z[t_off + j] - t[j]) / expf(s[j]) : t[j] + z[t_off + j] * expf(s[j]);
          ld += s[j];
        }
      } else if (b.kind == TNF_BIJ_BATCHNORM) {
        for (int d = 0; d < D; ++d) z[d] = c.inverse ? fmaf(z[d], b.b[d], b.a[d]) : (z[d] - b.a[d]) / b.b[d];
        ld += b.ld[0];
      } else {   // Affine: [alpha(D), shift(D)]
        const float* al = sp + b.poff;
        for (int d = 0; d < D; ++d) {
          z[d] = c.inverse ? (z[d] - al[D + d]) / expf(al[d]) : fmaf(z[d], expf(al[d]), al[D + d]);
          ld += al[d];
        }
      }
    }
    if (c.inverse) {
      float ss = 0.f;
      for (int d = 0; d < D; ++d) ss = fmaf(z[d], z[d], ss);
      out_lp[r] = -0.5f * ss - (float)D * 0.91893853320467274178f - ld;
    } else {
      log_q[r] -= (double)ld;
    }
    if (z_out != nullptr)
      for (int d = 0; d < D; ++d) z_out[r * D + d] = z[d];
  }
}

// ---- the same chain in the SAMPLE direction with LIVE BatchNorm statistics (bijectors.py:401-415): one cooperative
// launch, thread = sample (all CTAs co-resident), z in registers for the whole chain.  Each BatchNorm: per-CTA column
// sums in float64 -> global partials -> grid barrier -> every CTA adds the partials in CTA order (identical bits
// everywhere) and finalises mean / alpha / log-det with the arithmetic of bn_finalize_kernel; CTA 0 stores the state.
constexpr int kLiveMaxBn = 16;
constexpr int kLiveMaxGrid = 2400;

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(128) chain_small_live_kernel(SmallChain c, const float* __restrict__ z_in,
                                                               const float* __restrict__ params, int64_t rows,
                                                               float* __restrict__ z_out, double* __restrict__ log_q,
                                                               double* __restrict__ partial, unsigned int* __restrict__ counter) {
  extern __shared__ float sp[];
  __shared__ double s_red[4][2 * kSmallMaxD];
  __shared__ double s_part[128];
  __shared__ float s_mean[kSmallMaxD], s_alpha[kSmallMaxD], s_ld;
  for (int i = threadIdx.x; i < c.n_params; i += blockDim.x) sp[i] = params[i];
  __syncthreads();
  const int D = c.D, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = r < rows;
  float z[kSmallMaxD];
  float ht[kSmallMaxU], hs[kSmallMaxU], gt[kSmallMaxU], gs[kSmallMaxU], t[kSmallMaxD], s[kSmallMaxD];
  for (int d = 0; d < D; ++d) z[d] = valid ? z_in[r * D + d] : 0.f;
  float ld = 0.f;
  int n_bn = 0;
  for (int ii = 0; ii < c.n; ++ii) {
    const SmallBij& b = c.b[ii];
    if (b.kind == TNF_BIJ_REALNVP) {
      const int h = D / 2;
      const int d_in = b.upper ? h : D - h, d_out = D - d_in;
      const int c_off = b.upper ? 0 : h, t_off = b.upper ? h : 0;
      small_conditioner(sp + b.poff, z + c_off, d_in, d_out, b.U, b.L, ht, hs, gt, gs, t, s);
      for (int j = 0; j < d_out; ++j) {
        z[t_off + j] = t[j] + z[t_off + j] * expf(s[j]);
        ld += s[j];
      }
    } else if (b.kind == TNF_BIJ_BATCHNORM) {
      for (int d = 0; d < D; ++d) {   // per-CTA column sums, float64
        const double v = valid ? (double)z[d] : 0.0;
        const double s1 = warp_sum(v), s2 = warp_sum(v * v);
        if (lane == 0) { s_red[warp][d] = s1; s_red[warp][D + d] = s2; }
      }
      __syncthreads();
      double* mine = partial + ((size_t)n_bn * gridDim.x + blockIdx.x) * 2 * D;
      if ((int)threadIdx.x < 2 * D) mine[threadIdx.x] = ((s_red[0][threadIdx.x] + s_red[1][threadIdx.x]) + s_red[2][threadIdx.x]) + s_red[3][threadIdx.x];
      grid_barrier(counter, (unsigned int)(n_bn + 1) * gridDim.x);
      {   // all 128 threads add the CTAs' partials: thread = (slice, column), slices combined in slice order below
        const int C2 = 2 * D, S = 128 / C2;
        const double* all = partial + (size_t)n_bn * gridDim.x * C2;
        if ((int)threadIdx.x < S * C2) {
          const int col = threadIdx.x % C2, sl = threadIdx.x / C2;
          double acc = 0.0;
          for (unsigned int k = sl; k < gridDim.x; k += S) acc += __ldcg(all + (size_t)k * C2 + col);
          s_part[threadIdx.x] = acc;
        }
        __syncthreads();
      }
      if ((int)threadIdx.x < D) {
        const int C2 = 2 * D, S = 128 / C2;
        double a1 = 0.0, a2 = 0.0;
        for (int sl = 0; sl < S; ++sl) { a1 += s_part[sl * C2 + threadIdx.x]; a2 += s_part[sl * C2 + D + threadIdx.x]; }
        const double n = (double)rows, mu = a1 / n;
        double var = a2 / n - mu * mu;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mu;
        s_alpha[threadIdx.x] = (float)sqrt(var + (double)b.eps);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double acc = 0.0;
        for (int d = 0; d < D; ++d) acc += (double)logf(s_alpha[d]);
        s_ld = (float)(-acc);
      }
      __syncthreads();
      if (blockIdx.x == 0) {   // the remembered statistics (bijectors.py:412-415)
        if ((int)threadIdx.x < D) { const_cast<float*>(b.a)[threadIdx.x] = s_mean[threadIdx.x]; const_cast<float*>(b.b)[threadIdx.x] = s_alpha[threadIdx.x]; }
        if (threadIdx.x == 0) const_cast<float*>(b.ld)[0] = s_ld;
      }
      for (int d = 0; d < D; ++d) z[d] = (z[d] - s_mean[d]) / s_alpha[d];
      ld += s_ld;
      ++n_bn;
      __syncthreads();   // s_red / s_mean are reused by the next BatchNorm
    } else {   // Affine: [alpha(D), shift(D)]
      const float* al = sp + b.poff;
      for (int d = 0; d < D; ++d) {
        z[d] = fmaf(z[d], expf(al[d]), al[D + d]);
        ld += al[d];
      }
    }
  }
  if (valid) {
    log_q[r] -= (double)ld;
    for (int d = 0; d < D; ++d) z_out[r * D + d] = z[d];
  }
}

// ---------------------------------------------------------------- executor helpers
static bool small_chain_ok(const tnf_bijector_t* ch, int n, int D, int64_t Mp, int frozen_or_inverse) {
  if (Mp != 1 || D > kSmallMaxD || n > kSmallMaxBij || !frozen_or_inverse) return false;
  int64_t np = 0;
  for (int i = 0; i < n; ++i) {
    const tnf_bijector_t& b = ch[i];
    if (b.kind == TNF_BIJ_REALNVP) {
      if (b.packed != nullptr || b.num_units > kSmallMaxU || b.num_layers < 1 || b.num_layers > 5) return false;
      const int h = D / 2, d_in = b.transform_upper ? h : D - h, d_out = D - d_in, U = b.num_units, L = b.num_layers;
      np = b.param_offset + 2 * ((int64_t)d_in * U + (int64_t)d_out * U + d_out + U + (int64_t)(L - 1) * (U + 1) * U);
    } else if (b.kind == TNF_BIJ_AFFINE) {
      np = b.param_offset + 2 * D;
    } else if (b.kind == TNF_BIJ_BATCHNORM) {
      if (!b.bn_mean || !b.bn_alpha || !b.bn_log_det) return false;
    } else {
      return false;
    }
    if (np > kSmallMaxParams) return false;
  }
  return true;
}

static int launch_small(const tnf_bijector_t* ch, int n, int D, int inverse, const float* z_in, const float* params,
                        int64_t rows, float* z_out, float* out_lp, double* log_q, cudaStream_t st) {
  SmallChain c;
  c.n = n; c.D = D; c.inverse = inverse; c.n_params = 0;
  for (int i = 0; i < n; ++i) {
    const tnf_bijector_t& b = ch[i];
    c.b[i].kind = b.kind; c.b[i].L = b.num_layers; c.b[i].U = b.num_units; c.b[i].upper = b.transform_upper;
    c.b[i].poff = (int)b.param_offset;
    c.b[i].a = b.bn_mean; c.b[i].b = b.bn_alpha; c.b[i].ld = b.bn_log_det; c.b[i].eps = (float)b.bn_eps;
    int64_t end = b.param_offset;
    if (b.kind == TNF_BIJ_REALNVP) {
      const int h = D / 2, d_in = b.transform_upper ? h : D - h, d_out = D - d_in, U = b.num_units, L = b.num_layers;
      end += 2 * ((int64_t)d_in * U + (int64_t)d_out * U + d_out + U + (int64_t)(L - 1) * (U + 1) * U);
    } else if (b.kind == TNF_BIJ_AFFINE) {
      end += 2 * D;
    }
    if (end > c.n_params) c.n_params = (int)end;
  }
  const size_t smem = (size_t)c.n_params * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(chain_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tnf_chain: %s", cudaGetErrorString(e)); return (int)e; }
  }
  int64_t blocks = (rows + 127) / 128;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  chain_small_kernel<<<(int)blocks, 128, smem, st>>>(c, z_in, params, rows, z_out, out_lp, log_q);
  return check_launch("tnf_chain (fused small-chain kernel)");
}

// cooperative launch of chain_small_live_kernel, or 1 when the batch does not fit one co-resident grid
static int launch_small_live(const tnf_bijector_t* ch, int n, int D, const float* z_in, const float* params, int64_t rows,
                             float* z_out, double* log_q, double* partial, unsigned int* counter, cudaStream_t st) {
  int n_bn = 0;
  for (int i = 0; i < n; ++i) n_bn += ch[i].kind == TNF_BIJ_BATCHNORM;
  const int64_t grid = (rows + 127) / 128;
  if (n_bn > kLiveMaxBn || grid > kLiveMaxGrid) return 1;
  SmallChain c;
  c.n = n; c.D = D; c.inverse = 0; c.n_params = 0;
  for (int i = 0; i < n; ++i) {
    const tnf_bijector_t& b = ch[i];
    c.b[i].kind = b.kind; c.b[i].L = b.num_layers; c.b[i].U = b.num_units; c.b[i].upper = b.transform_upper;
    c.b[i].poff = (int)b.param_offset;
    c.b[i].a = b.bn_mean; c.b[i].b = b.bn_alpha; c.b[i].ld = b.bn_log_det; c.b[i].eps = (float)b.bn_eps;
    int64_t end = b.param_offset;
    if (b.kind == TNF_BIJ_REALNVP) {
      const int h = D / 2, d_in = b.transform_upper ? h : D - h, d_out = D - d_in, U = b.num_units, L = b.num_layers;
      end += 2 * ((int64_t)d_in * U + (int64_t)d_out * U + d_out + U + (int64_t)(L - 1) * (U + 1) * U);
    } else if (b.kind == TNF_BIJ_AFFINE) {
      end += 2 * D;
    }
    if (end > c.n_params) c.n_params = (int)end;
  }
  const size_t smem = (size_t)c.n_params * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(chain_small_live_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("tnf_chain: %s", cudaGetErrorString(e)); return (int)e; }
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chain_small_live_kernel, 128, smem) != cudaSuccess ||
      grid > (int64_t)per_sm * num_sms())
    return 1;
  cudaMemsetAsync(counter, 0, sizeof(unsigned int), st);
  void* args[] = {(void*)&c, (void*)&z_in, (void*)&params, (void*)&rows, (void*)&z_out, (void*)&log_q, (void*)&partial, (void*)&counter};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)chain_small_live_kernel, dim3((unsigned)grid), dim3(128), args, smem, st);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }   // not launchable cooperatively: the per-bijector path
  return check_launch("tnf_chain (fused small-chain kernel, live statistics)");
}

struct Ws {   // carve the caller's workspace
  unsigned char* p; size_t left;
  void* take(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes > left) return nullptr;
    void* r = p; p += bytes; left -= bytes; return r;
  }
};

static bool all_tc(const tnf_bijector_t* ch, int n) {
  bool any = false;
  for (int i = 0; i < n; ++i)
    if (ch[i].kind == TNF_BIJ_REALNVP) { if (!ch[i].packed) return false; any = true; }
  return any;
}

}  // namespace tnf

using namespace tnf;

#define TNF_TRY(call)        \
  do {                       \
    int rc_ = (call);        \
    if (rc_) return rc_;     \
  } while (0)

extern "C" {

size_t tnf_chain_workspace_bytes(int64_t M, int64_t N, int D) {
  const size_t rows = (size_t)(M * N);
  auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
  return 2 * al(rows * (size_t)(D + 1) * 4) + al(rows * 4) + al((size_t)M * 4) + 4 * al((size_t)D * 4) +
         al((size_t)(2 * D + 1) * 8) + al(tnf_colstats_workspace_bytes(D)) + 3 * al((size_t)D * 4) +
         2 * (size_t)kMaxFoldSteps * al((size_t)D * 4) +
         (D <= kSmallMaxD ? al((size_t)kLiveMaxBn * kLiveMaxGrid * 2 * (size_t)D * 8) + 256 : 0) + 1024;
}

int tnf_chain_logprob(const tnf_bijector_t* chain, int n_bij, const float* z, const float* params,
                      int64_t param_row_stride, int64_t M, int64_t N, int D, int tc_precision, float* log_prob,
                      void* workspace, size_t workspace_bytes, tnf_stream_t stream) {
  TNF_REQUIRE(chain && n_bij >= 1 && M >= 0 && N >= 0 && D >= 2, TNF_ERR_ARG, "tnf_chain_logprob: bad argument");
  const int64_t rows = M * N;
  if (rows == 0) return 0;
  TNF_REQUIRE(z && params && log_prob, TNF_ERR_ARG, "tnf_chain_logprob: null pointer");
  const int64_t Mp = param_row_stride == 0 ? 1 : M;
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n_bij; ++i)
    TNF_REQUIRE(chain[i].kind == TNF_BIJ_REALNVP || chain[i].kind == TNF_BIJ_BATCHNORM || chain[i].kind == TNF_BIJ_AFFINE ||
                    chain[i].kind == TNF_BIJ_TOINTERVAL,
                TNF_ERR_UNSUPPORTED, "tnf_chain_logprob: bijector kind %d has no inverse in a fused chain", chain[i].kind);
  if (small_chain_ok(chain, n_bij, D, Mp, 1))
    return launch_small(chain, n_bij, D, 1, z, params, rows, nullptr, log_prob, nullptr, st);

  TNF_REQUIRE(workspace && workspace_bytes >= tnf_chain_workspace_bytes(M, N, D), TNF_ERR_ARG,
              "tnf_chain_logprob: workspace of %zu bytes needed", tnf_chain_workspace_bytes(M, N, D));
  Ws ws{(unsigned char*)workspace, workspace_bytes};
  float* zb[2] = {(float*)ws.take((size_t)rows * (D + 1) * 4), (float*)ws.take((size_t)rows * (D + 1) * 4)};
  float* ld_acc = (float*)ws.take((size_t)rows * 4);
  float* scal = (float*)ws.take((size_t)Mp * 4);
  float* pend[2][2] = {{(float*)ws.take((size_t)D * 4), (float*)ws.take((size_t)D * 4)},
                       {(float*)ws.take((size_t)D * 4), (float*)ws.take((size_t)D * 4)}};
  cudaMemsetAsync(ld_acc, 0, (size_t)rows * 4, st);
  cudaMemsetAsync(scal, 0, (size_t)Mp * 4, st);
  const bool fold = all_tc(chain, n_bij) && Mp == 1;
  const int64_t Mk = Mp == 1 ? 1 : M, Nk = Mp == 1 ? rows : N;
  const float* cur = z;
  int nb = 0, np = 0;
  bool have_pend = false;
  auto out_buf = [&]() { float* o = zb[nb]; nb ^= 1; return o; };
  bool pure = fold && n_bij + 1 <= kMaxFoldSteps;
  for (int i = 0; i < n_bij && pure; ++i) pure = chain[i].kind != TNF_BIJ_TOINTERVAL;
  if (pure) {
    // Tensor-core chain of RealNVP / BatchNorm / Affine: ONE launch composes every BatchNorm / Affine into the per-column
    // pre-affine of the coupling layer executed after it (the arithmetic, in the order, of the tnf_fold_colaffine /
    // tnf_accum_bcast launches it replaces), and the last executed layer emits log N(z0) - log-dets itself (no z0
    // store, no base-density pass).
    FoldPlan plan;
    plan.n = 0; plan.D = D;
    const float* pre[kMaxFoldSteps][2];     // per bijector index: the emitted pre-affine of that coupling layer (or NULL)
    bool pending = false;
    int first_cp = -1;
    for (int i = n_bij - 1; i >= 0; --i) {
      const tnf_bijector_t& b = chain[i];
      if (b.kind == TNF_BIJ_BATCHNORM) {
        TNF_REQUIRE(b.bn_mean && b.bn_alpha && b.bn_log_det, TNF_ERR_ARG, "tnf_chain_logprob: BatchNorm state missing");
        plan.s[plan.n++] = FoldStep{0, b.bn_mean, b.bn_alpha, b.bn_log_det, nullptr, nullptr};
        pending = true;
      } else if (b.kind == TNF_BIJ_AFFINE) {
        plan.s[plan.n++] = FoldStep{1, params + b.param_offset, params + b.param_offset + D, nullptr, nullptr, nullptr};
        pending = true;
      } else {   // RealNVP
        first_cp = i;
        pre[i][0] = pre[i][1] = nullptr;
        if (pending) {
          float* ps = (float*)ws.take((size_t)D * 4);
          float* pb = (float*)ws.take((size_t)D * 4);
          plan.s[plan.n++] = FoldStep{2, nullptr, nullptr, nullptr, ps, pb};
          pending = false;
          pre[i][0] = ps; pre[i][1] = pb;
        }
      }
    }
    float* tail[2] = {nullptr, nullptr};
    if (pending) {   // BatchNorm / Affine BEFORE the first coupling layer in chain order: applied after the last executed one
      tail[0] = (float*)ws.take((size_t)D * 4); tail[1] = (float*)ws.take((size_t)D * 4);
      plan.s[plan.n++] = FoldStep{2, nullptr, nullptr, nullptr, tail[0], tail[1]};
    }
    if (plan.n > 0) TNF_TRY(chain_fold_inv_launch(plan, scal, st));
    for (int i = n_bij - 1; i >= 0; --i) {
      const tnf_bijector_t& b = chain[i];
      if (b.kind != TNF_BIJ_REALNVP) continue;
      const bool fuse_lp = i == first_cp && tail[0] == nullptr && tc_lp_fusable(D, b.num_units, b.num_layers, tc_precision);
      float* o = fuse_lp ? nullptr : out_buf();
      if (b.ev_start) cudaEventRecord((cudaEvent_t)b.ev_start, st);
      TNF_TRY(coupling_tc_impl(cur, o, ld_acc, b.packed, rows, D, b.num_units, b.num_layers, b.transform_upper, TNF_INVERSE,
                               TNF_LD_ADD, pre[i][0], pre[i][1], nullptr, nullptr, tc_precision, 0, nullptr,
                               fuse_lp ? log_prob : nullptr, fuse_lp ? scal : nullptr, stream));
      if (b.ev_stop) cudaEventRecord((cudaEvent_t)b.ev_stop, st);
      if (fuse_lp) return 0;
      cur = o;
    }
    if (tail[0]) {
      float* o = out_buf();
      TNF_TRY(tnf_colaffine(cur, o, tail[0], tail[1], rows, D, stream));
      cur = o;
    }
    return tnf_base_logprob(cur, ld_acc, scal, rows, log_prob, rows, D, TNF_F32, stream);
  }
  auto flush_pend = [&]() -> int {
    if (!have_pend) return 0;
    float* o = out_buf();
    TNF_TRY(tnf_colaffine(cur, o, pend[np ^ 1][0], pend[np ^ 1][1], rows, D, stream));
    cur = o; have_pend = false;
    return 0;
  };
  for (int i = n_bij - 1; i >= 0; --i) {
    const tnf_bijector_t& b = chain[i];
    const float* prm = params + b.param_offset;
    if (b.kind == TNF_BIJ_REALNVP) {
      float* o = out_buf();
      if (b.packed) {
        if (b.ev_start) cudaEventRecord((cudaEvent_t)b.ev_start, st);
        TNF_TRY(tnf_coupling_tc(cur, o, ld_acc, b.packed, rows, D, b.num_units, b.num_layers, b.transform_upper, TNF_INVERSE,
                                TNF_LD_ADD, have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr, nullptr,
                                nullptr, tc_precision, 0, nullptr, stream));
        if (b.ev_stop) cudaEventRecord((cudaEvent_t)b.ev_stop, st);
        have_pend = false;
      } else {
        TNF_TRY(flush_pend());
        TNF_TRY(tnf_coupling(cur, o, ld_acc, prm, param_row_stride, Mk, Nk, D, b.num_units, b.num_layers, b.transform_upper,
                             TNF_INVERSE, TNF_LD_ADD, TNF_F32, stream));
      }
      cur = o;
    } else if (b.kind == TNF_BIJ_BATCHNORM) {
      TNF_REQUIRE(b.bn_mean && b.bn_alpha && b.bn_log_det, TNF_ERR_ARG, "tnf_chain_logprob: BatchNorm state missing");
      if (fold) {
        TNF_TRY(tnf_fold_colaffine(have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr, TNF_FOLD_BN_INV,
                                   b.bn_mean, b.bn_alpha, pend[np][0], pend[np][1], nullptr, D, stream));
        np ^= 1; have_pend = true;
      } else {
        float* o = out_buf();
        TNF_TRY(tnf_bn_apply(cur, o, b.bn_mean, b.bn_alpha, rows, D, TNF_INVERSE, TNF_F32, stream));
        cur = o;
      }
      TNF_TRY(tnf_accum_bcast(scal, b.bn_log_det, Mp, Mp, TNF_F32, stream));
    } else if (b.kind == TNF_BIJ_AFFINE) {
      if (fold) {
        TNF_TRY(tnf_fold_colaffine(have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr, TNF_FOLD_AFF_INV,
                                   prm, prm + D, pend[np][0], pend[np][1], scal, D, stream));
        np ^= 1; have_pend = true;
      } else {
        float* o = out_buf();
        float* ldm = o + (size_t)rows * D;   // (Mp) log-dets: each z buffer has `rows` extra floats at its tail
        TNF_TRY(tnf_affine(cur, o, ldm, prm, param_row_stride, Mk, Nk, D, TNF_INVERSE, TNF_F32, stream));
        TNF_TRY(tnf_accum_bcast(scal, ldm, Mp, 1, TNF_F32, stream));
        cur = o;
      }
    } else {   // ToInterval
      TNF_TRY(flush_pend());
      float* o = out_buf();
      TNF_TRY(tnf_tointerval(cur, o, ld_acc, b.consts, rows, D, TNF_INVERSE, TNF_LD_ADD, TNF_F32, stream));
      cur = o;
    }
  }
  TNF_TRY(flush_pend());
  return tnf_base_logprob(cur, ld_acc, scal, Mp == M && M > 1 ? N : rows, log_prob, rows, D, TNF_F32, stream);
}

int tnf_chain_sample(const tnf_bijector_t* chain, int n_bij, const float* params, int64_t param_row_stride, int64_t M,
                     int64_t N, int D, int tc_precision, const float* omega, uint64_t seed, uint64_t offset, int freeze_bn,
                     tnf_allreduce_fn allreduce, void* allreduce_user, double* stats_buf, const tnf_peer_t* peer, float* z_out,
                     double* log_q, void* workspace, size_t workspace_bytes, tnf_stream_t stream) {
  TNF_REQUIRE(chain && n_bij >= 1 && M >= 0 && N >= 0 && D >= 2, TNF_ERR_ARG, "tnf_chain_sample: bad argument");
  const int64_t rows = M * N;
  if (rows == 0) return 0;
  TNF_REQUIRE(params && z_out && log_q, TNF_ERR_ARG, "tnf_chain_sample: null pointer");
  const int64_t Mp = param_row_stride == 0 ? 1 : M;
  cudaStream_t st = (cudaStream_t)stream;
  int D_out = D;
  for (int i = 0; i < n_bij; ++i) {
    TNF_REQUIRE(chain[i].kind >= TNF_BIJ_REALNVP && chain[i].kind <= TNF_BIJ_TOSIMPLEX, TNF_ERR_UNSUPPORTED,
                "tnf_chain_sample: bijector kind %d", chain[i].kind);
    TNF_REQUIRE(chain[i].kind != TNF_BIJ_TOSIMPLEX || i == n_bij - 1, TNF_ERR_UNSUPPORTED,
                "tnf_chain_sample: ToSimplex must be the last bijector");
    if (chain[i].kind == TNF_BIJ_TOSIMPLEX) D_out = D + 1;
  }
  TNF_REQUIRE(workspace && workspace_bytes >= tnf_chain_workspace_bytes(M, N, D), TNF_ERR_ARG,
              "tnf_chain_sample: workspace of %zu bytes needed", tnf_chain_workspace_bytes(M, N, D));
  Ws ws{(unsigned char*)workspace, workspace_bytes};
  float* zb[2] = {(float*)ws.take((size_t)rows * (D + 1) * 4), (float*)ws.take((size_t)rows * (D + 1) * 4)};
  float* ld_acc = (float*)ws.take((size_t)rows * 4);
  float* scal = (float*)ws.take((size_t)Mp * 4);
  float* pend[2][2] = {{(float*)ws.take((size_t)D * 4), (float*)ws.take((size_t)D * 4)},
                       {(float*)ws.take((size_t)D * 4), (float*)ws.take((size_t)D * 4)}};
  double* sums_own = (double*)ws.take((size_t)(2 * D + 1) * 8);
  void* stat_ws = ws.take(tnf_colstats_workspace_bytes(D));
  double* sums = stats_buf ? stats_buf : sums_own;

  // base draw (density_estimator.py:366-372): device Philox, or the injected omega
  float* z0 = zb[0];
  if (omega) {
    cudaMemcpyAsync(z0, omega, (size_t)rows * D * 4, cudaMemcpyDeviceToDevice, st);
    TNF_TRY(tnf_base_logq(z0, log_q, rows, D, stream));
  } else {
    TNF_TRY(tnf_base_sample(z0, log_q, rows, D, seed, offset, stream));
  }
  if (small_chain_ok(chain, n_bij, D, Mp, freeze_bn))
    return launch_small(chain, n_bij, D, 0, z0, params, rows, z_out, nullptr, log_q, st);
  if (!freeze_bn && allreduce == nullptr && (peer == nullptr || peer->world <= 1) && small_chain_ok(chain, n_bij, D, Mp, 1)) {
    // live BatchNorm statistics, one rank: ONE cooperative launch when the batch fits a co-resident grid
    double* partial = (double*)ws.take((size_t)kLiveMaxBn * kLiveMaxGrid * 2 * (size_t)D * 8);
    unsigned int* counter = (unsigned int*)ws.take(256);
    if (partial && counter) {
      int rc = launch_small_live(chain, n_bij, D, z0, params, rows, z_out, log_q, partial, counter, st);
      if (rc != 1) return rc;
    }
  }

  cudaMemsetAsync(ld_acc, 0, (size_t)rows * 4, st);
  cudaMemsetAsync(scal, 0, (size_t)Mp * 4, st);
  const bool fold = all_tc(chain, n_bij) && Mp == 1;
  // BatchNorm statistics across ranks: inside the fold kernel over NVLink peer memory when the caller set it up and the
  // chain folds; otherwise through the all-reduce hook
  const bool use_peer = fold && peer != nullptr && peer->world > 1 && !freeze_bn;
  int n_exchanged = 0;
  const int64_t Mk = Mp == 1 ? 1 : M, Nk = Mp == 1 ? rows : N;
  const float* cur = z0;
  int nb = 1, np = 0;
  bool have_pend = false, have_stats = false;
  // the chain's last writer stores straight into z_out when nothing is pending after it
  auto out_buf = [&]() { float* o = zb[nb]; nb ^= 1; return o; };
  auto flush_pend = [&](float* dst) -> int {
    if (!have_pend) return 0;
    float* o = dst ? dst : out_buf();
    TNF_TRY(tnf_colaffine(cur, o, pend[np ^ 1][0], pend[np ^ 1][1], rows, D, stream));
    cur = o; have_pend = false;
    return 0;
  };
  for (int i = 0; i < n_bij; ++i) {
    const tnf_bijector_t& b = chain[i];
    const float* prm = params + b.param_offset;
    const bool last = i == n_bij - 1;
    if (b.kind == TNF_BIJ_REALNVP) {
      float* o = (last && D_out == D) ? z_out : out_buf();
      if (b.packed) {
        const bool want = !last && chain[i + 1].kind == TNF_BIJ_BATCHNORM && !freeze_bn &&
                          tc_stats_fusable(D, b.num_units, b.num_layers, tc_precision);
        if (b.ev_start) cudaEventRecord((cudaEvent_t)b.ev_start, st);
        // (the profiling stop event is recorded right after the coupling kernel, before the statistics reduce launch)
        TNF_TRY(coupling_tc_impl(cur, o, ld_acc, b.packed, rows, D, b.num_units, b.num_layers, b.transform_upper, TNF_FORWARD,
                                 TNF_LD_ADD, have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr,
                                 want ? sums : nullptr, want ? stat_ws : nullptr, tc_precision, 0, nullptr, nullptr, nullptr, stream,
                                 b.ev_stop));
        have_pend = false; have_stats = want;
      } else {
        TNF_TRY(flush_pend(nullptr));
        TNF_TRY(tnf_coupling(cur, o, ld_acc, prm, param_row_stride, Mk, Nk, D, b.num_units, b.num_layers, b.transform_upper,
                             TNF_FORWARD, TNF_LD_ADD, TNF_F32, stream));
        have_stats = false;
      }
      cur = o;
    } else if (b.kind == TNF_BIJ_BATCHNORM) {
      TNF_REQUIRE(b.bn_mean && b.bn_alpha && b.bn_log_det, TNF_ERR_ARG, "tnf_chain_sample: BatchNorm state missing");
      if (fold) {
        // one launch: statistics -> mean / alpha / log-det, (z - mean) / alpha folded into the pending map, the Affine
        // that follows folded too, both log-dets added - the arithmetic of the four launches of the generic branch
        if (!freeze_bn) {
          if (have_pend) { TNF_TRY(flush_pend(nullptr)); have_stats = false; }   // statistics are taken on the materialised tensor
          if (!have_stats) TNF_TRY(tnf_colstats(cur, rows, D, sums, stat_ws, TNF_F32, stream));
          if (allreduce && !use_peer) {
            int rc = allreduce(sums, 2 * D + 1, allreduce_user);
            TNF_REQUIRE(rc == 0, TNF_ERR_ARG, "tnf_chain_sample: the statistics all-reduce callback failed (%d)", rc);
          }
        }
        have_stats = false;
        const bool aff_next = !last && chain[i + 1].kind == TNF_BIJ_AFFINE;
        TNF_TRY(bn_fold_fwd_launch(sums, D, b.bn_eps, b.bn_mean, b.bn_alpha, b.bn_log_det, have_pend ? pend[np ^ 1][0] : nullptr,
                                   have_pend ? pend[np ^ 1][1] : nullptr, aff_next ? params + chain[i + 1].param_offset : nullptr,
                                   pend[np][0], pend[np][1], scal, freeze_bn ? 0 : 1, use_peer ? peer : nullptr,
                                   use_peer ? peer->seq + (unsigned long long)(n_exchanged++) : 0ull, st));
        np ^= 1; have_pend = true;
        if (aff_next) ++i;
        continue;
      }
      if (!freeze_bn) {
        if (have_pend) { TNF_TRY(flush_pend(nullptr)); have_stats = false; }   // statistics are taken on the materialised tensor
        if (!have_stats) TNF_TRY(tnf_colstats(cur, rows, D, sums, stat_ws, TNF_F32, stream));
        if (allreduce) {
          int rc = allreduce(sums, 2 * D + 1, allreduce_user);
          TNF_REQUIRE(rc == 0, TNF_ERR_ARG, "tnf_chain_sample: the statistics all-reduce callback failed (%d)", rc);
        }
        TNF_TRY(tnf_bn_finalize(sums, D, b.bn_eps, b.bn_mean, b.bn_alpha, b.bn_log_det, TNF_F32, stream));
      }
      have_stats = false;
      if (fold) {
        TNF_TRY(tnf_fold_colaffine(have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr, TNF_FOLD_BN_FWD,
                                   b.bn_mean, b.bn_alpha, pend[np][0], pend[np][1], nullptr, D, stream));
        np ^= 1; have_pend = true;
      } else {
        float* o = (last && D_out == D) ? z_out : out_buf();
        TNF_TRY(tnf_bn_apply(cur, o, b.bn_mean, b.bn_alpha, rows, D, TNF_FORWARD, TNF_F32, stream));
        cur = o;
      }
      TNF_TRY(tnf_accum_bcast(scal, b.bn_log_det, Mp, Mp, TNF_F32, stream));
    } else if (b.kind == TNF_BIJ_AFFINE) {
      if (fold) {
        TNF_TRY(tnf_fold_colaffine(have_pend ? pend[np ^ 1][0] : nullptr, have_pend ? pend[np ^ 1][1] : nullptr, TNF_FOLD_AFF_FWD,
                                   prm, prm + D, pend[np][0], pend[np][1], scal, D, stream));
        np ^= 1; have_pend = true;
      } else {
        float* o = out_buf();     // its log-det scratch lives in the buffer's tail, so never z_out
        float* ldm = o + (size_t)rows * D;
        TNF_TRY(tnf_affine(cur, o, ldm, prm, param_row_stride, Mk, Nk, D, TNF_FORWARD, TNF_F32, stream));
        TNF_TRY(tnf_accum_bcast(scal, ldm, Mp, 1, TNF_F32, stream));
        cur = o;
      }
      have_stats = false;
    } else if (b.kind == TNF_BIJ_TOINTERVAL) {
      TNF_TRY(flush_pend(nullptr));
      float* o = last ? z_out : out_buf();
      TNF_TRY(tnf_tointerval(cur, o, ld_acc, b.consts, rows, D, TNF_FORWARD, TNF_LD_ADD, TNF_F32, stream));
      cur = o; have_stats = false;
    } else {   // ToSimplex (last): (rows, D) -> (rows, D + 1)
      TNF_TRY(flush_pend(nullptr));
      TNF_TRY(tnf_tosimplex(cur, z_out, ld_acc, rows, D, b.num_units /* D attribute */, TNF_LD_ADD, TNF_F32, stream));
      cur = z_out;
    }
  }
  if (have_pend) TNF_TRY(flush_pend(z_out));
  if (cur != z_out) cudaMemcpyAsync(z_out, cur, (size_t)rows * D_out * 4, cudaMemcpyDeviceToDevice, st);
  return tnf_finish_logq(log_q, ld_acc, scal, Mp == M && M > 1 ? N : rows, rows, TNF_F32, stream);
}

}  // extern "C"
