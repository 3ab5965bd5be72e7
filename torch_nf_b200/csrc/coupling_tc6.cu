// RealNVP coupling layer on tcgen05 tensor cores at the REFERENCE'S precision (fp32 parity): TNF_TC_FP32.
//
// The reference computes the conditioner in fp32 (torch_nf/bijectors.py:237-241).  Here every GEMM operand is split
// into two fp16 parts, x = hi + lo with hi = fp16(x), lo = fp16(x - hi) (22 significand bits together), and every
// product is three tcgen05.mma.kind::f16 into the SAME fp32 TMEM accumulator:
//     A_hi.W_hi + A_lo.W_hi + A_hi.W_lo            (the dropped A_lo.W_lo term is ~2^-22 relative)
// so the conditioner costs 3x the bf16 tensor time instead of running on CUDA cores (~50x slower, measured).  tanh and
// exp are evaluated to fp32 accuracy (ex2.approx + rcp.approx: abs error ~2e-7), biases are added exactly by a bias
// MMA with a three-way fp16 split (hi, mid, lo = 33 bits).
//
// One 128-row tile in flight per CTA, CTA pairs (cta_group::2, M = 256) share every weight operand as in
// coupling_tc4/5.  All 16 epilogue warps work on the tile (TMEM lane quadrant w%4, accumulator chunks c = w/4 mod 4).
// Activations never touch shared memory: the 512 TMEM columns are two 256-column regions R0, R1.  Layer l's
// accumulator lives in R(l%2) and is issued as two N = U/2 halves; the tanh phase of a half writes the fp16 hi / lo
// activations back IN PLACE (chunk c: columns [32c, 32c+16) = hi pairs, [32c+16, 32c+32) = lo pairs - exactly the 32
// columns it has just read), while the tensor pipe computes the other half; layer l+1 reads them as its A operand
// from tensor memory and accumulates into the other region.  The first layer's A operand (the conditioning half of z)
// comes from hi / lo shared-memory images written by the two I/O warps (double buffered across tiles).
//
// The same kernel without the split (kSplit = false: bf16 operands, one MMA per product, tanh.approx) is the bf16
// path of the D = 256 layer (BASELINE.json configuration 5), whose activation images do not fit shared memory next to
// two tiles: here they need none.
//
// Job sequence per tile and net (t, then s): (0,a) (0,b) (1,a) (1,b) ... (L-1,a) (L-1,b) final.
// Reference semantics: torch_nf/bijectors.py:145-242 (RealNVP).
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace tnf {
namespace tc {

constexpr int kThreads6 = (kEpiWarps2 + 4) * 32;   // 16 epilogue, MMA, producer, 2 I/O
constexpr int kBiasPad6 = 4096;                    // zero block after the resident bias region
constexpr int kMaxStages6 = 10;                    // weight ring: as many 16 KB stages as fit (no activation images here)

struct Shape6 {
  int D, U, L, upper, split, DH, Q, NBF, c_off, t_off;
  // split != 0: fp16 hi + lo images (fp32 parity); 0: one bf16 image
  __host__ __device__ Shape6(int D_, int U_, int L_, int upper_, int split_ = 1) : D(D_), U(U_), L(L_), upper(upper_), split(split_) {
    DH = D / 2; Q = U / 4; NBF = DH / 2;
    c_off = upper ? 0 : DH;
    t_off = upper ? DH : 0;
  }
  __host__ __device__ int K_of(int l) const { return l == 0 ? DH : U; }
  __host__ __device__ int NB_of(int l) const { return l < L ? Q : NBF; }          // B rows per CTA of one job
  __host__ __device__ int ebytes() const { return split ? 4 : 2; }                  // bytes per weight element in a stage
  __host__ __device__ int KS_of(int l) const {                                     // K rows per 16 KB stage
    const int ks = (kStageBytes / ebytes()) / NB_of(l);
    return ks < K_of(l) ? ks : K_of(l);
  }
  __host__ __device__ int64_t job_bytes(int l) const { return (int64_t)K_of(l) * NB_of(l) * ebytes(); }   // one half
  __host__ __device__ int64_t net_bytes() const {
    int64_t b = 0;
    for (int l = 0; l < L; ++l) b += 2 * job_bytes(l);
    return b + job_bytes(L);
  }
  __host__ __device__ int64_t w_off(int rank, int net, int l, int h) const {
    int64_t off = (int64_t)(rank * 2 + net) * net_bytes();
    for (int i = 0; i < l; ++i) off += 2 * job_bytes(i);
    return off + (int64_t)h * job_bytes(l);
  }
  __host__ __device__ int64_t bias_rank_bytes() const { return 2 * 16 * ((int64_t)L * 2 * Q + NBF); }
  __host__ __device__ int64_t bias_off(int net, int l, int h) const {
    return (int64_t)net * (bias_rank_bytes() / 2) + (int64_t)l * 2 * Q * 16 + (int64_t)h * Q * 16;
  }
  __host__ __device__ int64_t bias_base(int rank) const { return 4 * net_bytes() + (int64_t)rank * bias_rank_bytes(); }
  __host__ __device__ int64_t packed_bytes() const { return 4 * net_bytes() + 2 * bias_rank_bytes(); }
  __host__ __device__ size_t a1_bytes() const { return (size_t)kTileM * DH * 2; }   // one fp16 image (hi or lo)
};

bool shape_supported6(int D, int U, int L) {
  return (D == 64 || D == 128 || D == 256) && (U == 128 || U == 256) && L >= 1 && L <= 5;
}
size_t packed_bytes6(int D, int U, int L, int split) { return (size_t)Shape6(D, U, L, 1, split).packed_bytes(); }

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {   // low half = first (lower K index) element
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// x -> (fp16(x), fp16(x - fp16(x))) for a pair of values, each packed like pack_f16
__device__ __forceinline__ void split_f16(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16(a, b);
  const float2 h = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16(a - h.x, b - h.y);
}
// tanh to fp32 accuracy: (e - 1) / (e + 1), e = 2^(2 log2(e) |x|); |x| clamped at 10 (tanh = 1 in fp32 beyond 9.02)
__device__ __forceinline__ float tanh_f32(float x) {
  const float ax = fminf(fabsf(x), 10.0f);
  const float e = exp2_fast(ax * 2.8853900817779268f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return copysignf((e - 1.0f) * r, x);
}

// ---------------------------------------------------------------- weight packing (fp32 row -> fp16 hi / lo images)
__global__ void pack6_kernel(const float* __restrict__ params, unsigned char* __restrict__ packed, Shape6 sh) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t total = 0;
  for (int l = 0; l <= sh.L; ++l) total += 2 * (int64_t)sh.K_of(l) * (l < sh.L ? sh.U : sh.DH);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t rem = idx, src_off = 0;
    int l = 0;
    for (;; ++l) {
      const int64_t n_el = (int64_t)sh.K_of(l) * (l < sh.L ? sh.U : sh.DH);
      if (rem < 2 * n_el) break;
      rem -= 2 * n_el;
      src_off += 2 * n_el + 2 * (l < sh.L ? sh.U : sh.DH);
    }
    const int K = sh.K_of(l), J = l < sh.L ? sh.U : sh.DH;
    const int64_t n_el = (int64_t)K * J;
    const int net = rem >= n_el;
    rem -= net * n_el;
    const int k = (int)(rem / J), j = (int)(rem % J);
    const float w = params[src_off + net * n_el + rem];   // W_t then W_s, (K, J) row-major, x @ W (bijectors.py:224-235)
    const int NB = sh.NB_of(l), KS = sh.KS_of(l);
    const int h = l < sh.L ? j / (2 * sh.Q) : 0;
    const int rank = l < sh.L ? (j / sh.Q) & 1 : j / sh.NBF;
    const int n = j % NB;
    unsigned char* stage = packed + sh.w_off(rank, net, l, h) + (int64_t)(k / KS) * KS * NB * sh.ebytes();
    if (sh.split) {
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      *reinterpret_cast<__half*>(stage + img_off(n, k % KS, NB)) = hi;
      *reinterpret_cast<__half*>(stage + (int64_t)KS * NB * 2 + img_off(n, k % KS, NB)) = lo;
    } else {
      *reinterpret_cast<__nv_bfloat16*>(stage + img_off(n, k % KS, NB)) = __float2bfloat16_rn(w);
    }
  }
  int nb = sh.L * sh.U + sh.DH;   // biases per net
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * nb; idx += stride) {
    const int net = idx >= nb;
    int rem = (int)(idx - net * nb), l = 0;
    int64_t src_off = 0;
    for (;; ++l) {
      const int J = l < sh.L ? sh.U : sh.DH;
      if (rem < J) break;
      rem -= J;
      src_off += 2 * (int64_t)sh.K_of(l) * J + 2 * J;
    }
    const int K = sh.K_of(l), J = l < sh.L ? sh.U : sh.DH;
    const float b = params[src_off + 2 * (int64_t)K * J + (net ? J : 0) + rem];
    const int h = l < sh.L ? rem / (2 * sh.Q) : 0;
    const int rank = l < sh.L ? (rem / sh.Q) & 1 : rem / sh.NBF;
    const int n = rem % sh.NB_of(l);
    unsigned char* blkp = packed + sh.bias_base(rank) + sh.bias_off(net, l, h) + (int64_t)n * 16;
    if (sh.split) {
      __half* blk = reinterpret_cast<__half*>(blkp);
      const __half b0 = __float2half_rn(b);
      const float r1 = b - __half2float(b0);
      const __half b1 = __float2half_rn(r1);
      const __half b2 = __float2half_rn(r1 - __half2float(b1));
      blk[0] = b0; blk[1] = b1; blk[2] = b2;
      for (int kk = 3; kk < 8; ++kk) blk[kk] = __float2half_rn(0.f);
    } else {
      __nv_bfloat16* blk = reinterpret_cast<__nv_bfloat16*>(blkp);
      const __nv_bfloat16 b0 = __float2bfloat16_rn(b);
      blk[0] = b0; blk[1] = __float2bfloat16_rn(b - __bfloat162float(b0)); blk[2] = __float2bfloat16_rn(0.f);
      for (int kk = 3; kk < 8; ++kk) blk[kk] = __float2bfloat16_rn(0.f);
    }
  }
}

int pack6_launch(const float* params, void* packed, int D, int U, int L, int upper, int split, cudaStream_t st) {
  Shape6 sh(D, U, L, upper, split);
  pack6_kernel<<<num_sms() * 4, 256, 0, st>>>(params, (unsigned char*)packed, sh);
  return 0;
}

// ---------------------------------------------------------------- kernel
struct __align__(16) Ctrl6 {
  uint64_t w_full[kMaxStages6];
  uint64_t w_empty[kMaxStages6];
  uint64_t w_peer[kMaxStages6];   // leader only: the second CTA's half of the stage has landed
  uint64_t a1_ready[2];   // per A1 buffer, I/O warps of both CTAs (leader's barrier): images of a tile written
  uint64_t a1_free[2];    // per A1 buffer, tcgen05.commit: the tile's layer-0 jobs (both nets) have read the images
  uint64_t t_done;        // 16 epilogue warps of both CTAs (leader's barrier): a tanh phase is complete
  uint64_t e_fin;         // 16 epilogue warps of both CTAs (leader's barrier): the final accumulator has been read
  uint64_t h_ready[2];    // tcgen05.commit: N-half a / b of the current hidden job complete
  uint64_t f_ready;       // tcgen05.commit: final-layer accumulator complete
  uint64_t y_done;        // this CTA's 16 epilogue warps: the tile's transformed half is stored (fused statistics)
  uint32_t tmem_base;
  uint32_t pad;
};
// dynamic shared memory:
//   [ring: S x 16 KB][A1 hi/lo x 2 buffers][ones 4 KB][Ctrl6][pre_scale D][pre_shift D][ld partial 4 x 128]
//   [resident bias blocks of this rank][4 KB zeros]
size_t smem_bytes6(int D, int U, int L, int split, int n_stages) {
  Shape6 sh(D, U, L, 1, split);
  return (size_t)n_stages * kStageBytes + (split ? 4 : 2) * sh.a1_bytes() + kOnesBytes + sizeof(Ctrl6) +
         (size_t)(2 * sh.D + 4 * kTileM + 4 * kTileM) * sizeof(float) + (size_t)(16 * 64) * sizeof(double) +
         (size_t)sh.bias_rank_bytes() + kBiasPad6;
}

// dense, D = f32, A = B = f16 (split) or bf16, K-major, M = 256 (pair)
__host__ __device__ inline uint32_t make_idesc6(int N, bool split) { return split ? ((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24)) : make_idesc2(N); }

// One job of the MMA warp: bias MMA (ones image x resident bias block), then for all K the two CORRECTION products
// (A_lo, W_hi) (A_hi, W_lo), then for all K the main product (A_hi, W_hi).  Order matters: the tensor core rounds the
// fp32 accumulator once per MMA instruction (towards zero, as measured: error grows with the number of instructions
// executed while the accumulator is large); with the small products first only the K/16 main instructions see a
// full-size accumulator (3x fewer).  All stages of a job (<= 4 of the ring's >= 8) are therefore waited for up front
// and released after their main MMAs.
// kTS: A operand from tensor memory (in-place activations: chunk c = kk/2 holds the hi pairs of its units at columns
// 32c + 8*(kk%2) .. +8 and the lo pairs 16 columns further); else from the hi / lo shared-memory images (descriptor low
// words a0 / a1, +256 per K step).
template <int K, int N, bool kTS, bool kSplit>
__device__ __forceinline__ void mma_job6(uint32_t d_tmem, uint32_t a0, uint32_t a1, uint32_t a_dhi, uint32_t ring16,
                                         uint32_t b_hi, uint32_t ones_lo, uint32_t bias_lo, uint32_t wfull0,
                                         uint32_t wpeer0, uint32_t wempty0, uint32_t S, uint32_t& slot, uint32_t& phase,
                                         bool leader, long long* t_w) {
  constexpr int NB = N / 2;
  constexpr int KSMAX = (kStageBytes / (kSplit ? 4 : 2)) / NB;
  constexpr int KS = KSMAX < K ? KSMAX : K;
  constexpr int SPS = KS / 16;                                   // K steps per stage
  constexpr int NST = K / KS;                                    // stages of this job
  constexpr uint32_t kStage16 = kStageBytes >> 4;
  constexpr uint32_t kLo16 = (uint32_t)(KS * NB * 2) >> 4;       // W_lo image inside the stage
  const uint32_t idesc = make_idesc6(N, kSplit);
  uint32_t sl[NST];
  {
    const long long c0 = t_w ? clock64() : 0;
#pragma unroll
    for (int st = 0; st < NST; ++st) {
      uint32_t s_ = slot + (uint32_t)st, ph = phase;
      if (s_ >= S) { s_ -= S; ph ^= 1u; }
      sl[st] = s_;
      mbar_wait_addr(wfull0 + s_ * 8u, ph);
      mbar_wait_addr(wpeer0 + s_ * 8u, ph);
    }
    if (t_w) *t_w += clock64() - c0;
    tc_fence_after();
  }
  if (leader) {
    umma2_ss2(d_tmem, ones_lo, a_dhi, bias_lo, b_hi, idesc, 0u);
#pragma unroll
    for (int kk = 0; kk < (kSplit ? K / 16 : 0); ++kk) {
      const uint32_t bw = ring16 + sl[kk / SPS] * kStage16 + (uint32_t)((kk % SPS) * 2 * NB);   // 2 K groups x NB rows x 16 B
      if (kTS) {
        const uint32_t col = (uint32_t)(32 * (kk / 2) + 8 * (kk % 2));
        umma2_ts2(d_tmem, a0 + col + 16u, bw, b_hi, idesc, 1u);
        umma2_ts2(d_tmem, a0 + col, bw + kLo16, b_hi, idesc, 1u);
      } else {
        umma2_ss2(d_tmem, a1 + 256u * kk, a_dhi, bw, b_hi, idesc, 1u);
        umma2_ss2(d_tmem, a0 + 256u * kk, a_dhi, bw + kLo16, b_hi, idesc, 1u);
      }
    }
#pragma unroll
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint32_t bw = ring16 + sl[kk / SPS] * kStage16 + (uint32_t)((kk % SPS) * 2 * NB);
      if (kTS) umma2_ts2(d_tmem, a0 + (uint32_t)(32 * (kk / 2) + 8 * (kk % 2)), bw, b_hi, idesc, 1u);
      else umma2_ss2(d_tmem, a0 + 256u * kk, a_dhi, bw, b_hi, idesc, 1u);
      if (kk % SPS == SPS - 1) tc_commit2_addr(wempty0 + sl[kk / SPS] * 8u);
    }
  }
  slot += NST;
  if (slot >= S) { slot -= S; phase ^= 1u; }
}

template <bool kInverse, int DH, int U_, bool kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads6, 1) coupling_tc6_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Shape6 sh(a.D, a.U, a.L, a.upper, kSplit ? 1 : 0);
  constexpr int kImgs = kSplit ? 2 : 1;            // A1 images per buffer (hi, lo | bf16)
  const int L = sh.L;
  const uint32_t S = (uint32_t)a.n_stages;
  unsigned char* ring = smem_raw;
  unsigned char* sA1 = ring + (size_t)S * kStageBytes;            // [buffer][hi | lo]
  unsigned char* sOnes = sA1 + 2 * kImgs * sh.a1_bytes();
  Ctrl6& ct = *reinterpret_cast<Ctrl6*>(sOnes + kOnesBytes);
  float* s_pscale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(Ctrl6));
  float* s_pshift = s_pscale + sh.D;
  float* s_ldp = s_pshift + sh.D;                                  // [3][128] log-det partials of chunk owners 1..3
  // [tile & 3][128] sum z^2 of the conditioning half (fused base density).  Three tiles are live at once: the I/O warps
  // load tile k+2 as soon as tile k's layer-0 jobs are done, before tile k's epilogue has consumed its entry
  float* s_ss = s_ldp + 4 * kTileM;
  double* s_acc = reinterpret_cast<double*>(s_ss + 4 * kTileM);    // [16][64]: the I/O threads' float64 accumulators (fused statistics)
  // [2 I/O warps][2 * D] column sums, written once after the CTA's last tile: in the A1 images, dead by then (an I/O
  // warp gets there through y_done of the last tile, which follows every MMA that reads them, in both CTAs of the pair)
  double* s_stat = reinterpret_cast<double*>(sA1);
  unsigned char* sBias = reinterpret_cast<unsigned char*>(s_acc + 16 * 64);
  const bool lp_mode = kInverse && a.out_lp != nullptr;           // this is the chain's last executed layer: emit log_prob, not z
  // Column statistics of the OUTPUT (the next BatchNorm's batch statistics), taken by the two I/O warps in float64: the
  // conditioning half as it passes through, the transformed half read back (L2) once the epilogue warps stored a tile
  const bool want_stats = !kInverse && a.stat_partials != nullptr;   // sample direction only

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const uint32_t rank = cluster_ctarank();
  const int64_t P = gridDim.x / 2, pair = blockIdx.x / 2;
  const int64_t n_super = (n_tiles + 1) / 2;
  const int64_t cnt = pair < n_super ? (n_super - pair + P - 1) / P : 0;   // tiles of this CTA: 2*(k*P + pair) + rank
  constexpr int n_chunks = U_ / kChunk;
  constexpr int Q = U_ / 4;

  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < S; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); mbar_init(&ct.w_peer[i], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&ct.a1_ready[b], 4); mbar_init(&ct.a1_free[b], 1); mbar_init(&ct.h_ready[b], 1); }
    mbar_init(&ct.t_done, 2 * kEpiWarps2);
    mbar_init(&ct.e_fin, 2 * kEpiWarps2);
    mbar_init(&ct.f_ready, 1);
    mbar_init(&ct.y_done, kEpiWarps2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps2) tmem_alloc2(&ct.tmem_base, 512);
  {
    // ones image: row r, K columns 0, 1, 2 = 1.0 (fp16 0x3c00), everything else 0
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x) {
      uint32_t v = 0u;
      if (i < kTileM * 4) {
        if (kSplit) v = (i & 3) == 0 ? 0x3c003c00u : ((i & 3) == 1 ? 0x00003c00u : 0u);   // fp16 1.0 in K columns 0, 1, 2
        else v = (i & 3) == 0 ? 0x3f803f80u : 0u;                                          // bf16 1.0 in K columns 0, 1
      }
      reinterpret_cast<uint32_t*>(sOnes)[i] = v;
    }
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
    const int nb16 = (int)(sh.bias_rank_bytes() / 16);
    const uint4* gb = reinterpret_cast<const uint4*>(a.packed + sh.bias_base((int)rank));
    for (int i = threadIdx.x; i < nb16 + kBiasPad6 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sBias)[i] = i < nb16 ? __ldg(gb + i) : make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;
  const uint32_t lead_t_done = mapa_u32(smem_u32(&ct.t_done), 0u), lead_e_fin = mapa_u32(smem_u32(&ct.e_fin), 0u);
  const uint32_t lead_a1_ready = mapa_u32(smem_u32(&ct.a1_ready[0]), 0u), lead_w_peer = mapa_u32(smem_u32(&ct.w_peer[0]), 0u);

  if (warp == kEpiWarps2 + 1) {
    // =============================== weight producer (one elected lane) ===============================
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int64_t k = 0; k < cnt; ++k) {
        for (int net = 0; net < 2; ++net) {
          for (int l = 0; l <= L; ++l) {
            const int KS = sh.KS_of(l), NB = sh.NB_of(l);
            const uint32_t bytes = (uint32_t)(KS * NB * sh.ebytes());
            for (int h = 0; h < (l < L ? 2 : 1); ++h) {
              const unsigned char* src = a.packed + sh.w_off((int)rank, net, l, h);
              for (int st = 0; st < sh.K_of(l) / KS; ++st) {
                mbar_wait(&ct.w_empty[slot], phase ^ 1u);
                mbar_arrive_expect_tx(&ct.w_full[slot], bytes);
                bulk_g2s(ring + (size_t)slot * kStageBytes, src + (size_t)st * bytes, bytes, &ct.w_full[slot]);
                if (++slot == S) { slot = 0; phase ^= 1u; }
              }
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps2 && rank != 0) {
    // =============================== second CTA: forward "stage landed" to the leader ===============================
    uint32_t slot = 0, phase = 0;
    int64_t n_st = 0;
    for (int l = 0; l <= L; ++l) n_st += (l < L ? 2 : 1) * (sh.K_of(l) / sh.KS_of(l));
    n_st *= 2 * cnt;
    for (int64_t i = 0; i < n_st; ++i) {
      mbar_wait(&ct.w_full[slot], phase);
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_w_peer + slot * 8u);
      if (++slot == S) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == kEpiWarps2) {
    // =============================== leader CTA: MMA issuer ===============================
    const bool leader = elect_one();
    uint32_t slot = 0, phase = 0, t_par = 0, f_par = 0, a1_par = 0;
    bool fin_pending = false;
    const bool diag = a.dbg != nullptr;
    const long long t_all = diag ? clock64() : 0;
    long long t_dep = 0, t_wt = 0;
    const uint32_t wfull0 = smem_u32(&ct.w_full[0]), wpeer0 = smem_u32(&ct.w_peer[0]), wempty0 = smem_u32(&ct.w_empty[0]);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint64_t a1_desc = make_desc(smem_u32(sA1), kTileM);
    const uint32_t a_dhi = (uint32_t)(a1_desc >> 32);
    const uint32_t a1_sz16 = (uint32_t)sh.a1_bytes() >> 4;
    const uint64_t bH_desc = make_desc(0u, Q), bF_desc = make_desc(0u, DH / 2);
    const uint32_t bH_hi = (uint32_t)(bH_desc >> 32), bF_hi = (uint32_t)(bF_desc >> 32);
    const uint32_t lboH = (uint32_t)bH_desc, lboF = (uint32_t)bF_desc;
    const uint32_t ones_lo = (uint32_t)make_desc(smem_u32(sOnes), kTileM);
    const uint32_t bias16 = smem_u32(sBias) >> 4;
    long long* tw = diag ? &t_wt : nullptr;
    for (int64_t k = 0; k < cnt; ++k) {
      const uint32_t buf = (uint32_t)(k & 1);
      for (int net = 0; net < 2; ++net) {
        for (int l = 0; l <= L; ++l) {
          const long long c0 = diag ? clock64() : 0;
          if (l == 0) {
            if (fin_pending) {   // the previous final accumulator (same region as this layer's) has been read out
              mbar_wait_addr(smem_u32(&ct.e_fin), f_par);
              f_par ^= 1u;
              fin_pending = false;
            }
            if (net == 0) {
              mbar_wait_addr(smem_u32(&ct.a1_ready[buf]), (a1_par >> buf) & 1u);
              a1_par ^= 1u << buf;
            }
          } else {               // the previous tanh phase: all K of this layer's A operand written
            mbar_wait_addr(smem_u32(&ct.t_done), t_par);
            t_par ^= 1u;
          }
          tc_fence_after();
          if (diag) t_dep += clock64() - c0;
          const uint32_t rD = tmem + (uint32_t)((l & 1) * 256), rA = tmem + (uint32_t)(((l + 1) & 1) * 256);
          const uint32_t a1h = (uint32_t)a1_desc + ((uint32_t)kImgs * buf) * a1_sz16, a1l = a1h + a1_sz16;
          if (l < L) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t bias_lo = lboH + bias16 + (uint32_t)(sh.bias_off(net, l, h) >> 4);
              if (l == 0)
                mma_job6<DH, U_ / 2, false, kSplit>(rD + (uint32_t)(h * (U_ / 2)), a1h, a1l, a_dhi, lboH + ring16, bH_hi, ones_lo, bias_lo,
                                            wfull0, wpeer0, wempty0, S, slot, phase, leader, tw);
              else
                mma_job6<U_, U_ / 2, true, kSplit>(rD + (uint32_t)(h * (U_ / 2)), rA, 0u, a_dhi, lboH + ring16, bH_hi, ones_lo, bias_lo,
                                           wfull0, wpeer0, wempty0, S, slot, phase, leader, tw);
              if (leader) tc_commit2_addr(smem_u32(&ct.h_ready[h]));
            }
            if (l == 0 && net == 1 && leader) tc_commit2_addr(smem_u32(&ct.a1_free[buf]));
          } else {
            const uint32_t bias_lo = lboF + bias16 + (uint32_t)(sh.bias_off(net, L, 0) >> 4);
            mma_job6<U_, DH, true, kSplit>(rD, rA, 0u, a_dhi, lboF + ring16, bF_hi, ones_lo, bias_lo, wfull0, wpeer0, wempty0, S, slot,
                                   phase, leader, tw);
            if (leader) tc_commit2_addr(smem_u32(&ct.f_ready));
            fin_pending = true;
          }
          __syncwarp();
        }
      }
    }
    if (diag && blockIdx.x == 0 && leader) { a.dbg[2040] = t_dep; a.dbg[2041] = t_wt; a.dbg[2042] = clock64() - t_all; }
  } else if (warp < kEpiWarps2) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, cq = warp >> 2;           // TMEM lane quadrant, chunk owner 0..3
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r_tile = q * 32 + lane;
    constexpr int W = DH / 4;                         // final-layer columns per thread
    const float kLog2e = 1.4426950408889634f;
    uint32_t h_par = 0, f_par = 0;                    // h_par: bit h
    const float lp_cst = (float)((double)sh.D * 0.91893853320467274178);   // D log sqrt(2 pi)
    const float lp_scal0 = (lp_mode && a.lp_scal) ? a.lp_scal[0] : 0.f;
    // tanh phase of one accumulator chunk, in place: 32 fp32 columns -> 16 columns of hi pairs + 16 of lo pairs
    auto tanh_chunk = [&](uint32_t col) {
      uint32_t x[32];
      tmem_ld32(col, x);
      tc_wait_ld();
      if (kSplit) {
        uint32_t o[32];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float y0 = tanh_f32(__uint_as_float(x[j])), y1 = tanh_f32(__uint_as_float(x[j + 1]));
          split_f16(y0, y1, o[j >> 1], o[16 + (j >> 1)]);
        }
        tmem_st32(col, o);
      } else {   // bf16 pairs into the first 16 of the chunk's 32 columns
        uint32_t o[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2)
          o[j >> 1] = pack_bf16(tanh_fast(__uint_as_float(x[j])), tanh_fast(__uint_as_float(x[j + 1])));
        tmem_st16(col, o);
      }
    };
    for (int64_t k = 0; k < cnt; ++k) {
      const int64_t tile = 2 * (k * P + pair) + rank;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = row < a.rows;
      const float* zrow = a.z_in + row * sh.D + sh.t_off + cq * W;
      float tv[W];
#pragma unroll
      for (int net = 0; net < 2; ++net) {
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          const uint32_t reg = tmem + lane_addr + (uint32_t)((l & 1) * 256);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mbar_wait(&ct.h_ready[h], (h_par >> h) & 1u);
            h_par ^= 1u << h;
            tc_fence_after();
            // chunks of this half owned by this warp: c = cq, cq + 4, ... inside [h * n_chunks/2, (h+1) * n_chunks/2)
#pragma unroll
            for (int c = cq; c < n_chunks; c += 4)
              if (c / (n_chunks / 2) == h) tanh_chunk(reg + (uint32_t)(c * kChunk));
          }
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(lead_t_done);
        }
        // ---- final layer of this net: W columns per thread
        float zin[W];
        float ld_old = 0.f;
        if (net == 1) {
#pragma unroll
          for (int j = 0; j < W; j += 8) {   // 256-bit accesses: half the instructions, each 32 L1 wavefronts (one row per lane)
            if (valid) {
              ldg256_nc(zrow + j, zin + j);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) zin[j + e] = 0.f;
            }
          }
          if (cq == 0 && valid && a.accum != TNF_LD_WRITE) ld_old = a.log_det[row];
        }
        mbar_wait(&ct.f_ready, f_par);
        f_par ^= 1u;
        tc_fence_after();
        constexpr int PW = W < 16 ? W : 16;          // the final accumulator is read in pieces of <= 16 columns
        const uint32_t fcol = tmem + lane_addr + (uint32_t)((L & 1) * 256) + (uint32_t)(cq * W);
        float ld_sum = 0.f;
        float (&y)[W] = zin;
#pragma unroll
        for (int p0 = 0; p0 < W; p0 += PW) {
          uint32_t o[PW];
          if (PW == 8) tmem_ld8(fcol + (uint32_t)p0, reinterpret_cast<uint32_t(&)[8]>(o));
          else tmem_ld16(fcol + (uint32_t)p0, reinterpret_cast<uint32_t(&)[16]>(o));
          tc_wait_ld();
          if (p0 + PW == W) {     // accumulator read: the next layer-0 job may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(lead_e_fin);
          }
          if (net == 0) {
#pragma unroll
            for (int j = 0; j < PW; ++j) tv[p0 + j] = __uint_as_float(o[j]);
          } else {
#pragma unroll
            for (int j = 0; j < PW; ++j) {
              const int col = sh.t_off + cq * W + p0 + j;
              const float zz = fmaf(zin[p0 + j], s_pscale[col], s_pshift[col]);
              const float sv = __uint_as_float(o[j]);
              ld_sum += sv;
              // exp to fp32 accuracy (ex2.approx: 2 ulp); the inverse multiplies by exp(-s) = 1 / exp(s)
              y[p0 + j] = kInverse ? (zz - tv[p0 + j]) * exp2_fast(-sv * kLog2e) : fmaf(zz, exp2_fast(sv * kLog2e), tv[p0 + j]);
            }
          }
        }
        if (net == 1 && lp_mode) {   // log N(z_out) - log-dets instead of z_out: nothing reads the base sample itself
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < W; ++j) ss = fmaf(y[j], y[j], ss);
          if (cq != 0) s_ldp[(cq - 1) * kTileM + r_tile] = fmaf(0.5f, ss, ld_sum);
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
          if (cq == 0 && valid) {
            const float rest = (s_ldp[r_tile] + s_ldp[kTileM + r_tile]) + s_ldp[2 * kTileM + r_tile];
            const float ss_all = ss + s_ss[(int)(k & 3) * kTileM + r_tile];
            a.out_lp[row] = ((-0.5f * ss_all - lp_cst) - ((ld_old + ld_sum) + rest)) - lp_scal0;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
        } else if (net == 1) {
          if (valid) {
            float* orow = a.z_out + row * sh.D + sh.t_off + cq * W;
#pragma unroll
            for (int j = 0; j < W; j += 8) stg256(orow + j, y + j);
          }
          if (want_stats) {   // release the stored tile half to the I/O warps
            __syncwarp();
            if (lane == 0) mbar_arrive(&ct.y_done);
          }
          if (cq != 0) s_ldp[(cq - 1) * kTileM + r_tile] = ld_sum;
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // the four chunk owners of this lane quadrant
          if (cq == 0 && valid) {
            const float tot = ((ld_sum + s_ldp[r_tile]) + s_ldp[kTileM + r_tile]) + s_ldp[2 * kTileM + r_tile];
            float* op = a.log_det + row;
            if (a.accum == TNF_LD_WRITE) *op = tot;
            else if (a.accum == TNF_LD_ADD) *op = ld_old + tot;
            else *op = ld_old - tot;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // partials consumed before the next tile writes them
        }
      }
    }
  } else {
    // =============================== I/O warps: conditioning half, coalesced ===============================
    const int w2 = warp - (kEpiWarps2 + 2);
    const int row0 = w2 * (kTileM / 2);
    constexpr int LPR = DH / 4;
    constexpr int RPI = 32 / LPR;
    constexpr int NI = (kTileM / 2) / RPI;
    constexpr int kIoBatch = 8;                    // float4 loads in flight per lane (16 spills at 96 registers: measured 1.8x slower)
    const int hc = 4 * (lane % LPR);
    const int col = sh.c_off + hc;
    const int rsub = lane / LPR;
    const float4 ps = *reinterpret_cast<const float4*>(s_pscale + col);
    const float4 pb = *reinterpret_cast<const float4*>(s_pshift + col);
    // float64 accumulators of this thread's 4 + 4 columns (sum, sum of squares; conditioning half 0..7, transformed
    // half 8..15) live in shared memory: 16 doubles in registers for the whole kernel would spill the I/O loop.  Per
    // tile the (<= 16-row) partial sums are formed in fp32 and added once.
    double* my_acc = s_acc + (w2 * 32 + lane);   // element i at my_acc[64 * i]
    if (want_stats)
#pragma unroll
      for (int i = 0; i < 16; ++i) my_acc[64 * i] = 0.0;
    auto stats_tile = [&](int64_t k) {   // column sums of the transformed half of tile k, read back after its epilogue
      mbar_wait(&ct.y_done, (uint32_t)(k & 1));
      const int64_t tile = 2 * (k * P + pair) + rank;
      const int tcol = sh.t_off + hc;
      float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int n0 = 0; n0 < NI; n0 += kIoBatch) {
        float4 v[kIoBatch];
#pragma unroll
        for (int u = 0; u < kIoBatch; ++u) {
          const int64_t grow = tile * kTileM + row0 + (n0 + u) * RPI + rsub;
          v[u] = grow < a.rows ? __ldcg(reinterpret_cast<const float4*>(a.z_out + grow * sh.D + tcol))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kIoBatch; ++u) {
          p1[0] += v[u].x; p1[1] += v[u].y; p1[2] += v[u].z; p1[3] += v[u].w;
          p2[0] = fmaf(v[u].x, v[u].x, p2[0]); p2[1] = fmaf(v[u].y, v[u].y, p2[1]);
          p2[2] = fmaf(v[u].z, v[u].z, p2[2]); p2[3] = fmaf(v[u].w, v[u].w, p2[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) { my_acc[64 * (8 + e)] += (double)p1[e]; my_acc[64 * (12 + e)] += (double)p2[e]; }
    };
    for (int64_t k = 0; k < cnt; ++k) {
      const uint32_t buf = (uint32_t)(k & 1);
      if (k >= 2) mbar_wait(&ct.a1_free[buf], (uint32_t)(((k >> 1) - 1) & 1));   // tile k-2's layer-0 jobs are done with it
      const int64_t tile = 2 * (k * P + pair) + rank;
      unsigned char* a1h = sA1 + (size_t)(kImgs * buf) * sh.a1_bytes();
      unsigned char* a1l = a1h + sh.a1_bytes();
      float c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};   // this tile's partial sums, conditioning half
#pragma unroll 1
      for (int n0 = 0; n0 < NI; n0 += kIoBatch) {
        float4 v[kIoBatch];
#pragma unroll
        for (int u = 0; u < kIoBatch; ++u) {
          const int r = row0 + (n0 + u) * RPI + rsub;
          const int64_t grow = tile * kTileM + r;
          v[u] = grow < a.rows ? __ldg(reinterpret_cast<const float4*>(a.z_in + grow * sh.D + col))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kIoBatch; ++u) {
          const int r = row0 + (n0 + u) * RPI + rsub;
          const int64_t grow = tile * kTileM + r;
          float4 x;
          x.x = fmaf(v[u].x, ps.x, pb.x); x.y = fmaf(v[u].y, ps.y, pb.y);
          x.z = fmaf(v[u].z, ps.z, pb.z); x.w = fmaf(v[u].w, ps.w, pb.w);
          if (kSplit) {
            uint32_t h0, l0, h1, l1;
            split_f16(x.x, x.y, h0, l0);
            split_f16(x.z, x.w, h1, l1);
            *reinterpret_cast<uint2*>(a1h + img_off(r, hc, kTileM)) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(a1l + img_off(r, hc, kTileM)) = make_uint2(l0, l1);
          } else {
            *reinterpret_cast<uint2*>(a1h + img_off(r, hc, kTileM)) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
          }
          if (lp_mode) {   // sum of squares of this row's conditioning half (it passes through unchanged), for the epilogue
            float ssq = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
#pragma unroll
            for (int o = 1; o < LPR; o <<= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
            if (lane % LPR == 0) s_ss[(int)(k & 3) * kTileM + r] = ssq;
          } else if (grow < a.rows) {
            *reinterpret_cast<float4*>(a.z_out + grow * sh.D + col) = x;
            if (want_stats) {
              c1[0] += x.x; c1[1] += x.y; c1[2] += x.z; c1[3] += x.w;
              c2[0] = fmaf(x.x, x.x, c2[0]); c2[1] = fmaf(x.y, x.y, c2[1]);
              c2[2] = fmaf(x.z, x.z, c2[2]); c2[3] = fmaf(x.w, x.w, c2[3]);
            }
          }
        }
      }
      fence_async_smem();
      if (lp_mode) __threadfence_block();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_a1_ready + buf * 8u);
      if (want_stats) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { my_acc[64 * e] += (double)c1[e]; my_acc[64 * (4 + e)] += (double)c2[e]; }
        if (k >= 1) stats_tile(k - 1);   // a tile behind: its epilogue ends about when this load does
      }
    }
    if (want_stats) {
      if (cnt > 0) stats_tile(cnt - 1);
      double* rowp = s_stat + (size_t)w2 * 2 * sh.D;   // this warp's [2 * D] row: the two halves cover all D columns
#pragma unroll 1
      for (int i = 0; i < 16; ++i) {                   // lanes with the same lane % LPR own the same columns
        double v = my_acc[64 * i];
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int e = i & 3, half_off = i < 8 ? sh.c_off : sh.t_off, sq = (i >> 2) & 1;
        if (lane < LPR) rowp[sq * sh.D + half_off + 4 * lane + e] = v;
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (want_stats) {   // one [2 * D] row of doubles per CTA: the sum of the two I/O warps' rows
    for (int i = threadIdx.x; i < 2 * sh.D; i += blockDim.x)
      a.stat_partials[(size_t)blockIdx.x * 2 * sh.D + i] = s_stat[i] + s_stat[2 * sh.D + i];
  }
  cluster_sync_all();
  if (warp == kEpiWarps2) tmem_dealloc2(tmem, 512);
}

int launch_tc6(const Args& a, int grid, int split, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
#define TNF_TC6_LAUNCH(INV, DHV, UV, SP)                                                                          \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(coupling_tc6_kernel<INV, DHV, UV, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                                          \
    if (e == cudaSuccess) coupling_tc6_kernel<INV, DHV, UV, SP><<<grid, kThreads6, smem, st>>>(a);                \
  } while (0)
#define TNF_TC6_U(INV, DHV, SP)                                  \
  do {                                                           \
    if (a.U == 256) TNF_TC6_LAUNCH(INV, DHV, 256, SP);           \
    else TNF_TC6_LAUNCH(INV, DHV, 128, SP);                      \
  } while (0)
  if (split) {
    if (a.D == 64) { if (a.inverse) TNF_TC6_U(true, 32, true); else TNF_TC6_U(false, 32, true); }
    else if (a.D == 128) { if (a.inverse) TNF_TC6_U(true, 64, true); else TNF_TC6_U(false, 64, true); }
    else { if (a.inverse) TNF_TC6_U(true, 128, true); else TNF_TC6_U(false, 128, true); }
  } else {   // bf16: only the D = 256 layer runs here (D <= 128 has the two-tile kernels)
    if (a.inverse) TNF_TC6_U(true, 128, false); else TNF_TC6_U(false, 128, false);
  }
#undef TNF_TC6_U
#undef TNF_TC6_LAUNCH
  return (int)e;
}

}  // namespace tc
}  // namespace tnf
