// Hyper-network fusion on tensor cores: the same fused conditional log-density as cde_fused.cu, with the per-sample
// parameter rows  params[m, :] = h[m, :] . W_last + b_last  produced by tcgen05 MMAs into TENSOR MEMORY instead of
// fp32 FMAs in the consumer threads (conditional_density_estimator.py:34-37,101-104; SURVEY 8f #2: "turns the
// per-sample GEMV into a tensor-core GEMM ... and removes params from HBM entirely").
//
// A 128-sample tile is one M = 128 MMA tile and TMEM lane = sample: the accumulator row of a sample IS its parameter
// row, and `tcgen05.ld.32x32b` hands every consumer thread 32 consecutive parameters of ITS OWN sample - the transposed
// staging a CUDA-core version needs does not exist here.  Per tile:
//   * warps 6-7 (A loaders): the NEXT tile's A operand (h tile, K = H + 1 with a constant-one column that multiplies the
//     bias row of W) as fp16 hi / lo images, double-buffered;
//   * consumer warps 0-3 (thread = sample = TMEM lane): the fully unrolled inverse chain (cde_common.cuh), pulling
//     parameters from tensor memory in stream order, 32 per tcgen05.ld, the next load in flight;
//   * warp 4 (one lane): for each block of 128 stream parameters, 3 MMAs per K = 16 step into one of two 128-column
//     accumulators: fp32 parity by the fp16 hi / lo operand split of coupling_tc6.cu (corrections A_lo W_hi, A_hi W_lo
//     first, the main product last), tcgen05.commit -> "block ready";
//   * warp 5 (one lane): streams the packed W blocks (both images of a block contiguous) L2 -> shared-memory ring with
//     cp.async.bulk.
// Measured and rejected: clusters of four CTAs consuming the W stream in lockstep, each fetching a quarter of a block
// and multicasting it (cp.async.bulk ... .multicast::cluster, ring slots released by multicast tcgen05.commit): a
// quarter of the L2 -> SM traffic, but 0.46 instead of 0.26 ms at C4 - the 10 KB copies and the four-way lockstep cost
// more than the traffic saved, so the stream is latency-, not bandwidth-bound (ncu: tensor pipe 30 % of elapsed).
// HBM traffic per sample: H + D + 1 floats.  The tensor pipe does 3 x 2 (H + 1) D_params FLOP per sample (C4: 0.55
// MFLOP, 0.14 TFLOP per 2^18 batch); the kernel is bound by the consumer threads' instruction issue (~4 instructions
// per parameter: tcgen05.ld share, FMA, tanh / exp of the small nets).
#include <cuda_fp16.h>

#include "cde_common.cuh"
#include "tc_common.cuh"

namespace tnf {
namespace cde {

using namespace tnf::tc;

constexpr int kNB = 128;                 // stream parameters (accumulator columns) per block
constexpr int kAcc = 4;                  // accumulator buffers in tensor memory (4 x 128 columns = all of it): the MMA warp runs up
                                         // to three blocks ahead, which also covers the consumers' end-of-tile work
constexpr int kTcThreads = 8 * 32;       // 4 consumer warps, MMA warp, W producer warp, 2 A-loader warps
constexpr int kBulk = 16384;             // bytes per cp.async.bulk

__host__ __device__ inline int kpad(int H) { return (H + 1 + 15) / 16 * 16; }                    // K of the MMAs
__host__ __device__ inline size_t img_bytes(int Kp) { return (size_t)128 * Kp * 2; }            // A: one 128-row K-major image
__host__ __device__ inline size_t wimg_bytes(int Kp) { return (size_t)kNB * Kp * 2; }           // W: one kNB-row K-major image
__host__ __device__ inline size_t block_bytes(int Kp) { return 2 * wimg_bytes(Kp); }             // W block: hi image, lo image

__device__ __forceinline__ uint32_t pack_f16_pair(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void split_f16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_f16_pair(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = pack_f16_pair(a - hf.x, b - hf.y);
}

// W blocks in stream order: block b = [hi image | lo image], image = 128 rows (stream parameters) x Kp (k) K-major
// (img_off); k < H: weight[param][k], k == H: bias[param], beyond: 0; rows beyond the last parameter: 0
template <int D, int U, int L, int STAGES>
__global__ void tc_pack_kernel(ChainDesc c, const float* __restrict__ weight, const float* __restrict__ bias, int H,
                               unsigned char* __restrict__ packed) {
  constexpr int P = chain_params<D, U, L, STAGES>();
  constexpr int nblk = (P + kNB - 1) / kNB;
  const int Kp = kpad(H);
  const int64_t total = (int64_t)nblk * kNB * (Kp / 2);     // one thread per (block, row, k pair)
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int kp = (int)(e % (Kp / 2)), n = (int)((e / (Kp / 2)) % kNB), b = (int)(e / ((int64_t)(Kp / 2) * kNB));
    const int s = b * kNB + n;
    float v[2] = {0.f, 0.f};
    if (s < P) {
      const int64_t p = stream_to_param<D, U, L, STAGES>(s, c);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int k = 2 * kp + i;
        v[i] = k < H ? weight[p * H + k] : (k == H ? bias[p] : 0.f);
      }
    }
    uint32_t hi, lo;
    split_f16_pair(v[0], v[1], hi, lo);
    unsigned char* blk = packed + (size_t)b * block_bytes(Kp);
    *reinterpret_cast<uint32_t*>(blk + img_off(n, 2 * kp, kNB)) = hi;
    *reinterpret_cast<uint32_t*>(blk + wimg_bytes(Kp) + img_off(n, 2 * kp, kNB)) = lo;
  }
}

struct __align__(16) TcCtrl {
  uint64_t w_full[4], w_empty[4];   // W ring
  uint64_t acc_full[kAcc];          // tcgen05.commit: accumulator buffer holds a complete block
  uint64_t acc_empty[kAcc];         // 4 consumer warps: buffer drained
  uint64_t a_ready[2];              // per A buffer, 2 loader warps: the tile's A images written
  uint64_t a_free[2];               // per A buffer, tcgen05.commit: the tile's last MMAs have read the A images
  uint32_t tmem_base, pad;
};

// instruction descriptor: D = f32, A = B = f16, K-major, M = 128, N
__device__ __forceinline__ uint32_t idesc_f16(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

// The 3 * KS MMAs of one block, fully unrolled with constant descriptor offsets (low words only: start address field).
// A single thread retires dependent instructions slowly: the same MMAs from a run-time loop with 64-bit descriptor
// arithmetic cost ~125 issue cycles each (profiles/microbench/mma_loop.cu) against 64 cycles of tensor time, which
// left the tensor pipe 30 % busy and made the MMA warp the kernel's bound.
template <int KS>
__device__ __forceinline__ void issue_block(uint32_t d_tmem, uint32_t ahi_lo, uint32_t alo_lo, uint32_t a_hi32, uint32_t bhi_lo,
                                            uint32_t blo_lo, uint32_t b_hi32, uint32_t idesc) {
  constexpr uint32_t kA = 2 * 128, kB = 2 * kNB;   // descriptor units (16 B) per K = 16 step
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    umma_ss2(d_tmem, alo_lo + kA * ks, a_hi32, bhi_lo + kB * ks, b_hi32, idesc, ks > 0 ? 1u : 0u);
    umma_ss2(d_tmem, ahi_lo + kA * ks, a_hi32, blo_lo + kB * ks, b_hi32, idesc, 1u);
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma_ss2(d_tmem, ahi_lo + kA * ks, a_hi32, bhi_lo + kB * ks, b_hi32, idesc, 1u);
}

// parameter stream of a consumer thread: 32-parameter batches straight from the thread's TMEM lane, double-buffered in
// registers - the tcgen05.ld of batch n + 1 is in flight while batch n is consumed (one consumer warp per SM
// sub-partition: nothing else would hide the load latency)
template <int P>
struct TStream {
  static constexpr bool kFast = true;
  static constexpr int kBatches = (P + 31) / 32;
  uint32_t xb[2][32];
  uint32_t taddr;             // accumulator base + this warp's lane offset
  uint32_t full0, empty0;     // shared addresses of acc_full[0], acc_empty[0]
  uint32_t full_par;          // phase bits of acc_full[0 .. kAcc)
  int lane;
  long long* dbg;             // timing experiment: per-block stamps of one consumer thread (NULL = off)
  int dbg_n;
  template <int B>            // issue the load of batch B (waiting for its block when it opens one)
  __device__ __forceinline__ void issue() {
    constexpr int POS = 32 * B, buf = (POS / kNB) % kAcc;
    if constexpr (POS % kNB == 0) {
      if (dbg && dbg_n < 64) dbg[2 * dbg_n] = clock64();
      mbar_wait_addr(full0 + buf * 8u, (full_par >> buf) & 1u);
      full_par ^= 1u << buf;
      tc_fence_after();
      if (dbg && dbg_n < 64) { dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; }
    }
    tmem_ld32(taddr + (uint32_t)(buf * kNB + POS % kNB), xb[B & 1]);
  }
  template <int POS>
  __device__ __forceinline__ float get() {
    if constexpr (POS % 32 == 0) {
      constexpr int B = POS / 32, buf = (POS / kNB) % kAcc;
      if constexpr (B == 0) issue<0>();
      tc_wait_ld();                                             // batch B is in xb[B & 1]
      if constexpr (POS % kNB == kNB - 32 || B == kBatches - 1) {   // last batch of its block: the MMA warp may overwrite it
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + buf * 8u) : "memory");
      }
      if constexpr (B + 1 < kBatches) issue<B + 1>();
    }
    return __uint_as_float(xb[(POS / 32) & 1][POS % 32]);
  }
};

template <int D, int U, int L, int STAGES>
__global__ void __launch_bounds__(kTcThreads, 1) cde_logprob_tc_kernel(ChainDesc c, const float* __restrict__ h, int H,
                                                                        const unsigned char* __restrict__ packed,
                                                                        const float* __restrict__ z_in, int64_t M,
                                                                        float* __restrict__ out_lp, int n_ring, int dbgbits) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int P = chain_params<D, U, L, STAGES>();
  constexpr int nblk = (P + kNB - 1) / kNB;
  const int Kp = kpad(H);
  const uint32_t blk_bytes = (uint32_t)block_bytes(Kp), image = (uint32_t)img_bytes(Kp), wimage = (uint32_t)wimg_bytes(Kp);
  unsigned char* ring = smem_raw;                                   // n_ring x [hi | lo]
  unsigned char* sA = ring + (size_t)n_ring * blk_bytes;            // 2 buffers x [hi | lo]
  TcCtrl& ct = *reinterpret_cast<TcCtrl*>(sA + 4 * (size_t)image);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); }
    for (int i = 0; i < kAcc; ++i) { mbar_init(&ct.acc_full[i], 1); mbar_init(&ct.acc_empty[i], 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(&ct.a_ready[i], 2); mbar_init(&ct.a_free[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(&ct.tmem_base, kAcc * kNB);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;

  if (warp == 5) {
    // =============================== W producer ===============================
    if (lane == 0) {
      uint32_t slot = 0, par = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int b = 0; b < nblk; ++b) {
          if (dbgbits & 2) continue;   // timing experiment: no W stream
          mbar_wait(&ct.w_empty[slot], par ^ 1u);
          mbar_arrive_expect_tx(&ct.w_full[slot], blk_bytes);
          const unsigned char* src = packed + (size_t)b * blk_bytes;
          unsigned char* dst = ring + (size_t)slot * blk_bytes;
          for (uint32_t o = 0; o < blk_bytes; o += kBulk)
            bulk_g2s(dst + o, src + o, blk_bytes - o < (uint32_t)kBulk ? blk_bytes - o : (uint32_t)kBulk, &ct.w_full[slot]);
          if (++slot == (uint32_t)n_ring) { slot = 0; par ^= 1u; }
        }
      }
    }
  } else if (warp == 4) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t a_par = 0, e_par = 0, it = 0, slot = 0, w_par = 0, gdbg = 0;   // e_par: phase bit per accumulator buffer
      const uint32_t idesc = idesc_f16(kNB);
      const int ksteps = Kp / 16;
      const uint32_t ring_addr = smem_u32(ring), empty0 = smem_u32(&ct.acc_empty[0]), full0 = smem_u32(&ct.acc_full[0]);
      const uint32_t wfull0 = smem_u32(&ct.w_full[0]), wempty0 = smem_u32(&ct.w_empty[0]);
      const uint64_t bdesc0 = make_desc(0u, kNB);
      const uint32_t b_hi32 = (uint32_t)(bdesc0 >> 32), wimage16 = wimage >> 4, blk16 = blk_bytes >> 4;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ab = it & 1u;
        const uint64_t a_hi = make_desc(smem_u32(sA) + ab * 2u * image, 128), a_lo = make_desc(smem_u32(sA) + ab * 2u * image + image, 128);
        const uint32_t a_hi32 = (uint32_t)(a_hi >> 32);
        mbar_wait(&ct.a_ready[ab], (a_par >> ab) & 1u);
        a_par ^= 1u << ab;
#pragma unroll 1
        for (int b = 0; b < nblk; ++b, ++gdbg) {
          const uint32_t buf = (uint32_t)b % kAcc;
          long long* dbgp = ((dbgbits & 4) && blockIdx.x == 0 && gdbg < 64) ? reinterpret_cast<long long*>(out_lp) + 4 * gdbg : nullptr;
          if (dbgp) dbgp[0] = clock64();
          mbar_wait_addr(empty0 + buf * 8u, ((e_par >> buf) & 1u) ^ 1u);   // the consumers drained what this buffer held
          e_par ^= 1u << buf;
          if (dbgp) dbgp[1] = clock64();
          if (!(dbgbits & 2)) mbar_wait_addr(wfull0 + slot * 8u, w_par);
          tc_fence_after();
          if (dbgp) dbgp[2] = clock64();
          const uint32_t d_tmem = tmem + buf * (uint32_t)kNB;
          // descriptor low words: start address >> 4 of this ring slot's hi / lo image, LBO field from bdesc0
          const uint32_t bhi_lo = (uint32_t)bdesc0 + ((ring_addr >> 4) + slot * blk16), blo_lo = bhi_lo + wimage16;
          switch (ksteps) {   // K = 16 steps of the MMAs: (H + 1) padded to 16; common widths unrolled
#define TNF_KS(N) case N: issue_block<N>(d_tmem, (uint32_t)a_hi, (uint32_t)a_lo, a_hi32, bhi_lo, blo_lo, b_hi32, idesc); break;
            TNF_KS(1) TNF_KS(2) TNF_KS(3) TNF_KS(4) TNF_KS(5) TNF_KS(6) TNF_KS(7) TNF_KS(8) TNF_KS(9) TNF_KS(10) TNF_KS(12) TNF_KS(16)
#undef TNF_KS
            default: {
              constexpr uint32_t kA = 2 * 128, kB = 2 * kNB;
              for (int ks = 0; ks < ksteps; ++ks) {
                umma_ss2(d_tmem, (uint32_t)a_lo + kA * ks, a_hi32, bhi_lo + kB * ks, b_hi32, idesc, ks > 0 ? 1u : 0u);
                umma_ss2(d_tmem, (uint32_t)a_hi + kA * ks, a_hi32, blo_lo + kB * ks, b_hi32, idesc, 1u);
              }
              for (int ks = 0; ks < ksteps; ++ks) umma_ss2(d_tmem, (uint32_t)a_hi + kA * ks, a_hi32, bhi_lo + kB * ks, b_hi32, idesc, 1u);
            }
          }
          if (!(dbgbits & 2)) tc_commit_addr(wempty0 + slot * 8u);
          tc_commit_addr(full0 + buf * 8u);
          if (dbgp) dbgp[3] = clock64();
          if (++slot == (uint32_t)n_ring) { slot = 0; w_par ^= 1u; }
        }
        tc_commit(&ct.a_free[ab]);
      }
    }
  } else if (warp >= 6) {
    // =============================== A loaders: the h tile of the NEXT tile as fp16 hi / lo images ===============================
    // unit = (row, group of 8 k): consecutive threads take consecutive groups of a row (coalesced reads of h); row r =
    // [h[m, 0..H), 1, 0...]; one 16-byte store per image and unit
    const int t = threadIdx.x - 6 * 32, g8 = Kp / 8;
    uint32_t it = 0, f_par = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ab = it & 1u;
      if (it >= 2) {   // the MMAs of the tile that used this buffer two tiles ago are done with it
        mbar_wait(&ct.a_free[ab], (f_par >> ab) & 1u);
        f_par ^= 1u << ab;
      }
      unsigned char* a_hi = sA + (size_t)ab * 2 * image;
      unsigned char* a_lo = a_hi + image;
      const int64_t m0 = tile * 128;
#pragma unroll 4
      for (int u = t; u < 128 * g8; u += 64) {   // (10 units in flight measured slower: 0.197 vs 0.180 ms at C4)
        const int r = u / g8, k8 = u - r * g8;
        const bool valid = m0 + r < M;
        const float* hrow = h + (m0 + r) * H;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = 8 * k8 + i;
          v[i] = k < H ? (valid ? __ldg(hrow + k) : 0.f) : (k == H ? 1.0f : 0.f);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_f16_pair(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
        *reinterpret_cast<uint4*>(a_hi + (size_t)k8 * 2048 + r * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(a_lo + (size_t)k8 * 2048 + r * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.a_ready[ab]);
    }
  } else {
    // =============================== consumers: thread = sample = TMEM lane ===============================
    TStream<P> S;
    S.taddr = tmem + ((uint32_t)(warp * 32) << 16);
    S.full0 = smem_u32(&ct.acc_full[0]);
    S.empty0 = smem_u32(&ct.acc_empty[0]);
    S.full_par = 0;
    S.lane = lane;
    S.dbg = ((dbgbits & 4) && blockIdx.x == 0 && threadIdx.x == 0) ? reinterpret_cast<long long*>(out_lp) + 512 : nullptr;
    S.dbg_n = 0;
    const int r = threadIdx.x;
    float zn[D];   // the NEXT tile's sample, loaded a tile ahead (its global-load latency would stall this warp)
    {
      const int64_t m0 = (int64_t)blockIdx.x * 128 + r;
#pragma unroll
      for (int d = 0; d < D; ++d) zn[d] = m0 < M ? z_in[m0 * D + d] : 0.f;
    }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t m = tile * 128 + r;
      const bool valid = m < M;
      float z[D];
#pragma unroll
      for (int d = 0; d < D; ++d) z[d] = zn[d];
      {
        const int64_t mn = (tile + gridDim.x) * 128 + r;
#pragma unroll
        for (int d = 0; d < D; ++d) zn[d] = mn < M ? z_in[mn * D + d] : 0.f;
      }
      float lp;
      if (dbgbits & 1) {   // timing experiment: the stream alone, no chain arithmetic
        lp = z[0];
        static_for<0, (P + 31) / 32>([&](auto bi) { lp += S.template get<32 * bi>(); });
      } else {
        lp = chain_logprob<D, U, L, STAGES>(z, c, S);
      }
      if (valid && !(dbgbits & 4)) out_lp[m] = lp;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, kAcc * kNB);
}

size_t tc_packed_bytes(int64_t D_params, int H) {
  const int64_t nblk = (D_params + kNB - 1) / kNB;
  return (size_t)nblk * block_bytes(kpad(H));
}

int tc_pack(const ChainDesc& c, int D, int U, const float* weight, const float* bias, int H, void* packed, cudaStream_t st) {
#define X(DV, UV) if (D == DV && U == UV) tc_pack_kernel<DV, UV, 2, 1><<<num_sms() * 2, 256, 0, st>>>(c, weight, bias, H, (unsigned char*)packed);
  TNF_CDE_SHAPES(X)
#undef X
  return check_launch("tnf_cde_pack");
}

int tc_logprob(const ChainDesc& c, int D, int U, const float* h, int H, const void* packed, const float* z, int64_t M,
               float* log_prob, int dbgbits, cudaStream_t st) {
  const int Kp = kpad(H);
  // one CTA per SM (255 registers per consumer thread, no spills); the W ring as deep as shared memory allows.
  // Measured alternative (64-column blocks, two CTAs per SM at 168 registers with spills): C4 0.29 vs 0.32 ms, C2b 0.20
  // vs 0.15 ms per call - the consumers wait on the W stream either way.
  int n_ring = 4;
  auto smem_of = [&](int ring) { return (size_t)ring * block_bytes(Kp) + 4 * img_bytes(Kp) + sizeof(TcCtrl) + 1024; };
  while (n_ring > 2 && smem_of(n_ring) > 227 * 1024) --n_ring;
  const size_t smem = smem_of(n_ring);
  TNF_REQUIRE(smem <= 227 * 1024, TNF_ERR_UNSUPPORTED, "tnf_cde_logprob: H = %d needs %zu B shared memory", H, smem);
  const int64_t n_tiles = (M + 127) / 128;
  const int per_sm = 1;
  const int grid = (int)(n_tiles < (int64_t)per_sm * num_sms() ? n_tiles : (int64_t)per_sm * num_sms());
  cudaError_t e = cudaSuccess;
#define X(DV, UV)                                                                                                         \
  if (D == DV && U == UV) {                                                                                               \
    e = cudaFuncSetAttribute(cde_logprob_tc_kernel<DV, UV, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess)                                                                                                 \
      cde_logprob_tc_kernel<DV, UV, 2, 1><<<grid, kTcThreads, smem, st>>>(c, h, H, (const unsigned char*)packed, z, M, log_prob, n_ring, dbgbits); \
  }
  TNF_CDE_SHAPES(X)
#undef X
  if (e != cudaSuccess) { set_error("tnf_cde_logprob: %s", cudaGetErrorString(e)); return (int)e; }
  return check_launch("tnf_cde_logprob");
}

}  // namespace cde
}  // namespace tnf
