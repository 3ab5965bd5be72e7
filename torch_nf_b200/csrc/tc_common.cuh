// Shared definitions of the tensor-core coupling kernels (sm_100a): tile constants, the packed-weight
// layout (Shape), raw PTX wrappers for mbarrier / cp.async.bulk / tcgen05 / clusters, UMMA descriptors.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace tnf {
namespace tc {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kStageBytes = 16384;
constexpr int kStageElems = kStageBytes / 2;
constexpr int kMaxStages = 8;
constexpr int kChunk = 32;  // accumulator columns handled per epilogue step

struct Shape {
  int D, U, L, upper;
  int d_in, d_out, c_off, t_off;
  int Nh, nh;  // hidden layers are issued as nh column blocks of width Nh
  __host__ __device__ Shape(int D_, int U_, int L_, int upper_) : D(D_), U(U_), L(L_), upper(upper_) {
    int h = D / 2;
    d_in = h; d_out = h;
    c_off = upper ? 0 : h;
    t_off = upper ? h : 0;
    // Full-width jobs: splitting a layer into column halves would let the epilogue of the low half
    // overwrite the (single) A image while the high-half MMAs still read it.
    Nh = U;
    nh = 1;
  }
  __host__ __device__ int K_of(int l) const { return l == 0 ? d_in : U; }
  __host__ __device__ int J_of(int l) const { return l == L ? d_out : U; }
  __host__ __device__ int N_of(int l) const { return l == L ? d_out : Nh; }      // MMA N of one job
  __host__ __device__ int halves(int l) const { return l == L ? 1 : nh; }
  // weight stage capacity: 16 KB, or 8 KB when the A images leave too little shared memory (D = 256, U = 256)
  __host__ __device__ int stage_elems() const { return (D >= 256 && U >= 256) ? kStageElems / 2 : kStageElems; }
  // K rows of one weight stage for an N-wide layer
  __host__ __device__ int stage_k(int K, int N) const {
    int ks = stage_elems() / N;
    return ks < K ? ks : K;
  }
  __host__ __device__ int64_t net_weight_elems() const {
    return (int64_t)d_in * U + (int64_t)(L - 1) * U * U + (int64_t)U * d_out;
  }
  __host__ __device__ int net_bias_elems() const { return L * U + d_out; }
  // bias operand images (two-tile kernel): per (layer, net) one K=16 group of J columns, K-row 0 = bf16(b),
  // K-row 1 = bf16(b - bf16(b)), rest 0; multiplied by a constant [1, 1, 0, ...] A image the MMA adds the bias
  __host__ __device__ int64_t bias_img_bytes() const { return 2 * (int64_t)net_bias_elems() * 32; }
  __host__ __device__ int64_t bias_img_off(int l, int net) const {
    int64_t off = 2 * net_weight_elems() * 2 + 2 * (int64_t)net_bias_elems() * 4;
    for (int i = 0; i < l; ++i) off += 2 * (int64_t)J_of(i) * 32;
    return off + (int64_t)net * J_of(l) * 32;
  }
  __host__ __device__ int64_t unsplit_bytes() const {
    return 2 * net_weight_elems() * 2 + 2 * (int64_t)net_bias_elems() * 4 + bias_img_bytes();
  }
  // CTA-pair kernel: the same weight and bias operand images cut in two along N (J/2 columns per CTA), each half
  // contiguous so that one cp.async.bulk fetches a CTA's part of a stage: for layer, net, rank: [K x J/2] image
  __host__ __device__ int64_t split_w_off(int l, int net, int rank) const {
    int64_t off = unsplit_bytes();
    for (int i = 0; i < l; ++i) off += 2 * (int64_t)K_of(i) * J_of(i) * 2;
    return off + (int64_t)net * K_of(l) * J_of(l) * 2 + (int64_t)rank * K_of(l) * (J_of(l) / 2) * 2;
  }
  __host__ __device__ int64_t split_b_off(int l, int net, int rank) const {
    int64_t off = unsplit_bytes() + 2 * net_weight_elems() * 2;
    for (int i = 0; i < l; ++i) off += 2 * (int64_t)J_of(i) * 32;
    return off + (int64_t)net * J_of(l) * 32 + (int64_t)rank * (J_of(l) / 2) * 32;
  }
  __host__ __device__ int64_t pair_bytes() const { return unsplit_bytes() + 2 * net_weight_elems() * 2 + bias_img_bytes(); }
  // coupling_tc5_kernel (L >= 2): the LAST hidden layer (l = L-1) is issued as two N = U/2 halves: per net, half h,
  // rank r one [U x U/4] weight image (units U/2*h + U/4*r ..)
  __host__ __device__ int64_t half_w_off(int net, int h, int rank) const {
    return pair_bytes() + (int64_t)((net * 2 + h) * 2 + rank) * U * (U / 4) * 2;
  }
  // ... and the biases stay RESIDENT in shared memory as 8-K-row blocks (unit n at byte 16*n: k = 0 bf16(b),
  // k = 1 bf16(b - bf16(b)), k = 2..7 zero), per rank one contiguous region: for net: for layer: block(s) of this
  // rank's units (the split layer: half a block, half b block)
  __host__ __device__ int64_t bias8_rank_bytes() const { return 2 * 16 * ((int64_t)L * (U / 2) + d_out / 2); }
  __host__ __device__ int64_t bias8_off(int l, int net, int h) const {
    int64_t off = (int64_t)net * (bias8_rank_bytes() / 2);
    for (int i = 0; i < l; ++i) off += (int64_t)(J_of(i) / 2) * 16;
    return off + (int64_t)h * (J_of(l) / 4) * 16;
  }
  __host__ __device__ int64_t bias8_base(int rank) const {
    return pair_bytes() + 2 * (int64_t)U * U * 2 + (int64_t)rank * bias8_rank_bytes();
  }
  __host__ __device__ int64_t half_bytes() const { return L >= 2 ? 2 * (int64_t)U * U * 2 + 2 * bias8_rank_bytes() : 0; }
  __host__ __device__ int64_t packed_bytes() const { return pair_bytes() + half_bytes(); }
  __host__ __device__ size_t a1_bytes() const { return (size_t)kTileM * d_in * 2; }
  __host__ __device__ size_t act_bytes() const { return (size_t)kTileM * U * 2; }
};

__host__ __device__ inline bool shape_supported(int D, int U, int L) {
  if (!(D == 64 || D == 128 || D == 256)) return false;  // d_in = d_out = D/2 in {32, 64, 128}
  if (!(U == 64 || U == 128 || U == 256)) return false;
  if (L < 1 || L > 5) return false;
  return true;
}

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]   (M=128, K=16, bf16 -> fp32)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors passed as (low, high) 32-bit words: only the low word (start address, LBO) changes per step
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(bar_addr), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
}
// ---- CTA pairs (clusters of two, tcgen05 cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of both CTAs
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA's layout) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// Arrival on a barrier of the partner (or own) CTA through the cluster address space.  RELAXED on purpose: a
// release.cluster arrival costs ~500 cycles here (0.54 vs 0.47 ms per launch) and orders nothing we need - what the
// arriving warp published (activation / A1 images) is read by its OWN SM's tensor core and was made visible to the
// async proxy by fence.proxy.async before this message is even sent over the SM-to-SM network; the stage-landed
// forward has no memory operations of its own to order.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma2_ss2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// pair MMA with the A operand in tensor memory (each CTA's own 128 rows at the same TMEM address)
__device__ __forceinline__ void umma2_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit2_addr(uint32_t bar_addr) {   // arrives on this barrier in BOTH CTAs of the pair
  const uint16_t mask = 0x3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   bar_addr), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]  (kept for the TMEM-A diagnostic)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// instruction descriptor: dense, D=f32, A=B=bf16, both K-major, M=128, N
__host__ __device__ inline uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
// the same for a CTA pair: M = 256 (128 rows per CTA)
__host__ __device__ inline uint32_t make_idesc2(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE.  An operand image of R rows x K columns is
// stored as K/8 blocks of R x 16 bytes (row r at byte r*16 inside a block): core matrix = 8 rows x
// 16 bytes contiguous (128 B), SBO = 128 B between 8-row groups, LBO = R*16 B between the 8-column
// K groups.  One K=16 MMA step reads two consecutive K groups starting at `smem_addr`.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int rows) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(((uint32_t)rows * 16u >> 4) & 0x3FFF) << 16;  // LBO
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;                   // SBO
  d |= (uint64_t)1 << 46;                                        // descriptor version (sm_100)
  return d;
}
// byte offset of element (row, k) inside an operand image with `rows` rows
__host__ __device__ inline uint32_t img_off(int row, int k, int rows) {
  return (uint32_t)((k >> 3) * rows * 16 + row * 16 + (k & 7) * 2);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// 256-bit global accesses (sm_100): a thread that owns 8+ consecutive floats of a row moves them in half the
// instructions; with one row per lane every lane of such an instruction is in another 128-byte line, so the LSU pipe
// pays per instruction (32 wavefronts each)
__device__ __forceinline__ void ldg256_nc(const float* p, float* r) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* r) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r[0]), "f"(r[1]), "f"(r[2]),
               "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7])
               : "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh on the FMA pipe: odd minimax polynomial x*P(x^2) on |x| <= 3.3, input clamped (beyond it tanh rounds to +-1 in
// bf16 anyway).  Max abs error 1.4e-3 (bf16 rounding of the result: up to 2e-3); 10 FMA-pipe instructions against 8
// issue cycles of the 16-lane MUFU pipe, used for a share of the activations to take load off that pipe.
__device__ __forceinline__ float tanh_poly(float x) {
  const float xc = fminf(fmaxf(x, -3.3f), 3.3f);
  const float t = xc * xc;
  float p = fmaf(2.4607425075373612e-06f, t, -0.00010122240928467363f);
  p = fmaf(p, t, 0.0017052673501893878f);
  p = fmaf(p, t, -0.015382171608507633f);
  p = fmaf(p, t, 0.08308329433202744f);
  p = fmaf(p, t, -0.29954469203948975f);
  p = fmaf(p, t, 0.9929457902908325f);
  return p * xc;
}
// low half = first (lower K index) element
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- the fused kernel
struct Args {
  const float* z_in; float* z_out; float* log_det; const unsigned char* packed;
  const float* pre_scale; const float* pre_shift;
  int64_t rows;
  int D, U, L, upper, inverse, accum, n_stages, n_groups;
  int tune;   // diagnostic knob of the kernel under development (bits 8.. of `variant`); 0 = shipped configuration
  double* stat_partials;   // [grid*8 warps][2][D] per-warp column sums of the OUTPUT (NULL = off)
  long long* dbg;   // diagnostics: per-phase clock64 stamps of CTA 0 (NULL = off)
  // fused base density (inverse direction, the chain's LAST executed layer; coupling_tc5 / coupling_tc6 kernels):
  // out_lp[row] = -1/2 sum_d z_out^2 - D log sqrt(2 pi) - (log_det[row] + sum s) - lp_scal[0]; z_out / log_det are not written
  float* out_lp;
  const float* lp_scal;
};

struct __align__(16) Ctrl {
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t a1_ready[2];  // per group, 4 epilogue warps: A1 image written, last tile's outputs drained
  uint64_t e_done[2];    // per group, 4 epilogue warps: accumulator drained (and next A image written)
  uint64_t h_ready[2][2];  // per group and column half, MMA commit: accumulator (half) complete
  uint32_t tmem_base;
  uint32_t pad;
};
// dynamic shared memory:
//   [ring: n_stages x 16 KB][A1 g0][A1 g1][Act g0][Act g1][Ctrl][bias 2 x nb][pre_scale D][pre_shift D]

__host__ __device__ inline size_t smem_bytes(const Shape& sh, int n_stages) {
  return (size_t)n_stages * sh.stage_elems() * 2 + 2 * sh.a1_bytes() + 2 * sh.act_bytes() + sizeof(Ctrl) +
         (size_t)(2 * sh.net_bias_elems() + 2 * sh.D) * sizeof(float);
}

// Sum over the 32 lanes of v[j] for every j, by recursive halving: 31 shuffles; lane l returns column l's sum.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__device__ __forceinline__ float exp2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int kEpiWarps2 = 16;   // 4 per SM sub-partition
constexpr int kOnesBytes = 2 * kTileM * 16;   // A image of one K=16 step: [1, 1, 0, ..., 0] in every row

__host__ __device__ inline bool shape_supported2(int D, int U, int L) {
  return shape_supported(D, U, L) && D <= 128;
}
// column sums of v[0..15] over the 32 lanes: lane l returns the sum of column l % 16 (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = hi ? v[i] : v[i + off];
      const float keep = hi ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// coupling_tc4.cu: launches the CTA-pair two-tile kernel (the product path for D <= 128); returns cudaError_t
int launch_tc4(const Args& a, int grid, size_t smem, cudaStream_t st);
size_t smem_bytes4(const Shape& sh, int n_stages);
// coupling_tc5.cu: tc4 with the last hidden job issued as two N-halves and the final layer fed from tensor memory
int launch_tc5(const Args& a, int grid, size_t smem, cudaStream_t st);
size_t smem_bytes5(const Shape& sh, int n_stages);
// coupling_tc6.cu: fp32-parity mode (fp16 hi / lo operand split, three MMAs per product, activations in TMEM)
bool shape_supported6(int D, int U, int L);
size_t packed_bytes6(int D, int U, int L, int split);
int pack6_launch(const float* params, void* packed, int D, int U, int L, int upper, int split, cudaStream_t st);
size_t smem_bytes6(int D, int U, int L, int split, int n_stages);
int launch_tc6(const Args& a, int grid, int split, size_t smem, cudaStream_t st);
__host__ __device__ inline bool shape_supported5(int D, int U, int L) {
  return shape_supported2(D, U, L) && U >= 128 && L == 2;
}

}  // namespace tc
}  // namespace tnf
