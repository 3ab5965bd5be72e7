// RealNVP coupling layer on tcgen05 tensor cores: the CTA-pair two-tile kernel with
//   * the LAST hidden job (layer L-1, L >= 2) issued as two N = U/2 halves, each with its own commit: the group's
//     tanh phase starts on the first half of the accumulator while the tensor pipe computes the second half;
//   * the output of that tanh phase written back as bf16 INTO the accumulator columns it has just consumed
//     (tcgen05.st) and the final-layer MMAs fed from tensor memory (A operand in TMEM): no st.shared / shared-memory
//     A reads for that layer, and no write-after-read hazard on the activation image the second half still reads.
// Everything else is coupling_tc4_kernel (coupling_tc4.cu).
// Shared definitions: tc_common.cuh.  Reference semantics: torch_nf/bijectors.py:145-242 (RealNVP).
#include "tc_common.cuh"

namespace tnf {
namespace tc {

// TMEM column (relative to the group's accumulator) of the 8 packed bf16 columns holding the activations of units
// [16*kk, 16*kk + 16) after the last tanh phase: chunk c = kk/2 of parity p = ci%2 (ci = index inside its N-half h)
// was written by its own warp to (U/2)*h + 32*p + 16*(ci/2)  -- columns that warp had already consumed.
template <int U_>
__host__ __device__ constexpr uint32_t act_col5(int kk) {
  const int nc = U_ / kChunk, c = kk / 2, sub = kk % 2;
  const int h = c / (nc / 2), ci = c % (nc / 2), p = ci % 2, i = ci / 2;
  return (uint32_t)((U_ / 2) * h + 32 * p + 16 * i + 8 * sub);
}
template <int U_>
__host__ __device__ constexpr uint32_t fin_col5() { return U_ == 256 ? 64u : 128u; }   // free columns for the final accumulator

constexpr uint32_t kStages5 = 4;   // weight ring depth (compile-time: slot = counter & 3)

// One GEMM job of the MMA warp (see mma_job in coupling_tc.cu): the bias MMA first (accumulator := ones image x the
// job's RESIDENT bias block, no barrier), then per 32-wide K chunk two K=16 MMAs with compile-time descriptor offsets,
// an mbarrier wait pair (own stage landed / partner's stage landed) when a weight stage begins and a commit when it
// is used up.  `cnt` counts ring stages since kernel start: slot = cnt & 3, parity = (cnt >> 2) & 1.
template <int K, int N, bool kTS, int U_>
__device__ __forceinline__ void mma_job5(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo_ring, uint32_t b_hi,
                                         uint32_t ones_lo, uint32_t bias_lo, uint32_t wfull0, uint32_t wpeer0,
                                         uint32_t wempty0, uint32_t& cnt, bool leader, long long* t_w) {
  constexpr int NB = N / 2;                                            // B rows held per CTA
  constexpr int KS = (kStageElems / NB) < K ? (kStageElems / NB) : K;  // K rows per weight stage
  constexpr int CPS = KS / kChunk;                                     // chunks per stage
  constexpr uint32_t kStage16 = kStageBytes >> 4;
  const uint32_t idesc = make_idesc2(N);
  if (leader) umma2_ss2(d_tmem, ones_lo, a_hi, bias_lo, b_hi, idesc, 0u);
#pragma unroll
  for (int c = 0; c < K / kChunk; ++c) {
    const uint32_t slot = cnt & (kStages5 - 1);
    if (c % CPS == 0) {
      const long long c0 = t_w ? clock64() : 0;
      const uint32_t par = (cnt >> 2) & 1u;
      mbar_wait_addr(wfull0 + slot * 8u, par);
      mbar_wait_addr(wpeer0 + slot * 8u, par);
      if (t_w) *t_w += clock64() - c0;
      tc_fence_after();
    }
    if (leader) {
      const uint32_t b_lo = b_lo_ring + slot * kStage16 + (uint32_t)((c % CPS) * 4 * NB);
      if (kTS) {   // A operand = the bf16 activations the last tanh phase left in the accumulator's own columns
        umma2_ts2(d_tmem, a_lo + act_col5<U_>(2 * c), b_lo, b_hi, idesc, 1u);
        umma2_ts2(d_tmem, a_lo + act_col5<U_>(2 * c + 1), b_lo + 2u * NB, b_hi, idesc, 1u);
      } else {
        umma2_ss2(d_tmem, a_lo + 512u * c, a_hi, b_lo, b_hi, idesc, 1u);
        umma2_ss2(d_tmem, a_lo + 512u * c + 256u, a_hi, b_lo + 2u * NB, b_hi, idesc, 1u);
      }
      if (c % CPS == CPS - 1) tc_commit2_addr(wempty0 + slot * 8u);
    }
    if (c % CPS == CPS - 1) ++cnt;
  }
}

// ================================================================ two-tile ping-pong kernel on CTA pairs (D <= 128)
// The two-tile kernel above, run by clusters of two CTAs that share every weight operand (tcgen05 cta_group::2):
// the pair's MMAs have M = 256 (128 rows = one tile per CTA and group), each CTA keeps only ITS half of the N
// columns of a weight stage in shared memory, and the tensor cores of both SMs read both halves.  Per SM this
// halves the weight stream from L2, the ring bytes per stage (a 16 KB slot now holds 64 K rows = four MMAs, so the
// four-slot ring is twice as deep in MMA time) and the B-operand shared-memory reads.  The leader CTA (cluster
// rank 0) issues all MMAs; commits are multicast to both CTAs' barriers; the epilogue and I/O warps of the second
// CTA arrive on the leader's barriers through the cluster address space, and its otherwise idle MMA warp forwards
// "weight stage landed" from its own ring to the leader.  (Semantics checked in profiles/microbench/umma_2cta.cu.)
// Two tiles in flight per CTA.  Epilogue group g (8 warps: TMEM lane quadrant w%4, accumulator chunks of parity
// (w/4)%2) owns tile (2k+g)*grid + cta, the 256 TMEM columns [256g, 256g+256) and its own A1 / activation images.
// A layer's MMAs start when the group's previous epilogue phase is complete (one accumulator per group: the next
// layer would overwrite what the epilogue is still reading), so a single group alternates between the MUFU pipe and
// the tensor pipe - and the other group, half a tile out of phase, fills each pipe in the gaps.  Weights and the
// bias operand images stream through the ring in the static job order of the MMA warp; biases are added by a bias
// MMA (see mma_job), global I/O of the conditioning half is done by two dedicated warps with coalesced accesses.
constexpr int kThreads5 = (kEpiWarps2 + 4) * 32;   // 16 epilogue, MMA, producer, 2 I/O

struct __align__(16) Ctrl5 {
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t w_peer[kMaxStages];   // leader only: the second CTA's half of the stage has landed (forwarded by its MMA warp)
  uint64_t a1_ready[2];   // per group, 2 I/O warps: A1 image of the group's next tile written
  uint64_t a1_free[2];    // per group, tcgen05.commit: both layer-0 jobs of the tile have read the A1 image
  uint64_t e_done[2];     // per group, 8 epilogue warps: accumulator drained (and activation image written)
  uint64_t h_ready[2];    // per group, tcgen05.commit: accumulator of the group's current job complete
  uint64_t h_ready_b[2];  // per group, tcgen05.commit: second N-half of the split job complete
  uint64_t y_done[2];     // per group, this CTA's 8 epilogue warps: the tile's transformed half is stored (fused statistics)
  uint32_t tmem_base;
  uint32_t pad;
};
// dynamic shared memory:
//   [ring: 4 x 16 KB][A1 g0][A1 g1][Act g0][Act g1][ones 4 KB][Ctrl5][pre_scale D][pre_shift D][ld partial 2 x 128]
//   [sum z^2 of the conditioning half: group x tile parity x 128 (fused base density)][resident bias blocks of this rank][2 KB zeros]
constexpr int kBiasPad5 = 2048;   // zero block after the resident bias region (second K group of the last bias MMA)
size_t smem_bytes5(const Shape& sh, int n_stages) {
  return (size_t)n_stages * sh.stage_elems() * 2 + 2 * sh.a1_bytes() + 2 * sh.act_bytes() + kOnesBytes + sizeof(Ctrl5) +
         (size_t)(2 * sh.D + 2 * kTileM + 4 * kTileM) * sizeof(float) + (size_t)sh.bias8_rank_bytes() + kBiasPad5;
}

// FT: of every 8 activations, FT are evaluated by tanh_poly on the FMA pipe, the rest by MUFU.TANH
template <bool kInverse, int DH, int U_, int FT>   // DH = D/2 = d_in = d_out in {32, 64}; U_ = hidden units; L = 2
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads5, 1) coupling_tc5_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int L_ = 2;
  const Shape sh(a.D, a.U, L_, a.upper);
  constexpr int S = (int)kStages5;
  unsigned char* ring = smem_raw;
  const uint32_t stage_bytes = (uint32_t)sh.stage_elems() * 2;
  unsigned char* sA1 = ring + (size_t)S * stage_bytes;             // 2 images (group)
  unsigned char* sAct = sA1 + 2 * sh.a1_bytes();                    // 2 images (group)
  unsigned char* sOnes = sAct + 2 * sh.act_bytes();                 // constant A image for the bias MMA
  Ctrl5& ct = *reinterpret_cast<Ctrl5*>(sOnes + kOnesBytes);
  float* s_pscale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(Ctrl5));
  float* s_pshift = s_pscale + sh.D;
  float* s_ldp = s_pshift + sh.D;
  float* s_ss = s_ldp + 2 * kTileM;                                               // [group][tile parity][128]
  unsigned char* sBias = reinterpret_cast<unsigned char*>(s_ss + 4 * kTileM);    // resident bias blocks (+ zero pad)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const uint32_t rank = cluster_ctarank();          // 0 = leader
  const int64_t P = gridDim.x / 2, pair = blockIdx.x / 2;
  const int64_t n_super = (n_tiles + 1) / 2;          // 256-row super tiles: one tile per CTA of the pair
  // tiles of group g: 2 * ((2k + g) * P + pair) + rank, k = 0 .. cnt[g] - 1 (the same count in both CTAs)
  int64_t cnt[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int64_t first = (int64_t)g * P + pair;
    cnt[g] = first < n_super ? (n_super - first + 2 * P - 1) / (2 * P) : 0;
  }
  constexpr int n_chunks = U_ / kChunk;
  constexpr int JT = 2 * (L_ + 1);   // jobs per tile: t0 t1 tF s0 s1 sF
  constexpr int SH = L_;             // group 1 runs this many jobs behind group 0 (shifts 1..5 measured, 2 is best)

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); mbar_init(&ct.w_peer[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&ct.a1_ready[g], 4);                  // the I/O warps of both CTAs (leader's barrier)
      mbar_init(&ct.a1_free[g], 1);
      mbar_init(&ct.e_done[g], kEpiWarps2);           // the group's epilogue warps of both CTAs (leader's barrier)
      mbar_init(&ct.h_ready[g], 1);
      mbar_init(&ct.h_ready_b[g], 1);
      mbar_init(&ct.y_done[g], kEpiWarps2 / 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps2) tmem_alloc2(&ct.tmem_base, 512);
  {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x)   // row r: K columns 0 and 1 are 1.0 (bf16 0x3f80)
      reinterpret_cast<uint32_t*>(sOnes)[i] = (i < kTileM * 4 && (i & 3) == 0) ? 0x3f803f80u : 0u;
    fence_async_smem();
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
    const int nb16 = (int)(sh.bias8_rank_bytes() / 16);
    const uint4* gb = reinterpret_cast<const uint4*>(a.packed + sh.bias8_base((int)cluster_ctarank()));
    for (int i = threadIdx.x; i < nb16 + kBiasPad5 / 16; i += blockDim.x)
      reinterpret_cast<uint4*>(sBias)[i] = i < nb16 ? __ldg(gb + i) : make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers and images exist before anything crosses the pair
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;
  // the leader's barriers that both CTAs arrive on, as cluster addresses
  const uint32_t lead_e_done = mapa_u32(smem_u32(&ct.e_done[0]), 0u), lead_a1_ready = mapa_u32(smem_u32(&ct.a1_ready[0]), 0u);
  const uint32_t lead_w_peer = mapa_u32(smem_u32(&ct.w_peer[0]), 0u);
  const bool want_stats = a.stat_partials != nullptr;
  const bool lp_mode = kInverse && a.out_lp != nullptr;   // fused base density: this is the chain's last executed layer
  // Column statistics of the OUTPUT are taken by the two I/O warps (a lane owns four columns of each half, sums in
  // registers): the conditioning half as it passes through, the transformed half read back (L2) once the epilogue warps
  // have stored a tile - they have a tile of slack, the epilogue warps are the kernel's critical path.
  float io_sv1[4] = {0.f, 0.f, 0.f, 0.f}, io_sv2[4] = {0.f, 0.f, 0.f, 0.f};   // conditioning half: columns col..col+3
  float io_tv1[4] = {0.f, 0.f, 0.f, 0.f}, io_tv2[4] = {0.f, 0.f, 0.f, 0.f};   // transformed half: columns t_off+hc..+3
  // per-I/O-warp statistics rows ([2*D] doubles) are gathered in the (then dead) activation images
  double* stat_rows = reinterpret_cast<double*>(sAct);
  constexpr int kStatWarps = 2;

  // Static job schedule (L = 2): per tile and group six jobs j = 0..5 = t0 t1 tF s0 s1 sF (net j/3, layer j%3); step
  // n = it*6 + s serves group 0's job s of its tile `it`, then group 1's job (s-2) mod 6 of its tile it (s >= 2) or
  // it-1.  The three control warps walk the same sequence with the job type a compile-time constant.
#define TNF_FOR_JOBS(...)                                                               \
  for (int64_t it = 0; it <= cnt[0]; ++it) {                                            \
    _Pragma("unroll") for (int s = 0; s < JT; ++s) {                                    \
      if (it < cnt[0]) { constexpr int g = 0; const int j = s; __VA_ARGS__ }                   \
      {                                                                                 \
        const int64_t t1 = s >= SH ? it : it - 1;                                       \
        if (t1 >= 0 && t1 < cnt[1]) { constexpr int g = 1; const int j = (s + JT - SH) % JT; __VA_ARGS__ } \
      }                                                                                 \
    }                                                                                   \
  }
  if (warp == kEpiWarps2 + 1) {
    // =============================== weight producer (one elected lane) ===============================
    // streams THIS CTA's half of every weight stage; biases are resident.  Stages per job: t0/s0 one (DH x U/2), t1/s1
    // 2 x K/KSh (two N-halves of U x U/4), tF/sF one (U x DH/2)
    if (elect_one()) {
      uint32_t cntp = 0;
      auto push = [&](const unsigned char* src, uint32_t bytes) {
        const uint32_t slot = cntp & (kStages5 - 1), par = (cntp >> 2) & 1u;
        mbar_wait(&ct.w_empty[slot], par ^ 1u);
        mbar_arrive_expect_tx(&ct.w_full[slot], bytes);
        bulk_g2s(ring + (size_t)slot * kStageBytes, src, bytes, &ct.w_full[slot]);
        ++cntp;
      };
      const unsigned char* w0[2] = {a.packed + sh.split_w_off(0, 0, (int)rank), a.packed + sh.split_w_off(0, 1, (int)rank)};
      const unsigned char* wF[2] = {a.packed + sh.split_w_off(L_, 0, (int)rank), a.packed + sh.split_w_off(L_, 1, (int)rank)};
      const unsigned char* wH[2][2] = {{a.packed + sh.half_w_off(0, 0, (int)rank), a.packed + sh.half_w_off(0, 1, (int)rank)},
                                       {a.packed + sh.half_w_off(1, 0, (int)rank), a.packed + sh.half_w_off(1, 1, (int)rank)}};
      constexpr int KSH = (kStageElems / (U_ / 4)) < U_ ? (kStageElems / (U_ / 4)) : U_;   // K rows per stage of an N-half
      constexpr uint32_t kHB = (uint32_t)(KSH * (U_ / 4) * 2);
      TNF_FOR_JOBS({
        (void)g;
        const int net = j / 3, l = j % 3;
        if (l == 0) push(w0[net], (uint32_t)(DH * (U_ / 2) * 2));
        else if (l == 1) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int st = 0; st < U_ / KSH; ++st) push(wH[net][h] + (size_t)st * kHB, kHB);
        } else push(wF[net], (uint32_t)(U_ * (DH / 2) * 2));
      })
    }
  } else if (warp == kEpiWarps2 && rank != 0) {
    // =============================== second CTA: forward "stage landed" to the leader ===============================
    uint32_t cntf = 0;
    constexpr int KSH = (kStageElems / (U_ / 4)) < U_ ? (kStageElems / (U_ / 4)) : U_;
    TNF_FOR_JOBS({
      (void)g;
      const int l = j % 3;
      const int n_st = l == 1 ? 2 * (U_ / KSH) : 1;
      for (int st = 0; st < n_st; ++st) {
        const uint32_t slot = cntf & (kStages5 - 1), par = (cntf >> 2) & 1u;
        mbar_wait(&ct.w_full[slot], par);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(lead_w_peer + slot * 8u);
        ++cntf;
      }
    })
  } else if (warp == kEpiWarps2) {
    // =============================== leader CTA: MMA issuer (warp-uniform, elected lane issues) ===============================
    const bool leader = elect_one();
    uint32_t cntm = 0, e_par = 0, a1_par = 0;   // parities: bit g
    bool started[2] = {false, false};
    const bool diag = a.dbg != nullptr;
    const long long t_all = diag ? clock64() : 0;
    long long t_dep = 0, t_w3[3] = {0, 0, 0};
    int mma_n = 0;
    const uint32_t wfull0 = smem_u32(&ct.w_full[0]), wpeer0 = smem_u32(&ct.w_peer[0]), wempty0 = smem_u32(&ct.w_empty[0]);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint64_t a1_desc = make_desc(smem_u32(sA1), kTileM), act_desc = make_desc(smem_u32(sAct), kTileM);
    const uint32_t a_hi = (uint32_t)(a1_desc >> 32);
    const uint32_t a1_sz16 = (uint32_t)sh.a1_bytes() >> 4, act_sz16 = (uint32_t)sh.act_bytes() >> 4;
    const uint64_t bU_desc = make_desc(0u, U_ / 2), bF_desc = make_desc(0u, DH / 2);   // each CTA holds half of the N rows
    const uint32_t bU_lo = (uint32_t)bU_desc + ring16, bU_hi = (uint32_t)(bU_desc >> 32);
    const uint32_t bF_lo = (uint32_t)bF_desc + ring16, bF_hi = (uint32_t)(bF_desc >> 32);
    const uint64_t bH_desc = make_desc(0u, U_ / 4);      // N-half jobs: each CTA holds U/4 of the half's U/2 weight rows
    const uint32_t bH_lo = (uint32_t)bH_desc + ring16, bH_hi = (uint32_t)(bH_desc >> 32);
    const uint32_t ones_lo = (uint32_t)make_desc(smem_u32(sOnes), kTileM);
    const uint32_t bias16 = smem_u32(sBias) >> 4;       // resident bias blocks: descriptor low word = (addr >> 4) | LBO field
    const uint32_t lboU = (uint32_t)bU_desc, lboF = (uint32_t)bF_desc, lboH = (uint32_t)bH_desc;
    TNF_FOR_JOBS({
      const int net = j / 3, l = j % 3;
      const long long c0 = diag ? clock64() : 0;
      if (started[g]) {   // the group's previous epilogue phase in BOTH CTAs: accumulators drained, activations written
        mbar_wait_addr(smem_u32(&ct.e_done[g]), (e_par >> g) & 1u);
        e_par ^= 1u << g;
      }
      started[g] = true;
      if (j == 0) {
        mbar_wait_addr(smem_u32(&ct.a1_ready[g]), (a1_par >> g) & 1u);
        a1_par ^= 1u << g;
      }
      tc_fence_after();
      if (diag) t_dep += clock64() - c0;
      if (diag && blockIdx.x == 0 && leader && mma_n < 480) {
        a.dbg[3072 + 2 * mma_n] = 1000 + g * 100 + j; a.dbg[3072 + 2 * mma_n + 1] = clock64(); ++mma_n;
      }
      const uint32_t d_tmem = tmem + (uint32_t)g * 256u;
      const uint32_t a1_lo = (uint32_t)a1_desc + (uint32_t)g * a1_sz16;
      const uint32_t act_lo = (uint32_t)act_desc + (uint32_t)g * act_sz16;
      if (l == 0) {
        mma_job5<DH, U_, false, U_>(d_tmem, a1_lo, a_hi, bU_lo, bU_hi, ones_lo, lboU + bias16 + (uint32_t)(sh.bias8_off(0, net, 0) >> 4),
                                    wfull0, wpeer0, wempty0, cntm, leader, diag ? &t_w3[0] : nullptr);
        if (leader) {
          tc_commit2_addr(smem_u32(&ct.h_ready[g]));
          if (net == 1) tc_commit2_addr(smem_u32(&ct.a1_free[g]));
        }
      } else if (l == 1) {   // last hidden job: N-half a -> commit -> N-half b (the group's tanh phase starts on half a)
        mma_job5<U_, U_ / 2, false, U_>(d_tmem, act_lo, a_hi, bH_lo, bH_hi, ones_lo, lboH + bias16 + (uint32_t)(sh.bias8_off(1, net, 0) >> 4),
                                        wfull0, wpeer0, wempty0, cntm, leader, diag ? &t_w3[1] : nullptr);
        if (leader) tc_commit2_addr(smem_u32(&ct.h_ready[g]));
        mma_job5<U_, U_ / 2, false, U_>(d_tmem + (uint32_t)(U_ / 2), act_lo, a_hi, bH_lo, bH_hi, ones_lo,
                                        lboH + bias16 + (uint32_t)(sh.bias8_off(1, net, 1) >> 4), wfull0, wpeer0, wempty0, cntm, leader,
                                        diag ? &t_w3[1] : nullptr);
        if (leader) tc_commit2_addr(smem_u32(&ct.h_ready_b[g]));
      } else {
        mma_job5<U_, DH, true, U_>(d_tmem + fin_col5<U_>(), d_tmem, a_hi, bF_lo, bF_hi, ones_lo,
                                   lboF + bias16 + (uint32_t)(sh.bias8_off(2, net, 0) >> 4), wfull0, wpeer0, wempty0, cntm, leader,
                                   diag ? &t_w3[2] : nullptr);
        if (leader) tc_commit2_addr(smem_u32(&ct.h_ready[g]));
      }
      if (diag && blockIdx.x == 0 && leader && mma_n < 480) {
        a.dbg[3072 + 2 * mma_n] = 2000 + g * 100 + j; a.dbg[3072 + 2 * mma_n + 1] = clock64(); ++mma_n;
      }
      __syncwarp();
    })
    if (diag && blockIdx.x == 0 && leader) {
      a.dbg[2040] = t_dep; a.dbg[2041] = t_w3[0] + t_w3[1] + t_w3[2]; a.dbg[2042] = clock64() - t_all;
      a.dbg[2043] = t_w3[0]; a.dbg[2044] = t_w3[1]; a.dbg[2045] = t_w3[2];
    }
#undef TNF_FOR_JOBS
  } else if (warp < kEpiWarps2) {
    // =============================== epilogue warps ===============================
    const int g = warp >> 3, q = warp & 3, par = (warp >> 2) & 1;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r_tile = q * 32 + lane;
    constexpr int W = DH / 2;                 // final-layer columns per thread (16 or 32)
    const float kLog2e = 1.4426950408889634f;
    const uint32_t hcol = tmem + lane_addr + (uint32_t)g * 256u;
    unsigned char* myAct = sAct + (size_t)g * sh.act_bytes();
    uint32_t h_par = 0, hb_par = 0;
    const float lp_cst = (float)((double)sh.D * 0.91893853320467274178);   // D log sqrt(2 pi)
    const float lp_scal0 = (lp_mode && a.lp_scal) ? a.lp_scal[0] : 0.f;

    int dbg_n = 0;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && q == 0 && par == 0 && lane == 0;
    long long* dbg = a.dbg + g * 1024;
#define TNF_STAMP(tag)                                                                     \
  do {                                                                                     \
    if (dbg_on && dbg_n < 500) { dbg[2 * dbg_n] = (tag); dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; } \
  } while (0)

    // accumulator chunk c (bias already added by the bias MMA) -> MUFU.TANH -> bf16 -> A image.  No software pipelining
    // inside the warp (a double-buffered variant measured no faster): the other warps of the sub-partition cover it.
    auto epi_step = [&](int c) {
      uint32_t x[32];
      tmem_ld32(hcol + (uint32_t)(c * kChunk), x);
      tc_wait_ld();
      unsigned char* dst = myAct + img_off(r_tile, c * kChunk, kTileM);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          x[j + e] = __float_as_uint(e < FT ? tanh_poly(__uint_as_float(x[j + e])) : tanh_fast(__uint_as_float(x[j + e])));
        *reinterpret_cast<uint4*>(dst + (j >> 3) * (kTileM * 16)) =
            make_uint4(pack_bf16(__uint_as_float(x[j]), __uint_as_float(x[j + 1])),
                       pack_bf16(__uint_as_float(x[j + 2]), __uint_as_float(x[j + 3])),
                       pack_bf16(__uint_as_float(x[j + 4]), __uint_as_float(x[j + 5])),
                       pack_bf16(__uint_as_float(x[j + 6]), __uint_as_float(x[j + 7])));
      }
    };
    // last tanh phase: the bf16 activations go back into accumulator columns this warp has already consumed
    auto epi_step_tm = [&](int c) {
      uint32_t x[32];
      tmem_ld32(hcol + (uint32_t)(c * kChunk), x);
      tc_wait_ld();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          x[j + e] = __float_as_uint(e < FT ? tanh_poly(__uint_as_float(x[j + e])) : tanh_fast(__uint_as_float(x[j + e])));
#pragma unroll
        for (int e = 0; e < 8; e += 2) pk[(j + e) >> 1] = pack_bf16(__uint_as_float(x[j + e]), __uint_as_float(x[j + e + 1]));
      }
      tmem_st16(hcol + act_col5<U_>(2 * c), pk);
    };
    // end of an epilogue phase: accumulator reads done, activation image visible to the async proxy
    auto phase_done = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_e_done + (uint32_t)g * 8u);
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + q) : "memory"); };

    for (int64_t k = 0; k < cnt[g]; ++k) {
      const int64_t tile = 2 * ((2 * k + g) * P + pair) + rank;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = row < a.rows;
      const float* zrow = a.z_in + row * sh.D + sh.t_off + par * W;
      TNF_STAMP(100);
      float tv[W];
#pragma unroll
      for (int net = 0; net < 2; ++net) {
#pragma unroll 1
        for (int l = 0; l < sh.L - 1; ++l) {
          TNF_STAMP(200 + net * 10 + l);
          mbar_wait(&ct.h_ready[g], h_par);
          h_par ^= 1;
          tc_fence_after();
          TNF_STAMP(300 + net * 10 + l);
#pragma unroll 1
          for (int c = par; c < n_chunks; c += 2) epi_step(c);
          phase_done();
        }
        {   // last tanh phase, on the split job: N-half a while the tensor pipe still computes N-half b
          TNF_STAMP(200 + net * 10 + sh.L - 1);
          mbar_wait(&ct.h_ready[g], h_par);
          h_par ^= 1;
          tc_fence_after();
          TNF_STAMP(300 + net * 10 + sh.L - 1);
#pragma unroll 1
          for (int c = par; c < n_chunks / 2; c += 2) epi_step_tm(c);
          TNF_STAMP(250 + net * 10);
          mbar_wait(&ct.h_ready_b[g], hb_par);
          hb_par ^= 1;
          tc_fence_after();
          TNF_STAMP(350 + net * 10);
#pragma unroll 1
          for (int c = n_chunks / 2 + par; c < n_chunks; c += 2) epi_step_tm(c);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(lead_e_done + (uint32_t)g * 8u);
        }
        // ---- final layer of this net: W columns per thread
        float zin[W];
        float ld_old = 0.f;
        if (net == 1) {   // the transformed half and the old log-det arrive while the final-layer MMAs run
#pragma unroll
          for (int j = 0; j < W; j += 8) {
            if (valid) {
              ldg256_nc(zrow + j, zin + j);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) zin[j + e] = 0.f;
            }
          }
          if (par == 0 && valid && a.accum != TNF_LD_WRITE) ld_old = a.log_det[row];
        }
        TNF_STAMP(400 + net);
        mbar_wait(&ct.h_ready[g], h_par);
        h_par ^= 1;
        tc_fence_after();
        TNF_STAMP(500 + net);
        uint32_t o[W];
        if (W == 16) tmem_ld16(hcol + fin_col5<U_>() + (uint32_t)(par * W), reinterpret_cast<uint32_t(&)[16]>(o));
        else tmem_ld32(hcol + fin_col5<U_>() + (uint32_t)(par * W), reinterpret_cast<uint32_t(&)[32]>(o));
        tc_wait_ld();
        if (net == 0) {
#pragma unroll
          for (int j = 0; j < W; ++j) tv[j] = __uint_as_float(o[j]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(lead_e_done + (uint32_t)g * 8u);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(lead_e_done + (uint32_t)g * 8u);   // accumulator free: the next tile's first job may start
          float ld_sum = 0.f;
          float (&y)[W] = zin;   // transformed in place
#pragma unroll
          for (int j = 0; j < W; ++j) {
            const int col = sh.t_off + par * W + j;
            const float zz = fmaf(zin[j], s_pscale[col], s_pshift[col]);
            const float sv = __uint_as_float(o[j]);
            ld_sum += sv;
            y[j] = kInverse ? (zz - tv[j]) * exp2_fast(-sv * kLog2e) : fmaf(zz, exp2_fast(sv * kLog2e), tv[j]);
          }
          TNF_STAMP(601);
          if (lp_mode) {   // log N(z_out) - log-dets instead of z_out: nothing reads the base sample itself
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < W; ++j) ss = fmaf(y[j], y[j], ss);
            if (par == 1) s_ldp[g * kTileM + r_tile] = fmaf(0.5f, ss, ld_sum);
            pair_sync();
            if (par == 0 && valid) {
              const float ss_all = ss + s_ss[(g * 2 + (int)(k & 1)) * kTileM + r_tile];
              a.out_lp[row] = ((-0.5f * ss_all - lp_cst) - ((ld_old + ld_sum) + s_ldp[g * kTileM + r_tile])) - lp_scal0;
            }
          } else {
          if (valid) {
            float* orow = a.z_out + row * sh.D + sh.t_off + par * W;
#pragma unroll
            for (int j = 0; j < W; j += 8) stg256(orow + j, y + j);
          }
          if (want_stats) {   // release the stored tile half to the I/O warps (statistics)
            __syncwarp();
            if (lane == 0) mbar_arrive(&ct.y_done[g]);
          }
          if (par == 1) s_ldp[g * kTileM + r_tile] = ld_sum;
          pair_sync();
          if (par == 0 && valid) {
            const float tot = ld_sum + s_ldp[g * kTileM + r_tile];
            float* op = a.log_det + row;
            if (a.accum == TNF_LD_WRITE) *op = tot;
            else if (a.accum == TNF_LD_ADD) *op = ld_old + tot;
            else *op = ld_old - tot;
          }
          }
        }
      }
      TNF_STAMP(600);
    }
#undef TNF_STAMP
  } else {
    // =============================== I/O warps: conditioning half, coalesced ===============================
    const int w2 = warp - (kEpiWarps2 + 2);   // warps 18, 19
    const int row0 = w2 * (kTileM / 2);
    constexpr int LPR = DH / 4;          // lanes per row of the conditioning half (16 B each)
    constexpr int RPI = 32 / LPR;        // rows per warp instruction
    constexpr int NI = (kTileM / 2) / RPI;
    const int hc = 4 * (lane % LPR);     // column inside the half
    const int col = sh.c_off + hc;
    const int rsub = lane / LPR;
    const float4 ps = *reinterpret_cast<const float4*>(s_pscale + col);
    const float4 pb = *reinterpret_cast<const float4*>(s_pshift + col);
    auto load_tile = [&](int g, int64_t k) {
      const int64_t tile = 2 * ((2 * k + g) * P + pair) + rank;
      unsigned char* a1 = sA1 + (size_t)g * sh.a1_bytes();
#pragma unroll 1
      for (int n0 = 0; n0 < NI; n0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = row0 + (n0 + u) * RPI + rsub;
          const int64_t grow = tile * kTileM + r;
          v[u] = grow < a.rows ? __ldg(reinterpret_cast<const float4*>(a.z_in + grow * sh.D + col))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int r = row0 + (n0 + u) * RPI + rsub;
          const int64_t grow = tile * kTileM + r;
          float4 x;
          x.x = fmaf(v[u].x, ps.x, pb.x); x.y = fmaf(v[u].y, ps.y, pb.y);
          x.z = fmaf(v[u].z, ps.z, pb.z); x.w = fmaf(v[u].w, ps.w, pb.w);
          *reinterpret_cast<uint2*>(a1 + img_off(r, hc, kTileM)) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
          if (lp_mode) {   // sum of squares of this row's conditioning half (it passes through unchanged), for the epilogue
            float ssq = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, x.w * x.w)));
#pragma unroll
            for (int o = 1; o < LPR; o <<= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
            if (lane % LPR == 0) s_ss[(g * 2 + (int)(k & 1)) * kTileM + r] = ssq;
          } else if (grow < a.rows) {
            *reinterpret_cast<float4*>(a.z_out + grow * sh.D + col) = x;
            if (want_stats) {
              io_sv1[0] += x.x; io_sv1[1] += x.y; io_sv1[2] += x.z; io_sv1[3] += x.w;
              io_sv2[0] = fmaf(x.x, x.x, io_sv2[0]); io_sv2[1] = fmaf(x.y, x.y, io_sv2[1]);
              io_sv2[2] = fmaf(x.z, x.z, io_sv2[2]); io_sv2[3] = fmaf(x.w, x.w, io_sv2[3]);
            }
          }
        }
      }
      fence_async_smem();
      if (lp_mode) __threadfence_block();   // s_ss before the (relaxed) arrival the epilogue's h_ready chain descends from
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(lead_a1_ready + (uint32_t)g * 8u);
    };
    // column sums of the transformed half of the group's tile k, read back (L2) after its epilogue warps stored it
    auto stats_tile = [&](int g, int64_t k) {
      mbar_wait(&ct.y_done[g], (uint32_t)(k & 1));
      const int64_t tile = 2 * ((2 * k + g) * P + pair) + rank;
      const int tcol = sh.t_off + hc;
#pragma unroll 1
      for (int n0 = 0; n0 < NI; n0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t grow = tile * kTileM + row0 + (n0 + u) * RPI + rsub;
          v[u] = grow < a.rows ? __ldcg(reinterpret_cast<const float4*>(a.z_out + grow * sh.D + tcol))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          io_tv1[0] += v[u].x; io_tv1[1] += v[u].y; io_tv1[2] += v[u].z; io_tv1[3] += v[u].w;
          io_tv2[0] = fmaf(v[u].x, v[u].x, io_tv2[0]); io_tv2[1] = fmaf(v[u].y, v[u].y, io_tv2[1]);
          io_tv2[2] = fmaf(v[u].z, v[u].z, io_tv2[2]); io_tv2[3] = fmaf(v[u].w, v[u].w, io_tv2[3]);
        }
      }
    };
    for (int g = 0; g < 2; ++g)
      if (cnt[g] > 0) load_tile(g, 0);
    for (int64_t k = 0; k < cnt[0]; ++k) {
      for (int g = 0; g < 2; ++g) {
        if (k + 1 < cnt[g]) {
          mbar_wait(&ct.a1_free[g], (uint32_t)(k & 1));   // layer-0 jobs of the group's tile k are done with the image
          load_tile(g, k + 1);
        }
      }
      if (want_stats)
        for (int g = 0; g < 2; ++g)
          if (k < cnt[g]) stats_tile(g, k);
    }
    if (want_stats) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
          io_sv1[e] += __shfl_xor_sync(0xffffffffu, io_sv1[e], o);
          io_sv2[e] += __shfl_xor_sync(0xffffffffu, io_sv2[e], o);
          io_tv1[e] += __shfl_xor_sync(0xffffffffu, io_tv1[e], o);
          io_tv2[e] += __shfl_xor_sync(0xffffffffu, io_tv2[e], o);
        }
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    // gather: one [2*D] row of doubles per CTA = fixed-order sum over the epilogue warps' rows and the I/O warps' sums
    if (warp == kEpiWarps2 + 2 || warp == kEpiWarps2 + 3) {   // I/O warps: expand their sums into rows w2
      const int w2 = warp - (kEpiWarps2 + 2);
      constexpr int LPR = DH / 4;
      double* rowp = stat_rows + (size_t)w2 * 2 * sh.D;
      if (lane < LPR) {   // the two halves cover all D columns
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          rowp[sh.c_off + 4 * lane + e] = (double)io_sv1[e];
          rowp[sh.D + sh.c_off + 4 * lane + e] = (double)io_sv2[e];
          rowp[sh.t_off + 4 * lane + e] = (double)io_tv1[e];
          rowp[sh.D + sh.t_off + 4 * lane + e] = (double)io_tv2[e];
        }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * sh.D; i += blockDim.x) {
      double acc = 0.0;
      for (int w = 0; w < kStatWarps; ++w) acc += stat_rows[(size_t)w * 2 * sh.D + i];
      a.stat_partials[(size_t)blockIdx.x * 2 * sh.D + i] = acc;
    }
  }
  cluster_sync_all();   // no CTA leaves (or frees TMEM) while its partner can still signal it
  if (warp == kEpiWarps2) tmem_dealloc2(tmem, 512);
}

constexpr int kDefaultFT5 = 0;

int launch_tc5(const Args& a, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  if (!shape_supported5(a.D, a.U, a.L)) return (int)cudaErrorInvalidValue;
#define TNF_TC5_LAUNCH(INV, DHV, UV, FTV)                                                                         \
  do {                                                                                                            \
    e = cudaFuncSetAttribute(coupling_tc5_kernel<INV, DHV, UV, FTV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             (int)smem);                                                                          \
    if (e == cudaSuccess) coupling_tc5_kernel<INV, DHV, UV, FTV><<<grid, kThreads5, smem, st>>>(a);               \
  } while (0)
#define TNF_TC5_U(INV, DHV)                                               \
  do {                                                                    \
    if (a.U == 256) TNF_TC5_LAUNCH(INV, DHV, 256, kDefaultFT5);           \
    else TNF_TC5_LAUNCH(INV, DHV, 128, kDefaultFT5);                      \
  } while (0)
  // FMA-pipe tanh share: 1 of 8 activations at the C3 shape (measured best: 0 -> 0.416, 1 -> 0.400, 2+ slower);
  // tune bit 8 (variant 3) turns it off = the arithmetic of coupling_tc4_kernel bit for bit
  const bool ft_off = (a.tune & 0x100) != 0;
  if (a.D == 64 && a.U == 256 && !ft_off) {
    if (a.inverse) TNF_TC5_LAUNCH(true, 32, 256, 1); else TNF_TC5_LAUNCH(false, 32, 256, 1);
  } else if (a.D == 64) { if (a.inverse) TNF_TC5_U(true, 32); else TNF_TC5_U(false, 32); }
  else { if (a.inverse) TNF_TC5_U(true, 64); else TNF_TC5_U(false, 64); }
#undef TNF_TC5_U
#undef TNF_TC5_LAUNCH
  return (int)e;
}

}  // namespace tc
}  // namespace tnf
