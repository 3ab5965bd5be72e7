// RealNVP coupling layer fused on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-sample tiles.  For each tile and each of
// the two conditioner nets (shift t, scale s) the whole MLP runs on-chip:
//
//   z1 (fp32, HBM) --cvt--> A1 bf16 in TMEM
//   H  = A_l . W_l          tcgen05.mma kind::f16, A from TMEM, B from SMEM, D fp32 in TMEM
//   A_{l+1} = bf16(tanh(H + b_l))   epilogue warps: tcgen05.ld -> MUFU.TANH -> pack -> tcgen05.st
//   ...
//   (t, s) = A_L . W_L + b_L ;  z2' = t + z2 e^s  |  (z2 - t)/e^s ;  log_det += sum s
//
// Activations never leave the SM: the only HBM traffic is z in, z out and the
// log-det read-modify-write (520 B per sample-layer at D = 64).  Weights are
// pre-packed (tnf_tc_pack) into the exact shared-memory images the UMMA
// descriptors expect (K-major, no swizzle, 8x16-byte core matrices) and are
// streamed from L2 through a ring of 8 KB stages by one producer thread with
// cp.async.bulk + mbarrier complete_tx.
//
// TMEM plan (512 columns x 128 lanes, lane = sample row of the tile):
//   [  0,128) H_lo   fp32 accumulator, hidden units   0..127 (also the final t/s output)
//   [128,256) H_hi   fp32 accumulator, hidden units 128..255
//   [256,384) R0     bf16 A operand (two values per column)  layers 0, 2, 4
//   [384,512) R1     bf16 A operand                          layers 1, 3, 5
// Hidden layers are issued as two N=128 halves so the epilogue of one half
// overlaps the MMAs of the other; the A operand of the next layer is published
// in 32-column chunks (mbarrier per chunk) so its MMAs start while the
// epilogue is still running.
//
// Warp roles (320 threads): warps 0-7 epilogue (warp w owns TMEM lane quadrant
// w%4 and the 32-column chunks c with c%2 == w/4), warp 8 lane 0 MMA issuer,
// warp 9 lane 0 weight producer.
//
// Reference semantics: torch_nf/bijectors.py:145-242 (RealNVP).
#include <cuda_bf16.h>

#include "common.cuh"

namespace tnf {
namespace tc {

constexpr int kTileM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kStageBytes = 8192;  // K=32 x N=128 bf16
constexpr int kStages = 20;
constexpr int kChunk = 32;         // hidden columns per published A chunk
constexpr int kMaxChunks = 8;      // U <= 256
constexpr uint32_t kColHlo = 0, kColHhi = 128, kColR0 = 256, kColR1 = 384;

struct Shape {
  int D, U, L, upper;
  int d_in, d_out, c_off, t_off;
  int Nh, nh, chunks, cph;  // hidden half width, halves, chunks per layer, chunks per half
  __host__ __device__ Shape(int D_, int U_, int L_, int upper_) : D(D_), U(U_), L(L_), upper(upper_) {
    int h = D / 2;
    d_in = h; d_out = h;
    c_off = upper ? 0 : h;
    t_off = upper ? h : 0;
    Nh = U < 128 ? U : 128;
    nh = U / Nh;
    chunks = U / kChunk;
    cph = Nh / kChunk;
  }
  // elements of one net's weights
  __host__ __device__ int64_t net_weight_elems() const {
    return (int64_t)d_in * U + (int64_t)(L - 1) * U * U + (int64_t)U * d_out;
  }
  __host__ __device__ int net_bias_elems() const { return L * U + d_out; }
  __host__ __device__ int64_t packed_bytes() const { return 2 * net_weight_elems() * 2 + 2 * (int64_t)net_bias_elems() * 4; }
  __host__ __device__ int K_of(int l) const { return l == 0 ? d_in : U; }
  __host__ __device__ int J_of(int l) const { return l == L ? d_out : U; }
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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

// D[tmem] (+)= A[tmem] . B[smem desc]   (M=128, K=16, bf16 -> fp32)
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
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE:
//   core matrix = 8 rows x 16 bytes stored contiguously (128 B);
//   SBO = distance between 8-row groups (128 B: groups are adjacent),
//   LBO = distance between the two 8-element K groups of one K=16 step (= N*16 B).
__device__ __forceinline__ uint64_t make_bdesc(uint32_t smem_addr, int N) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(((uint32_t)N * 16u >> 4) & 0x3FFF) << 16;  // LBO
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;                // SBO
  d |= (uint64_t)1 << 46;                                     // descriptor version (sm_100)
  return d;
}
// byte offset of element (n, k) inside one packed stage holding Kc x N
__host__ __device__ inline uint32_t stage_elem_off(int n, int k, int N) {
  return (uint32_t)((k >> 3) * N * 16 + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2);
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// low half = first (lower K index) element
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- weight packing
// Packed buffer: for net in {t, s}: for layer l in 0..L: for N-half hb: for K-chunk kc (32 rows of K):
//   one stage image of Kc x N bf16 in the UMMA no-swizzle K-major layout (see stage_elem_off),
// followed by the fp32 biases [net][layer][unit].
__global__ void pack_kernel(const float* __restrict__ params, unsigned char* __restrict__ packed, Shape sh) {
  const int64_t per_net = sh.net_weight_elems();
  const int64_t total = 2 * per_net;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // idx enumerates SOURCE elements: net, layer, k, j
    const int net = idx >= per_net;
    int64_t rem = idx - net * per_net;
    int l = 0;
    int64_t src_off = 0;   // offset of layer l inside the reference parameter row
    int64_t dst_layer = 0; // element offset of layer l inside this net's packed weights
    for (;; ++l) {
      const int64_t n_el = (int64_t)sh.K_of(l) * sh.J_of(l);
      if (rem < n_el) break;
      rem -= n_el;
      src_off += 2 * n_el + 2 * sh.J_of(l);
      dst_layer += n_el;
    }
    const int K = sh.K_of(l), J = sh.J_of(l);
    const int k = (int)(rem / J), j = (int)(rem % J);
    const float w = params[src_off + (net ? (int64_t)K * J : 0) + rem];
    // destination: half hb (width N), K-chunk kc
    const int N = (l == sh.L) ? J : sh.Nh;
    const int hb = j / N, n = j % N;
    const int kc = k / 32, kk = k % 32;
    const int kcs = (K + 31) / 32;                       // K-chunks per half
    // bytes of the stages before (hb, kc): all full chunks hold 32 x N elements; only the last can be shorter
    int64_t before = ((int64_t)hb * K + (int64_t)kc * 32) * N;  // elements
    (void)kcs;
    unsigned char* dst = packed + ((int64_t)net * per_net + dst_layer + before) * 2 + stage_elem_off(n, kk, N);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(w);
  }
  // biases
  const int nb = sh.net_bias_elems();
  float* bias_dst = reinterpret_cast<float*>(packed + 2 * per_net * 2);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * nb; idx += stride) {
    const int net = idx >= nb;
    int rem = (int)(idx - net * nb);
    int l = 0;
    int64_t src_off = 0;
    for (;; ++l) {
      if (rem < sh.J_of(l)) break;
      rem -= sh.J_of(l);
      src_off += 2 * (int64_t)sh.K_of(l) * sh.J_of(l) + 2 * sh.J_of(l);
    }
    const int K = sh.K_of(l), J = sh.J_of(l);
    bias_dst[idx] = params[src_off + 2 * (int64_t)K * J + (net ? J : 0) + rem];
  }
}

// ---------------------------------------------------------------- the fused kernel
struct Args {
  const float* z_in; float* z_out; float* log_det; const unsigned char* packed;
  const float* pre_scale; const float* pre_shift;
  int64_t rows;
  int D, U, L, upper, inverse, accum;
};

struct __align__(16) Smem {
  unsigned char ring[kStages][kStageBytes];
  uint64_t w_full[kStages];
  uint64_t w_empty[kStages];
  uint64_t a1_ready;
  uint64_t a_ready[kMaxChunks];
  uint64_t h_ready[2];
  uint32_t tmem_base;
  float ld_xchg[kTileM];
  // followed by: bias[2][L*U + d_out] floats, pre_scale[D], pre_shift[D]
};

template <bool kInverse, int DH>   // DH = D/2 = d_in = d_out
__global__ void __launch_bounds__(kThreads, 1) coupling_tc_kernel(Args a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const Shape sh(a.D, a.U, a.L, a.upper);
  float* s_bias = reinterpret_cast<float*>(smem_raw + sizeof(Smem));
  const int nb = sh.net_bias_elems();
  float* s_pscale = s_bias + 2 * nb;
  float* s_pshift = s_pscale + sh.D;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const int64_t per_net_bytes = sh.net_weight_elems() * 2;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    mbar_init(&sm.a1_ready, kEpiWarps);
    for (int i = 0; i < kMaxChunks; ++i) mbar_init(&sm.a_ready[i], 4);
    mbar_init(&sm.h_ready[0], 1);
    mbar_init(&sm.h_ready[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) tmem_alloc(&sm.tmem_base, 512);
  {
    const float* gb = reinterpret_cast<const float*>(a.packed + 2 * per_net_bytes);
    for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) s_bias[i] = gb[i];
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == kEpiWarps + 1) {
    // =============================== weight producer ===============================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int net = 0; net < 2; ++net) {
          const unsigned char* src = a.packed + (int64_t)net * per_net_bytes;
          for (int l = 0; l <= sh.L; ++l) {
            const int K = sh.K_of(l);
            const int N = (l == sh.L) ? sh.d_out : sh.Nh;
            const int halves = (l == sh.L) ? 1 : sh.nh;
            for (int hb = 0; hb < halves; ++hb) {
              for (int k0 = 0; k0 < K; k0 += 32) {
                const int kc = (K - k0) < 32 ? (K - k0) : 32;
                const uint32_t bytes = (uint32_t)(kc * N * 2);
                mbar_wait(&sm.w_empty[slot], phase ^ 1);
                mbar_arrive_expect_tx(&sm.w_full[slot], bytes);
                bulk_g2s(sm.ring[slot], src, bytes, &sm.w_full[slot]);
                src += bytes;
                if (++slot == kStages) { slot = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      uint32_t slot = 0, phase = 0;
      uint32_t a1_phase = 0;
      uint32_t a_phase = 0;  // bit c = parity to wait for on a_ready[c]
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int net = 0; net < 2; ++net) {
          for (int l = 0; l <= sh.L; ++l) {
            const int K = sh.K_of(l);
            const int N = (l == sh.L) ? sh.d_out : sh.Nh;
            const int halves = (l == sh.L) ? 1 : sh.nh;
            const uint32_t idesc = make_idesc(N);
            const uint32_t a_col = (l & 1) ? kColR1 : kColR0;
            uint32_t waited = 0;  // chunks of A_l already waited for in this layer
            if (l == 0) {
              mbar_wait(&sm.a1_ready, a1_phase);
              a1_phase ^= 1;
              tc_fence_after();
            }
            for (int hb = 0; hb < halves; ++hb) {
              const uint32_t d_col = hb ? kColHhi : kColHlo;
              if (l > 0) {
                // the accumulator region must have been drained by the epilogue of layer l-1:
                // H_lo holds chunks [0, cph), H_hi chunks [cph, 2 cph)
                for (int c = hb * sh.cph; c < (hb + 1) * sh.cph && c < sh.chunks; ++c) {
                  if (!(waited >> c & 1)) {
                    mbar_wait(&sm.a_ready[c], (a_phase >> c) & 1);
                    a_phase ^= 1u << c;
                    waited |= 1u << c;
                  }
                }
                tc_fence_after();
              }
              for (int k0 = 0; k0 < K; k0 += 32) {
                const int kc = (K - k0) < 32 ? (K - k0) : 32;
                if (l > 0) {
                  const int c = k0 / kChunk;
                  if (!(waited >> c & 1)) {
                    mbar_wait(&sm.a_ready[c], (a_phase >> c) & 1);
                    a_phase ^= 1u << c;
                    waited |= 1u << c;
                    tc_fence_after();
                  }
                }
                mbar_wait(&sm.w_full[slot], phase);
                tc_fence_after();
                const uint32_t b_addr = smem_u32(sm.ring[slot]);
                for (int ks = 0; ks < kc; ks += 16) {
                  const uint64_t bdesc = make_bdesc(b_addr + (uint32_t)(ks >> 3) * (uint32_t)N * 16u, N);
                  umma_ts(tmem + d_col, tmem + a_col + (uint32_t)((k0 + ks) >> 1), bdesc, idesc,
                          (k0 + ks) > 0 ? 1u : 0u);
                }
                tc_commit(&sm.w_empty[slot]);
                if (++slot == kStages) { slot = 0; phase ^= 1; }
              }
              tc_commit(&sm.h_ready[hb]);
            }
          }
        }
      }
    }
  } else {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    uint32_t h_phase = 0;  // bit hb = parity to wait for on h_ready[hb]
    constexpr int nc_in = DH / 2;    // conditioning columns handled by this thread (its half)
    constexpr int nc_out = DH / 2;   // transformed columns handled by this thread
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * kTileM + q * 32 + lane;
      const bool valid = row < a.rows;
      const float* zrow = a.z_in + row * sh.D;
      float* orow = a.z_out + row * sh.D;
      // ---- conditioning half: load, pre-affine, pass through, pack to bf16
      uint32_t a1[nc_in / 2];  // packed bf16 pairs
      {
        const int c0 = sh.c_off + half * nc_in;
#pragma unroll
        for (int j = 0; j < nc_in; j += 4) {
          float4 v = valid ? *reinterpret_cast<const float4*>(zrow + c0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          v.x = fmaf(v.x, s_pscale[c0 + j + 0], s_pshift[c0 + j + 0]);
          v.y = fmaf(v.y, s_pscale[c0 + j + 1], s_pshift[c0 + j + 1]);
          v.z = fmaf(v.z, s_pscale[c0 + j + 2], s_pshift[c0 + j + 2]);
          v.w = fmaf(v.w, s_pscale[c0 + j + 3], s_pshift[c0 + j + 3]);
          if (valid) *reinterpret_cast<float4*>(orow + c0 + j) = v;
          a1[j / 2] = pack_bf16(v.x, v.y);
          a1[j / 2 + 1] = pack_bf16(v.z, v.w);
        }
      }
      float tv[nc_out];  // shift outputs of the t-net for this thread's columns
      float ld_part = 0.f;
      for (int net = 0; net < 2; ++net) {
        const float* bias = s_bias + net * nb;
        // ---- publish A1 (bf16) into R0: this half's columns
        {
          const uint32_t dst = tmem + lane_addr + kColR0 + (uint32_t)(half * nc_in / 2);
#pragma unroll
          for (int j = 0; j < nc_in / 2; j += 8) tmem_st8(dst + j, &a1[j]);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.a1_ready);
        }
        // ---- hidden layers: H -> tanh -> next A operand
        for (int l = 0; l < sh.L; ++l) {
          const uint32_t dst_col = ((l + 1) & 1) ? kColR1 : kColR0;
          const float* bl = bias + l * sh.U;
          for (int c = half; c < sh.chunks; c += 2) {
            const int hb = c / sh.cph;
            if (c - hb * sh.cph < 2) {  // first chunk this warp touches in this half
              mbar_wait(&sm.h_ready[hb], (h_phase >> hb) & 1);
              h_phase ^= 1u << hb;
              tc_fence_after();
            }
            uint32_t acc[32];
            tmem_ld32(tmem + lane_addr + (hb ? kColHhi : kColHlo) + (uint32_t)((c - hb * sh.cph) * kChunk), acc);
            tc_wait_ld();
            uint32_t packed[16];
            const float4* b4 = reinterpret_cast<const float4*>(bl + c * kChunk);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = b4[j / 4];
              const float x0 = tanh_fast(__uint_as_float(acc[j]) + b.x);
              const float x1 = tanh_fast(__uint_as_float(acc[j + 1]) + b.y);
              const float x2 = tanh_fast(__uint_as_float(acc[j + 2]) + b.z);
              const float x3 = tanh_fast(__uint_as_float(acc[j + 3]) + b.w);
              packed[j / 2] = pack_bf16(x0, x1);
              packed[j / 2 + 1] = pack_bf16(x2, x3);
            }
            tmem_st16(tmem + lane_addr + dst_col + (uint32_t)(c * (kChunk / 2)), packed);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.a_ready[c]);
          }
        }
        // ---- final layer output (t or s) for this thread's columns
        mbar_wait(&sm.h_ready[0], h_phase & 1);
        h_phase ^= 1u;
        tc_fence_after();
        const float* bL = bias + sh.L * sh.U + half * nc_out;
        const int z0 = sh.t_off + half * nc_out;
#pragma unroll
        for (int j0 = 0; j0 < nc_out; j0 += 16) {
          uint32_t o[16];
          tmem_ld16(tmem + lane_addr + kColHlo + (uint32_t)(half * nc_out + j0), o);
          tc_wait_ld();
          if (net == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) tv[j0 + j] = __uint_as_float(o[j]) + bL[j0 + j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 v = valid ? *reinterpret_cast<const float4*>(zrow + z0 + j0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              float zz[4] = {v.x, v.y, v.z, v.w};
              float yy[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int col = z0 + j0 + j + e;
                const float zin = fmaf(zz[e], s_pscale[col], s_pshift[col]);
                const float s = __uint_as_float(o[j + e]) + bL[j0 + j + e];
                const float t = tv[j0 + j + e];
                ld_part += s;
                yy[e] = kInverse ? __fdiv_rn(zin - t, expf(s)) : fmaf(zin, expf(s), t);
              }
              if (valid) *reinterpret_cast<float4*>(orow + z0 + j0 + j) = make_float4(yy[0], yy[1], yy[2], yy[3]);
            }
          }
        }
        tc_fence_before();
      }
      // ---- log-det: combine the two column halves of a row, one writer per row
      if (half == 1) sm.ld_xchg[q * 32 + lane] = ld_part;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      if (half == 0 && valid) {
        const float ld = ld_part + sm.ld_xchg[q * 32 + lane];
        float* o = a.log_det + row;
        if (a.accum == TNF_LD_WRITE) *o = ld;
        else if (a.accum == TNF_LD_ADD) *o += ld;
        else *o -= ld;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------- diagnostic: one UMMA GEMM
// out[128 x N] = bf16(A[128 x K]) . bf16(W[K x N]), through the same packing, descriptors and TMEM
// layouts as the fused kernel (A in TMEM via tcgen05.st, B image in SMEM, D read with tcgen05.ld).
__global__ void __launch_bounds__(128, 1) selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                           float* __restrict__ out, int K, int N) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* bimg = smem_raw;                                  // K/32 stages of 32 x N
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)K * N * 2);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tbase, 512);
  for (int idx = threadIdx.x; idx < K * N; idx += blockDim.x) {
    const int k = idx / N, n = idx % N;
    unsigned char* dst = bimg + (size_t)(k / 32) * 32 * N * 2 + stage_elem_off(n, k % 32, N);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(W[idx]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tbase;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + lane;
  for (int k = 0; k < K; k += 16) {
    uint32_t p[8];
    for (int j = 0; j < 8; ++j) p[j] = pack_bf16(A[row * K + k + 2 * j], A[row * K + k + 2 * j + 1]);
    tmem_st8(tmem + lane_addr + kColR0 + (uint32_t)(k / 2), p);
  }
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(N);
    for (int k0 = 0; k0 < K; k0 += 32) {
      const uint32_t b_addr = smem_u32(bimg + (size_t)(k0 / 32) * 32 * N * 2);
      for (int ks = 0; ks < 32 && k0 + ks < K; ks += 16) {
        const uint64_t bdesc = make_bdesc(b_addr + (uint32_t)(ks >> 3) * (uint32_t)N * 16u, N);
        umma_ts(tmem + kColHlo, tmem + kColR0 + (uint32_t)((k0 + ks) >> 1), bdesc, idesc, (k0 + ks) > 0 ? 1u : 0u);
      }
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t o[16];
    tmem_ld16(tmem + lane_addr + kColHlo + (uint32_t)n0, o);
    tc_wait_ld();
    for (int j = 0; j < 16; ++j) out[row * N + n0 + j] = __uint_as_float(o[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace tc
}  // namespace tnf

using namespace tnf;

extern "C" {

int tnf_tc_supported(int D, int U, int L) { return tc::shape_supported(D, U, L) ? 1 : 0; }

size_t tnf_tc_packed_bytes(int D, int U, int L) {
  if (!tc::shape_supported(D, U, L)) return 0;
  return (size_t)tc::Shape(D, U, L, 1).packed_bytes();
}

int tnf_tc_pack(const float* params, void* packed, int D, int U, int L, int transform_upper, tnf_stream_t stream) {
  TNF_REQUIRE(tc::shape_supported(D, U, L), TNF_ERR_UNSUPPORTED, "tnf_tc_pack: shape D=%d U=%d L=%d not supported", D,
              U, L);
  TNF_REQUIRE(params && packed, TNF_ERR_ARG, "tnf_tc_pack: null pointer");
  TNF_REQUIRE(((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN, "tnf_tc_pack: packed buffer must be 16-byte aligned");
  tc::Shape sh(D, U, L, transform_upper != 0);
  tc::pack_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(params, (unsigned char*)packed, sh);
  return check_launch("tnf_tc_pack");
}

int tnf_coupling_tc(const float* z_in, float* z_out, float* log_det, const void* packed, int64_t rows, int D, int U,
                    int L, int transform_upper, int direction, int accum, const float* pre_scale,
                    const float* pre_shift, double* col_stats, tnf_stream_t stream) {
  TNF_REQUIRE(tc::shape_supported(D, U, L), TNF_ERR_UNSUPPORTED,
              "tnf_coupling_tc: shape D=%d U=%d L=%d not supported", D, U, L);
  TNF_REQUIRE(rows >= 0, TNF_ERR_ARG, "tnf_coupling_tc: rows < 0");
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && z_out && log_det && packed, TNF_ERR_ARG, "tnf_coupling_tc: null pointer");
  TNF_REQUIRE((((uintptr_t)z_in | (uintptr_t)z_out | (uintptr_t)packed) & 15) == 0, TNF_ERR_ALIGN,
              "tnf_coupling_tc: z and packed weights must be 16-byte aligned");
  TNF_REQUIRE(col_stats == nullptr, TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: fused column statistics not built yet");
  tc::Shape sh(D, U, L, transform_upper != 0);
  tc::Args a{z_in, z_out, log_det, (const unsigned char*)packed, pre_scale, pre_shift, rows,
             D, U, L, transform_upper != 0, direction == TNF_INVERSE, accum};
  const size_t smem = sizeof(tc::Smem) + (size_t)(2 * sh.net_bias_elems() + 2 * D) * sizeof(float);
  const int64_t n_tiles = (rows + tc::kTileM - 1) / tc::kTileM;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e;
#define TNF_TC_LAUNCH(INV, DHV)                                                                               \
  do {                                                                                                        \
    e = cudaFuncSetAttribute(tc::coupling_tc_kernel<INV, DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                             (int)smem);                                                                      \
    if (e == cudaSuccess) tc::coupling_tc_kernel<INV, DHV><<<grid, tc::kThreads, smem, st>>>(a);              \
  } while (0)
  const bool inv = direction == TNF_INVERSE;
  if (D == 64) { if (inv) TNF_TC_LAUNCH(true, 32); else TNF_TC_LAUNCH(false, 32); }
  else if (D == 128) { if (inv) TNF_TC_LAUNCH(true, 64); else TNF_TC_LAUNCH(false, 64); }
  else { if (inv) TNF_TC_LAUNCH(true, 128); else TNF_TC_LAUNCH(false, 128); }
#undef TNF_TC_LAUNCH
  if (e != cudaSuccess) {
    set_error("tnf_coupling_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  return check_launch("tnf_coupling_tc");
}

int tnf_tc_selftest_gemm(const float* A, const float* W, float* out, int K, int N, tnf_stream_t stream) {
  TNF_REQUIRE(A && W && out, TNF_ERR_ARG, "tnf_tc_selftest_gemm: null pointer");
  TNF_REQUIRE(K >= 16 && K <= 256 && K % 16 == 0 && N >= 16 && N <= 256 && N % 16 == 0, TNF_ERR_ARG,
              "tnf_tc_selftest_gemm: need 16 <= K,N <= 256, multiples of 16");
  const size_t smem = (size_t)K * N * 2 + 64;
  cudaError_t e = cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tnf_tc_selftest_gemm: %s", cudaGetErrorString(e));
    return (int)e;
  }
  tc::selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, out, K, N);
  return check_launch("tnf_tc_selftest_gemm");
}

}  // extern "C"
