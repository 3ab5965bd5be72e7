// Backward of the RealNVP coupling layer on tcgen05 tensor cores (shared weights, bf16 conditioner): tnf_coupling_tc_bwd.
//
// Reference: autograd through torch_nf/bijectors.py:145-242 (RealNVP.forward / inverse_and_log_det, _t_s_layer).
// One 128-row tile per CTA at a time, thread = sample row = TMEM lane.  Per tile and net (t, then s) the conditioner is
// RECOMPUTED from the saved layer input (jobs F1, F2, F3: the forward GEMMs), the coupling's output gradients are turned
// into the gradients of t and s, and these are propagated back through the nets (jobs B3, B2, B1: the same GEMM shape
// with TRANSPOSED weight images):
//     d3 = dL/d(t|s)            (128 x D/2)
//     d2 = (d3 . W3^T) * (1 - h2^2)      d1 = (d2 . W2^T) * (1 - h1^2)      dx = d1 . W1^T
// All activations live in tensor memory (two 256-column regions R0, R1 that ping-pong between jobs; a tanh / tanh' phase
// rewrites its accumulator chunk IN PLACE as the bf16 A operand of the next job, as in coupling_tc6).  The matrices
// the WEIGHT gradients need (h1, h2, d1, d2, d3 of both nets, bf16) are stored to a caller-provided workspace; the
// weight gradients themselves are plain GEMMs over the whole batch (dW2 = h1^T d2 ...: K = rows), which the host side
// runs as library GEMMs (ops.coupling_tc_bwd_param_grads) - they have no fusion partner here: their accumulators
// (2 x 82k floats) exceed tensor memory, and their K dimension is the one this kernel tiles over.
//
// Deliberately simple control: the 16 warps (thread = row, four column owners per TMEM lane quadrant) move through the
// twelve jobs of a tile in lockstep (one mbarrier for "MMAs of this job complete", __syncthreads between phases);
// thread 0 also issues the MMAs and keeps four weight slots (two big, two small) filled with cp.async.bulk one use ahead.
#include "tc_common.cuh"

namespace tnf {
namespace tcb {
using namespace tc;

constexpr int kEpi = 16;
constexpr int kThreadsB = kEpi * 32;   // 4 warps per SM sub-partition: up to 128 registers per thread

__host__ __device__ inline bool shape_supported_b(int D, int U, int L) {
  return (D == 64 || D == 128) && (U == 128 || U == 256) && L == 2;
}

struct ShapeB {
  int D, U, DH, upper, c_off, t_off, NH;
  __host__ __device__ ShapeB(int D_, int U_, int upper_) : D(D_), U(U_), upper(upper_) {
    DH = D / 2;
    c_off = upper ? 0 : DH;
    t_off = upper ? DH : 0;
    NH = U / 128;
  }
  __host__ __device__ int64_t small_bytes() const { return (int64_t)U * DH * 2; }
  __host__ __device__ int64_t big_bytes() const { return (int64_t)128 * U * 2; }
  // per net: F1 | F2[NH] | F3 | B3 | B2[NH] | B1
  __host__ __device__ int64_t oF1() const { return 0; }
  __host__ __device__ int64_t oF2() const { return small_bytes(); }
  __host__ __device__ int64_t oF3() const { return oF2() + NH * big_bytes(); }
  __host__ __device__ int64_t oB3() const { return oF3() + small_bytes(); }
  __host__ __device__ int64_t oB2() const { return oB3() + small_bytes(); }
  __host__ __device__ int64_t oB1() const { return oB2() + NH * big_bytes(); }
  __host__ __device__ int64_t net_bytes() const { return oB1() + small_bytes(); }
  __host__ __device__ int bias_per_net() const { return 2 * U + DH; }
  __host__ __device__ int64_t bias_off() const { return 2 * net_bytes(); }
  __host__ __device__ int64_t packed_bytes() const { return bias_off() + 2 * (int64_t)bias_per_net() * 4; }
  // flat parameter row (bijectors.py:224-235): per layer [W_t | W_s | b_t | b_s]
  __host__ __device__ int64_t src_layer(int l) const {
    int64_t o = 0;
    if (l >= 1) o += 2 * (int64_t)DH * U + 2 * U;
    if (l >= 2) o += 2 * (int64_t)U * U + 2 * U;
    return o;
  }
  // workspace (bf16 elements): [net][h1, h2, d1, d2][rows][U + 16] then [net][rows][DH] (d3).  The 16 pad columns of
  // h1 / h2 hold [1, 0, ..., 0]: (h | 1)^T d is the weight gradient AND (last row) the bias gradient in one GEMM
  __host__ __device__ int64_t pitch() const { return U + 16; }
  __host__ __device__ int64_t ws_mat(int net, int kind, int64_t rows) const { return ((int64_t)(net * 4 + kind) * rows) * pitch(); }
  __host__ __device__ int64_t ws_d3(int net, int64_t rows) const { return 8 * rows * pitch() + (int64_t)net * rows * DH; }
  // ... then xa [rows][xw()]: the conditioning half as the layer saw it (after the folded affine), bf16, followed by a
  // ones column and zeros - the activation operand of the first layer's weight-gradient GEMM
  __host__ __device__ int xw() const { return DH + 1 <= 64 ? 64 : 128; }
  __host__ __device__ int64_t ws_xa(int64_t rows) const { return 8 * rows * pitch() + 2 * rows * DH; }
  __host__ __device__ int64_t ws_elems(int64_t rows) const { return ws_xa(rows) + rows * xw(); }
};

// ---------------------------------------------------------------- weight packing
// blockIdx.y = image: net * 8 + {F1, F2a, F2b, F3, B3, B2a, B2b, B1}; element (n, k) of an N x K image is
// params[base + n * sn + k * sk]
__global__ void pack_b_kernel(const float* __restrict__ params, unsigned char* __restrict__ packed, ShapeB sh) {
  const int img = blockIdx.y, net = img >> 3, kind = img & 7;
  const int U = sh.U, DH = sh.DH;
  int N = 0, K = 0;
  int64_t base = 0, sn = 0, sk = 0, dst = (int64_t)net * sh.net_bytes();
  const int64_t w1 = sh.src_layer(0) + (int64_t)net * DH * U;       // (DH, U): W1[k][j]
  const int64_t w2 = sh.src_layer(1) + (int64_t)net * U * U;        // (U, U)
  const int64_t w3 = sh.src_layer(2) + (int64_t)net * U * DH;       // (U, DH)
  switch (kind) {
    case 0: N = U; K = DH; base = w1; sn = 1; sk = U; dst += sh.oF1(); break;
    case 1: case 2: {
      const int h = kind - 1;
      if (h >= sh.NH) return;
      N = 128; K = U; base = w2 + 128 * h; sn = 1; sk = U; dst += sh.oF2() + h * sh.big_bytes();
      break;
    }
    case 3: N = DH; K = U; base = w3; sn = 1; sk = DH; dst += sh.oF3(); break;
    case 4: N = U; K = DH; base = w3; sn = DH; sk = 1; dst += sh.oB3(); break;
    case 5: case 6: {
      const int h = kind - 5;
      if (h >= sh.NH) return;
      N = 128; K = U; base = w2 + (int64_t)128 * h * U; sn = U; sk = 1; dst += sh.oB2() + h * sh.big_bytes();
      break;
    }
    default: N = DH; K = U; base = w1; sn = U; sk = 1; dst += sh.oB1(); break;
  }
  const int total = N * K;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int n = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(packed + dst + img_off(n, k, N)) = __float2bfloat16_rn(params[base + n * sn + k * sk]);
  }
  if (kind == 0 && blockIdx.x == 0) {   // biases of this net: [b1 U][b2 U][b3 DH], fp32
    float* bdst = reinterpret_cast<float*>(packed + sh.bias_off()) + net * sh.bias_per_net();
    for (int i = threadIdx.x; i < sh.bias_per_net(); i += blockDim.x) {
      int l, j;
      if (i < U) { l = 0; j = i; } else if (i < 2 * U) { l = 1; j = i - U; } else { l = 2; j = i - 2 * U; }
      const int64_t K_l = l == 0 ? DH : U, J_l = l == 2 ? DH : U;
      bdst[i] = params[sh.src_layer(l) + 2 * K_l * J_l + (net ? J_l : 0) + j];
    }
  }
}

// ---------------------------------------------------------------- kernel
struct ArgsB {
  const float* z_in; const unsigned char* packed; const float* g_z_out; const float* g_ld; float* g_z_in;
  __nv_bfloat16* ws;
  const float* pre_scale; const float* pre_shift;   // per-column affine applied to z_in on load (a folded BatchNorm), or NULL
  int64_t rows;
  int D, U, upper, inverse;
};

struct __align__(16) CtrlB {
  uint64_t full_s[2], full_b[2], mma_done;
  uint32_t tmem_base, pad;
};

template <int DH, int U>
__host__ __device__ constexpr size_t smem_bytes_b() {
  return (size_t)2 * 128 * U * 2 + (size_t)2 * U * DH * 2 + (size_t)kTileM * DH * 2 + (size_t)2 * (2 * U + DH) * 4 +
         sizeof(CtrlB) + (size_t)4 * DH * 4;
}

// 256-bit global accesses (sm_100): a thread's 32-column piece of a bf16 row is two of them instead of four 128-bit
// ones - every lane of such an instruction is in another 128-byte line (32 L1 wavefronts), so the instruction count is
// what the LSU pipe pays for.  (Measured alternative: the pieces transposed through a per-warp shared-memory tile into
// 8-rows-x-64-byte stores - fewer global wavefronts but as many shared-memory ones: 2.90 -> 2.62 ms, dropped for this.)
__device__ __forceinline__ void ld256_cg(const void* p, uint32_t* r) {
  asm volatile("ld.global.cg.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc as two words]   (cta_group::1)
__device__ __forceinline__ void umma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int V> struct IC { static constexpr int value = V; };
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

template <int DH, int U>
__global__ void __launch_bounds__(kThreadsB, 1) coupling_tcb_kernel(ArgsB a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const ShapeB sh(a.D, a.U, a.upper);
  constexpr int NH = U / 128;          // N = 128 column blocks of a hidden layer
  constexpr int NC = U / 32;           // 32-column chunks of a hidden accumulator
  constexpr int W = DH / 4;            // final-layer columns per thread
  constexpr uint32_t kBig = 128u * U * 2u, kSmall = (uint32_t)U * DH * 2u;
  unsigned char* sBig = smem_raw;                      // 2 slots
  unsigned char* sSmall = sBig + 2 * (size_t)kBig;     // 2 slots
  unsigned char* sX = sSmall + 2 * (size_t)kSmall;     // conditioning half, bf16 A image (128 x DH)
  float* sBias = reinterpret_cast<float*>(sX + (size_t)kTileM * DH * 2);   // [net][b1 U | b2 U | b3 DH]
  CtrlB& ct = *reinterpret_cast<CtrlB*>(sBias + 2 * (2 * U + DH));
  float* sPs = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(CtrlB));   // pre_scale [D]
  float* sPb = sPs + 2 * DH;                                                                       // pre_shift [D]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ctl = threadIdx.x == 0;                   // no dedicated control warp: thread 0 also issues copies and MMAs
  const int q = warp & 3, cq = (warp >> 2) & 3;        // TMEM lane quadrant, chunk owner
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const int64_t cnt = (int64_t)blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(&ct.full_s[0], 1); mbar_init(&ct.full_s[1], 1);
    mbar_init(&ct.full_b[0], 1); mbar_init(&ct.full_b[1], 1);
    mbar_init(&ct.mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(&ct.tmem_base, 512);
  {
    const float* gb = reinterpret_cast<const float*>(a.packed + sh.bias_off());
    for (int i = threadIdx.x; i < 2 * (2 * U + DH); i += blockDim.x) sBias[i] = gb[i];
    for (int i = threadIdx.x; i < 2 * DH; i += blockDim.x) {
      sPs[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      sPb[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;
  const uint32_t R0 = tmem, R1 = tmem + 256u;
  const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
  const int r_tile = q * 32 + lane;

  // ---- weight slots (thread 0): use u of the small sequence / v of the big sequence
  const int64_t u_total = 8 * cnt, v_total = 4 * NH * cnt;
  auto load_small = [&](int64_t u) {
    const int u8 = (int)(u & 7), net = (u8 >> 1) & 1, slot = (int)(u & 1);
    const int64_t off = (u8 < 4) ? ((u8 & 1) ? sh.oF3() : sh.oF1()) : ((u8 & 1) ? sh.oB1() : sh.oB3());
    mbar_arrive_expect_tx(&ct.full_s[slot], kSmall);
    bulk_g2s(sSmall + (size_t)slot * kSmall, a.packed + (int64_t)net * sh.net_bytes() + off, kSmall, &ct.full_s[slot]);
  };
  auto load_big = [&](int64_t v) {
    const int v8 = (int)(v % (4 * NH)), g = v8 / NH, h = v8 % NH, slot = (int)(v & 1);
    const int64_t off = (g < 2 ? sh.oF2() : sh.oB2()) + (int64_t)h * kBig;
    mbar_arrive_expect_tx(&ct.full_b[slot], kBig);
    bulk_g2s(sBig + (size_t)slot * kBig, a.packed + (int64_t)(g & 1) * sh.net_bytes() + off, kBig, &ct.full_b[slot]);
  };
  int64_t u_use = 0, v_use = 0;
  uint32_t ps = 0, pb = 0, dpar = 0;   // parities: small slots (bit = slot), big slots, mma_done
  if (ctl) {
    if (u_total > 0) { load_small(0); load_small(1); }
    if (v_total > 0) { load_big(0); load_big(1); }
  }

  // K-major SWIZZLE_NONE descriptors (tc_common.cuh): one K = 16 step reads two K groups.  The descriptors are kept as
  // 32-bit words (only the start address in the low word moves) and the K loops are unrolled with compile-time shapes:
  // one thread issues every MMA, its instruction count per MMA is the kernel's serial overhead.
  const uint32_t a_dhi = (uint32_t)(make_desc(0u, kTileM) >> 32);
  const uint32_t x_lo = (uint32_t)make_desc(smem_u32(sX), kTileM);
  // AMODE 0: A = the conditioning-half image in shared memory; 1: tensor memory, in-place (strided) layout;
  // 2: tensor memory, compact layout
  auto issue_small = [&](uint32_t d_tmem, uint32_t a_tmem, auto amode_c, auto K_c, auto N_c) {
    constexpr int AMODE = decltype(amode_c)::value, K = decltype(K_c)::value, N = decltype(N_c)::value;
    const int slot = (int)(u_use & 1);
    mbar_wait(&ct.full_s[slot], (ps >> slot) & 1u);
    ps ^= 1u << slot;
    tc_fence_after();
    const uint32_t idesc = make_idesc(N);
    const uint64_t bd = make_desc(smem_u32(sSmall + (size_t)slot * kSmall), N);
    const uint32_t b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
#pragma unroll
    for (int kk = 0; kk < K / 16; ++kk) {
      if (AMODE == 0) {
        umma_ss2(d_tmem, x_lo + (uint32_t)(kk * 2 * kTileM), a_dhi, b_lo + (uint32_t)(kk * 2 * N), b_hi, idesc, kk > 0 ? 1u : 0u);
      } else {
        const uint32_t col = AMODE == 1 ? (uint32_t)(32 * (kk >> 1) + 8 * (kk & 1)) : (uint32_t)(8 * kk);
        umma_ts2(d_tmem, a_tmem + col, b_lo + (uint32_t)(kk * 2 * N), b_hi, idesc, kk > 0 ? 1u : 0u);
      }
    }
    tc_commit(&ct.mma_done);
  };
  auto issue_big = [&](uint32_t d_tmem, uint32_t a_tmem) {   // hidden -> hidden: NH column blocks of N = 128, K = U, A strided
    const uint32_t idesc = make_idesc(128);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int slot = (int)((v_use + h) & 1);
      mbar_wait(&ct.full_b[slot], (pb >> slot) & 1u);
      pb ^= 1u << slot;
      tc_fence_after();
      const uint64_t bd = make_desc(smem_u32(sBig + (size_t)slot * kBig), 128);
      const uint32_t b_lo = (uint32_t)bd, b_hi = (uint32_t)(bd >> 32);
#pragma unroll
      for (int kk = 0; kk < U / 16; ++kk)
        umma_ts2(d_tmem + (uint32_t)(128 * h), a_tmem + (uint32_t)(32 * (kk >> 1) + 8 * (kk & 1)), b_lo + (uint32_t)(kk * 256),
                 b_hi, idesc, kk > 0 ? 1u : 0u);
    }
    tc_commit(&ct.mma_done);
  };
  // every thread: the job's MMAs are complete; thread 0 then refills the slot(s) the job used
  auto wait_job = [&](bool big) {
    mbar_wait(&ct.mma_done, dpar);
    dpar ^= 1u;
    tc_fence_after();
    if (ctl) {
      if (big) {
        for (int h = 0; h < NH; ++h)
          if (v_use + 2 + h < v_total && (NH == 2 || h == 0)) load_big(v_use + 2 + h);
        v_use += NH;
      } else {
        if (u_use + 2 < u_total) load_small(u_use + 2);
        u_use += 1;
      }
    }
    __syncwarp();
  };
  auto phase_end = [&]() {
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };

  // 32 columns (16 packed words) of this thread's row -> workspace matrix `mat` (rows x U, bf16), columns 32c..
  auto store_chunk = [&](__nv_bfloat16* mat, int c, const uint32_t* o, bool valid_own, int64_t own_row) {
    if (valid_own) {
      __nv_bfloat16* dst = mat + own_row * (U + 16) + 32 * c;
      st256(dst, o);
      st256(dst + 16, o + 8);
    }
  };

  for (int64_t it = 0; it < cnt; ++it) {
    const int64_t tile = (int64_t)blockIdx.x + it * gridDim.x;
    const int64_t row = tile * kTileM + r_tile;
    const bool valid = row < a.rows;
    // ---- conditioning half -> bf16 A image, and (with the ones column) -> workspace for the first layer's weight gradient
    {
      constexpr int XW = DH + 1 <= 64 ? 64 : 128;
      __nv_bfloat16* xa = a.ws + sh.ws_xa(a.rows);
      for (int i = threadIdx.x; i < kTileM * XW / 8; i += kEpi * 32) {
        const int r = i / (XW / 8), k8 = (i % (XW / 8)) * 8;
        const int64_t grow = tile * kTileM + r;
        uint4 pk = make_uint4(k8 == DH ? 0x00003f80u : 0u, 0u, 0u, 0u);      // [1, 0, ..] at column DH, zeros after it
        if (k8 < DH) {
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (grow < a.rows) {
            const float4* src = reinterpret_cast<const float4*>(a.z_in + grow * sh.D + sh.c_off + k8);
            v0 = __ldg(src); v1 = __ldg(src + 1);
          }
          const float4 s0 = *reinterpret_cast<const float4*>(sPs + sh.c_off + k8), s1 = *reinterpret_cast<const float4*>(sPs + sh.c_off + k8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(sPb + sh.c_off + k8), b1 = *reinterpret_cast<const float4*>(sPb + sh.c_off + k8 + 4);
          v0.x = fmaf(v0.x, s0.x, b0.x); v0.y = fmaf(v0.y, s0.y, b0.y); v0.z = fmaf(v0.z, s0.z, b0.z); v0.w = fmaf(v0.w, s0.w, b0.w);
          v1.x = fmaf(v1.x, s1.x, b1.x); v1.y = fmaf(v1.y, s1.y, b1.y); v1.z = fmaf(v1.z, s1.z, b1.z); v1.w = fmaf(v1.w, s1.w, b1.w);
          pk = make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
          *reinterpret_cast<uint4*>(sX + img_off(r, k8, kTileM)) = pk;
        }
        if (grow < a.rows) *reinterpret_cast<uint4*>(xa + grow * XW + k8) = pk;
      }
      fence_async_smem();
    }
    __syncthreads();

    float tv[W], sv[W];      // t, then reused; s -> gradient of s
    float dxt[W];
    constexpr int NI = NC / 4;
    uint32_t hv[NI][16];     // tanh outputs of this thread's chunks, re-read for the tanh' phases of the backward part
    // Requested one phase AHEAD of their use (before the MMAs of the previous, short job are even issued), so that the
    // L2 latency hides behind that job and its epilogue
    auto prefetch_h = [&](int net, int l) {
      const __nv_bfloat16* hsrc = a.ws + sh.ws_mat(net, l, a.rows) + row * (U + 16);
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        if (valid) {
          ld256_cg(hsrc + 32 * (cq + 4 * i), hv[i]);
          ld256_cg(hsrc + 32 * (cq + 4 * i) + 16, hv[i] + 8);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) hv[i][j] = 0u;
        }
      }
    };
    if (valid) {   // L2 prefetch of the rows' other operands (read once, a few phases from now): no registers held
      const float* pf = cq == 0 ? a.z_in + row * sh.D + sh.t_off
                                : (cq == 1 ? (a.g_z_out ? a.g_z_out + row * sh.D + sh.t_off : nullptr)
                                           : (cq == 2 ? (a.g_z_out ? a.g_z_out + row * sh.D + sh.c_off : nullptr) : nullptr));
      if (pf) {
#pragma unroll
        for (int j = 0; j < DH * 4; j += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + j / 4));
      }
    }
    // ======================= forward recompute, both nets =======================
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      const float* bias = sBias + net * (2 * U + DH);
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const uint32_t reg = l == 0 ? R0 : R1;
        if (ctl) {
          if (l == 0) issue_small(R0, 0u, IC<0>{}, IC<DH>{}, IC<U>{});
          else issue_big(R1, R0);
        }
        __syncwarp();
        wait_job(l == 1);
        {
          __nv_bfloat16* hmat = a.ws + sh.ws_mat(net, l, a.rows);
          uint32_t x[NI][32];
#pragma unroll
          for (int i = 0; i < NI; ++i) tmem_ld32(reg + lane_addr + (uint32_t)(32 * (cq + 4 * i)), x[i]);   // all in flight
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int c = cq + 4 * i;
            uint32_t o[16];
            const float* bl = bias + l * U + 32 * c;
#pragma unroll
            for (int j = 0; j < 32; j += 2)
              o[j >> 1] = pack_bf16(tanh_fast(__uint_as_float(x[i][j]) + bl[j]), tanh_fast(__uint_as_float(x[i][j + 1]) + bl[j + 1]));
            tmem_st16(reg + lane_addr + (uint32_t)(32 * c), o);
            store_chunk(hmat, c, o, valid, row);
          }
          if (cq == 0 && valid) {   // the pad columns: [1, 0, ..., 0] (bias-gradient row of the weight-gradient GEMM)
            const uint32_t one[8] = {0x00003f80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            st256(hmat + row * (U + 16) + U, one);
          }
          tc_wait_st();
        }
        phase_end();
      }
      // final layer: t or s
      float z2[W], g2[W], gl = 0.f;
      if (net == 1) {        // operands of the coupling phase and of the first tanh' phase: in flight during this job
        gl = (valid && a.g_ld) ? a.g_ld[row] : 0.f;
#pragma unroll
        for (int j = 0; j < W; j += 4) {
          float4 zz = make_float4(0.f, 0.f, 0.f, 0.f), gg = zz;
          if (valid) {
            zz = __ldg(reinterpret_cast<const float4*>(a.z_in + row * sh.D + sh.t_off + cq * W + j));
            if (a.g_z_out) gg = __ldg(reinterpret_cast<const float4*>(a.g_z_out + row * sh.D + sh.t_off + cq * W + j));
          }
          const float* ps = sPs + sh.t_off + cq * W + j;
          const float* pb = sPb + sh.t_off + cq * W + j;
          z2[j] = fmaf(zz.x, ps[0], pb[0]); z2[j + 1] = fmaf(zz.y, ps[1], pb[1]);
          z2[j + 2] = fmaf(zz.z, ps[2], pb[2]); z2[j + 3] = fmaf(zz.w, ps[3], pb[3]);
          g2[j] = gg.x; g2[j + 1] = gg.y; g2[j + 2] = gg.z; g2[j + 3] = gg.w;
        }
        prefetch_h(0, 1);
      }
      if (ctl) issue_small(R0, R1, IC<1>{}, IC<U>{}, IC<DH>{});
      __syncwarp();
      wait_job(false);
      {
        uint32_t o[W];
        if (W == 8) tmem_ld8(R0 + lane_addr + (uint32_t)(cq * W), reinterpret_cast<uint32_t(&)[8]>(o));
        else tmem_ld16(R0 + lane_addr + (uint32_t)(cq * W), reinterpret_cast<uint32_t(&)[16]>(o));
        tc_wait_ld();
        const float* bl = bias + 2 * U + cq * W;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const float v = __uint_as_float(o[j]) + bl[j];
          if (net == 0) tv[j] = v; else sv[j] = v;
        }
        if (net == 1) {
          // ---- the coupling itself: gradients of t, s and of the transformed half (bijectors.py:172,198)
          float gz2[W];
          uint32_t pt[W / 2], psn[W / 2];
#pragma unroll
          for (int j = 0; j < W; ++j) {
            float gt, gs;
            if (a.inverse) {       // x2 = (z2 - t) exp(-s)
              const float e = exp2_fast(-sv[j] * 1.4426950408889634f);
              const float x2 = (z2[j] - tv[j]) * e;
              gz2[j] = g2[j] * e;
              gt = -gz2[j];
              gs = gl - g2[j] * x2;
            } else {               // y2 = z2 exp(s) + t
              const float e = exp2_fast(sv[j] * 1.4426950408889634f);
              gz2[j] = g2[j] * e;
              gt = g2[j];
              gs = gl + g2[j] * z2[j] * e;
            }
            tv[j] = gt;
            sv[j] = gs;
          }
#pragma unroll
          for (int j = 0; j < W; j += 2) { pt[j >> 1] = pack_bf16(tv[j], tv[j + 1]); psn[j >> 1] = pack_bf16(sv[j], sv[j + 1]); }
          // d3 of the t net: compact bf16 A operand behind the final accumulator (columns DH .. DH + DH/2 of R0)
          if (W == 8) tmem_st4(R0 + lane_addr + (uint32_t)(DH + cq * (W / 2)), pt);
          else tmem_st8(R0 + lane_addr + (uint32_t)(DH + cq * (W / 2)), pt);
          if (valid) {
            float* gdst = a.g_z_in + row * sh.D + sh.t_off + cq * W;
#pragma unroll
            for (int j = 0; j < W; j += 4) {   // gradient w.r.t. the layer input BEFORE the folded affine
              const float* ps = sPs + sh.t_off + cq * W + j;
              *reinterpret_cast<float4*>(gdst + j) = make_float4(gz2[j] * ps[0], gz2[j + 1] * ps[1], gz2[j + 2] * ps[2], gz2[j + 3] * ps[3]);
            }
            uint4* d3t = reinterpret_cast<uint4*>(a.ws + sh.ws_d3(0, a.rows) + row * DH + cq * W);
            uint4* d3s = reinterpret_cast<uint4*>(a.ws + sh.ws_d3(1, a.rows) + row * DH + cq * W);
#pragma unroll
            for (int j = 0; j < W / 8; ++j) {
              d3t[j] = make_uint4(pt[4 * j], pt[4 * j + 1], pt[4 * j + 2], pt[4 * j + 3]);
              d3s[j] = make_uint4(psn[4 * j], psn[4 * j + 1], psn[4 * j + 2], psn[4 * j + 3]);
            }
          }
          tc_wait_st();
        }
      }
      phase_end();
    }
    // ======================= backward through the nets =======================
#pragma unroll
    for (int net = 0; net < 2; ++net) {
      // B3: d h2 = d3 . W3^T   (K = DH, N = U), then B2: d h1 = d2 . W2^T
#pragma unroll
      for (int l = 1; l >= 0; --l) {
        const uint32_t reg = l == 1 ? R1 : R0;
        if (ctl) {
          if (l == 1) issue_small(R1, R0 + (uint32_t)DH, IC<2>{}, IC<DH>{}, IC<U>{});
          else issue_big(R0, R1);
        }
        __syncwarp();
        // tanh outputs of this phase: the B3 phases' were requested a phase ago; the B2 phases' go out now and hide
        // behind the long hidden-layer job
        if (l == 0) prefetch_h(net, 0);
        wait_job(l == 0);
        {
          __nv_bfloat16* dmat = a.ws + sh.ws_mat(net, 2 + l, a.rows);
#pragma unroll
          for (int i = 0; i < NI; ++i) {
            const int c = cq + 4 * i;
            uint32_t x[32], o[16];
            tmem_ld32(reg + lane_addr + (uint32_t)(32 * c), x);
            tc_wait_ld();
            const uint32_t* hp = hv[i];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float h0 = bf_lo(hp[j]), h1 = bf_hi(hp[j]);
              o[j] = pack_bf16(__uint_as_float(x[2 * j]) * fmaf(-h0, h0, 1.f), __uint_as_float(x[2 * j + 1]) * fmaf(-h1, h1, 1.f));
            }
            tmem_st16(reg + lane_addr + (uint32_t)(32 * c), o);
            store_chunk(dmat, c, o, valid, row);
          }
          tc_wait_st();
        }
        phase_end();
      }
      // B1: dx = d1 . W1^T   (K = U, N = DH)
      float4 gx1[W / 4];
      if (net == 0) {
        prefetch_h(1, 1);
      } else {
#pragma unroll
        for (int j = 0; j < W / 4; ++j)
          gx1[j] = (valid && a.g_z_out) ? __ldg(reinterpret_cast<const float4*>(a.g_z_out + row * sh.D + sh.c_off + cq * W) + j)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (ctl) issue_small(R1, R0, IC<1>{}, IC<U>{}, IC<DH>{});
      __syncwarp();
      wait_job(false);
      {
        uint32_t o[W];
        if (W == 8) tmem_ld8(R1 + lane_addr + (uint32_t)(cq * W), reinterpret_cast<uint32_t(&)[8]>(o));
        else tmem_ld16(R1 + lane_addr + (uint32_t)(cq * W), reinterpret_cast<uint32_t(&)[16]>(o));
        tc_wait_ld();
        if (net == 0) {
#pragma unroll
          for (int j = 0; j < W; ++j) dxt[j] = __uint_as_float(o[j]);
          uint32_t psn[W / 2];
#pragma unroll
          for (int j = 0; j < W; j += 2) psn[j >> 1] = pack_bf16(sv[j], sv[j + 1]);
          if (W == 8) tmem_st4(R0 + lane_addr + (uint32_t)(DH + cq * (W / 2)), psn);
          else tmem_st8(R0 + lane_addr + (uint32_t)(DH + cq * (W / 2)), psn);
          tc_wait_st();
        } else if (valid) {
          float* gdst = a.g_z_in + row * sh.D + sh.c_off + cq * W;
#pragma unroll
          for (int j = 0; j < W; j += 4) {
            const float4 gg = gx1[j >> 2];
            const float* ps = sPs + sh.c_off + cq * W + j;
            *reinterpret_cast<float4*>(gdst + j) =
                make_float4((gg.x + dxt[j] + __uint_as_float(o[j])) * ps[0], (gg.y + dxt[j + 1] + __uint_as_float(o[j + 1])) * ps[1],
                            (gg.z + dxt[j + 2] + __uint_as_float(o[j + 2])) * ps[2], (gg.w + dxt[j + 3] + __uint_as_float(o[j + 3])) * ps[3]);
          }
        }
      }
      phase_end();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

template <int DH, int U>
static int launch_b(const ArgsB& a, int grid, cudaStream_t st) {
  constexpr size_t smem = smem_bytes_b<DH, U>();
  cudaError_t e = cudaFuncSetAttribute(coupling_tcb_kernel<DH, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tnf_coupling_tc_bwd: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  coupling_tcb_kernel<DH, U><<<grid, kThreadsB, smem, st>>>(a);
  return check_launch("tnf_coupling_tc_bwd");
}

}  // namespace tcb
}  // namespace tnf

using namespace tnf;

extern "C" {

int tnf_tc_bwd_supported(int D, int U, int L) { return tcb::shape_supported_b(D, U, L) ? 1 : 0; }

size_t tnf_tc_bwd_packed_bytes(int D, int U, int L) {
  if (!tcb::shape_supported_b(D, U, L)) return 0;
  return (size_t)tcb::ShapeB(D, U, 1).packed_bytes();
}

size_t tnf_tc_bwd_workspace_bytes(int64_t rows, int D, int U, int L) {
  if (!tcb::shape_supported_b(D, U, L) || rows <= 0) return 0;
  return (size_t)tcb::ShapeB(D, U, 1).ws_elems(rows) * 2;
}

int tnf_tc_bwd_pack(const float* params, void* packed, int D, int U, int L, int transform_upper, tnf_stream_t stream) {
  TNF_REQUIRE(params && packed, TNF_ERR_ARG, "tnf_tc_bwd_pack: null pointer");
  TNF_REQUIRE(tcb::shape_supported_b(D, U, L), TNF_ERR_UNSUPPORTED, "tnf_tc_bwd_pack: D=%d U=%d L=%d not supported", D, U, L);
  TNF_REQUIRE(((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN, "tnf_tc_bwd_pack: packed must be 16-byte aligned");
  tcb::ShapeB sh(D, U, transform_upper);
  tcb::pack_b_kernel<<<dim3(32, 16), 256, 0, (cudaStream_t)stream>>>(params, (unsigned char*)packed, sh);
  return check_launch("tnf_tc_bwd_pack");
}

int tnf_coupling_tc_bwd(const float* z_in, const void* packed, const float* g_z_out, const float* g_log_det,
                        float* g_z_in, void* workspace, int64_t rows, int D, int U, int L, int transform_upper,
                        int direction, const float* pre_scale, const float* pre_shift, tnf_stream_t stream) {
  TNF_REQUIRE(z_in && packed && g_z_in && workspace, TNF_ERR_ARG, "tnf_coupling_tc_bwd: null pointer");
  TNF_REQUIRE(tcb::shape_supported_b(D, U, L), TNF_ERR_UNSUPPORTED, "tnf_coupling_tc_bwd: D=%d U=%d L=%d not supported", D, U, L);
  TNF_REQUIRE(rows >= 0, TNF_ERR_ARG, "tnf_coupling_tc_bwd: rows < 0");
  TNF_REQUIRE((((uintptr_t)z_in | (uintptr_t)packed | (uintptr_t)g_z_in | (uintptr_t)workspace | (uintptr_t)g_z_out) & 15) == 0,
              TNF_ERR_ALIGN, "tnf_coupling_tc_bwd: pointers must be 16-byte aligned");
  TNF_REQUIRE((rows * (int64_t)U * 2) % 16 == 0 && (rows * (int64_t)(D / 2) * 2) % 16 == 0, TNF_ERR_ALIGN,
              "tnf_coupling_tc_bwd: workspace matrices must stay 16-byte aligned");
  if (rows == 0) return 0;
  tcb::ArgsB a;
  a.z_in = z_in; a.packed = (const unsigned char*)packed; a.g_z_out = g_z_out; a.g_ld = g_log_det; a.g_z_in = g_z_in;
  a.ws = (__nv_bfloat16*)workspace; a.rows = rows; a.D = D; a.U = U; a.upper = transform_upper;
  a.inverse = direction == TNF_INVERSE;
  a.pre_scale = pre_scale; a.pre_shift = pre_shift;
  const int64_t n_tiles = (rows + tc::kTileM - 1) / tc::kTileM;
  const int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  if (D == 64 && U == 256) return tcb::launch_b<32, 256>(a, grid, st);
  if (D == 64 && U == 128) return tcb::launch_b<32, 128>(a, grid, st);
  if (D == 128 && U == 256) return tcb::launch_b<64, 256>(a, grid, st);
  return tcb::launch_b<64, 128>(a, grid, st);
}

}  // extern "C"
