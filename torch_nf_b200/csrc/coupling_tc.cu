// RealNVP coupling layer fused on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-sample tiles.  For each tile the two
// conditioner nets (shift t, scale s) run entirely on-chip:
//
//   z1 (fp32, HBM) --cvt--> A1 bf16 in SMEM (UMMA K-major operand image)
//   H_net = A . W_l         tcgen05.mma kind::f16 (SS), D fp32 in TMEM, 256 columns per net
//   A_net = bf16(tanh(H_net + b_l))   epilogue warps: tcgen05.ld -> MUFU.TANH -> pack -> st.shared
//   ...
//   (t, s) = A . W_L + b_L ;  z2' = t + z2 e^s  |  (z2 - t)/e^s ;  log_det (+)= sum s
//
// Activations never leave the SM: HBM traffic is z in, z out and the log-det
// read-modify-write (520 B per sample-layer at D = 64).  Weights are pre-packed
// (tnf_tc_pack) into the shared-memory images the UMMA descriptors expect
// (K-major, SWIZZLE_NONE, 8-row x 16-byte core matrices) and streamed from L2
// through a ring of 16 KB stages by one producer warp (cp.async.bulk + mbarrier
// complete_tx).
//
// Three kernels share this file (tnf_coupling_tc picks by shape):
//
// coupling_tc3_kernel (D <= 128, the C3 path): two tiles in flight per CTA.  16 epilogue
// warps in two groups of 8 (per group two warps per TMEM lane quadrant, each taking the
// accumulator chunks of one parity); each group owns a tile, 256 TMEM columns and its own
// A images.  The MMA warp serves the groups alternately in a static order, so while one
// group runs its MUFU-bound tanh epilogue the tensor pipe computes the other group's next
// layer.  Biases are added by one extra K=16 MMA per layer (constant [1,1,0..] A image x
// bf16 hi/lo bias image), the MMA jobs are fully unrolled per hidden width, and two I/O
// warps load the conditioning half with coalesced 16-byte accesses, build the bf16 A1
// image and write the pass-through half.  Warp roles (640 threads): 0-15 epilogue,
// 16 MMA issuer, 17 weight producer, 18-19 I/O.
//
// coupling_tc2_kernel (D = 256): the images of two tiles do not fit shared memory, so one
// tile is in flight, with two 256-column accumulators: layer l+1's MMAs trail the tanh
// epilogue of layer l chunk by chunk (per-chunk mbarriers).  Same building blocks.
//
// coupling_tc_kernel (diagnostic variant 1): the first design - 8 epilogue warps in two
// groups of 4 (thread = one sample row, all I/O and the bias adds in the epilogue
// threads), generic MMA issue loop, 320 threads.  Kept as an independent cross-check.
//
// Reference semantics: torch_nf/bijectors.py:145-242 (RealNVP).
#include "tc_common.cuh"

namespace tnf {
namespace tc {

// ---------------------------------------------------------------- weight packing
// Packed buffer = the weight stream in consumption order:
//   for layer l in 0..L: for net in {t, s}: for column half hb: for stage s:
//     image of stage_k(K,N) x N bf16 (N = N_of(l); see img_off, rows = N)
// followed by the fp32 biases [net][layer][unit] and the bias operand images (Shape::bias_img_off).
__global__ void pack_kernel(const float* __restrict__ params, unsigned char* __restrict__ packed, Shape sh) {
  const int64_t per_net = sh.net_weight_elems();
  const int64_t total = 2 * per_net;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // idx enumerates SOURCE elements in (layer, net, k, j) order
    int64_t rem = idx;
    int l = 0;
    int64_t src_off = 0;  // offset of layer l inside the reference parameter row
    int64_t dst_off = 0;  // element offset of layer l inside the packed stream
    for (;; ++l) {
      const int64_t n_el = (int64_t)sh.K_of(l) * sh.J_of(l);
      if (rem < 2 * n_el) break;
      rem -= 2 * n_el;
      src_off += 2 * n_el + 2 * sh.J_of(l);
      dst_off += 2 * n_el;
    }
    const int K = sh.K_of(l), J = sh.J_of(l), N = sh.N_of(l);
    const int64_t n_el = (int64_t)K * J;
    const int net = rem >= n_el;
    rem -= net * n_el;
    const int k = (int)(rem / J), j = (int)(rem % J);
    const float w = params[src_off + net * n_el + rem];   // W_t then W_s, (K, J) row-major, x @ W
    const int hb = j / N, n = j % N;
    const int ks = sh.stage_k(K, N);
    const int st = k / ks, kk = k % ks;
    unsigned char* dst =
        packed + (dst_off + net * n_el + (int64_t)hb * K * N + (int64_t)st * ks * N) * 2 + img_off(n, kk, N);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(w);
    const int NB = J / 2;   // split-N copy for the CTA-pair kernel
    *reinterpret_cast<__nv_bfloat16*>(packed + sh.split_w_off(l, net, j / NB) + img_off(j % NB, k, NB)) = __float2bfloat16_rn(w);
    if (sh.L >= 2 && l == sh.L - 1) {   // N-half copy of the last hidden layer: half h = j / (U/2), rank = (j / (U/4)) % 2
      const int Q = sh.U / 4;
      *reinterpret_cast<__nv_bfloat16*>(packed + sh.half_w_off(net, j / (2 * Q), (j / Q) & 1) + img_off(j % Q, k, Q)) =
          __float2bfloat16_rn(w);
    }
  }
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
    const float b = params[src_off + 2 * (int64_t)K * J + (net ? J : 0) + rem];
    bias_dst[idx] = b;
    unsigned char* img = packed + sh.bias_img_off(l, net);
    const __nv_bfloat16 hi = __float2bfloat16_rn(b);
    const __nv_bfloat16 lo = __float2bfloat16_rn(b - __bfloat162float(hi));
    unsigned char* img2 = packed + sh.split_b_off(l, net, rem / (J / 2));
    for (int kk = 0; kk < 16; ++kk) {
      const __nv_bfloat16 v = kk == 0 ? hi : (kk == 1 ? lo : __float2bfloat16_rn(0.f));
      *reinterpret_cast<__nv_bfloat16*>(img + img_off(rem, kk, J)) = v;
      *reinterpret_cast<__nv_bfloat16*>(img2 + img_off(rem % (J / 2), kk, J / 2)) = v;
    }
    if (sh.L >= 2) {   // resident 8-K-row bias blocks of coupling_tc5_kernel
      const bool split = l == sh.L - 1;
      const int Q = sh.U / 4;
      const int rk = split ? (rem / Q) & 1 : rem / (J / 2);
      const int h = split ? rem / (2 * Q) : 0;
      const int n = split ? rem % Q : rem % (J / 2);
      __nv_bfloat16* blk = reinterpret_cast<__nv_bfloat16*>(packed + sh.bias8_base(rk) + sh.bias8_off(l, net, h) + (int64_t)n * 16);
      blk[0] = hi; blk[1] = lo;
      for (int kk = 2; kk < 8; ++kk) blk[kk] = __float2bfloat16_rn(0.f);
    }
  }
}


// Tile ping-pong: epilogue group g (warps 4g..4g+3, one warp per TMEM lane quadrant) owns tile
// (2*it+g)*grid + cta of every iteration, TMEM columns [256g, 256g+256) and the A images of group g.
// The MMA warp issues the GEMMs in the static order  for net: for layer: for group  so that while one
// group runs its MUFU-bound epilogue the tensor pipe works for the other group.
template <bool kInverse, int DH>   // DH = D/2 = d_in = d_out
__global__ void __launch_bounds__(kThreads, 1) coupling_tc_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Shape sh(a.D, a.U, a.L, a.upper);
  const int S = a.n_stages;
  unsigned char* ring = smem_raw;
  const uint32_t stage_bytes = (uint32_t)sh.stage_elems() * 2;
  unsigned char* sA1 = ring + (size_t)S * stage_bytes;           // 2 images
  unsigned char* sAct = sA1 + 2 * sh.a1_bytes();                  // 2 images
  Ctrl& ct = *reinterpret_cast<Ctrl*>(sAct + 2 * sh.act_bytes());
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(Ctrl));
  const int nb = sh.net_bias_elems();
  float* s_pscale = s_bias + 2 * nb;
  float* s_pshift = s_pscale + sh.D;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const int G = a.n_groups == 1 ? 1 : 2;   // epilogue groups in use (2; 1 = diagnostic solo mode)
  const int64_t iters = (n_tiles + G * (int64_t)gridDim.x - 1) / (G * (int64_t)gridDim.x);
  const int64_t weight_bytes = 2 * sh.net_weight_elems() * 2;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&ct.a1_ready[g], 4);
      mbar_init(&ct.e_done[g], 4);
      mbar_init(&ct.h_ready[g][0], 1);
      mbar_init(&ct.h_ready[g][1], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps) tmem_alloc(&ct.tmem_base, 512);
  {
    const float* gb = reinterpret_cast<const float*>(a.packed + weight_bytes);
    for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) s_bias[i] = gb[i];
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;
  if (warp == kEpiWarps + 1) {
    // =============================== weight producer (one elected lane) ===============================
    // Job j of a group = (iteration, net, layer) = (j / JT, (j % JT) / (L+1), j % (L+1)), JT = 2(L+1) jobs per tile.
    // Static service order: group 1 runs half a tile (L+1 jobs) behind group 0, so one group's tile tail / head
    // (global loads and stores, no MUFU work) coincides with the other group's mid-tile tanh epilogues.
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      const int JT = 2 * (sh.L + 1);
      const int64_t n_jobs = iters * JT;
      const int shift = (G > 1) ? sh.L + 1 : 0;
      for (int64_t n = 0; n < n_jobs + shift; ++n) {
        for (int g = 0; g < G; ++g) {
          const int64_t j = n - (g ? shift : 0);
          if (j < 0 || j >= n_jobs) continue;
          const int jj = (int)(j % JT);
          const int net = jj / (sh.L + 1), l = jj % (sh.L + 1);
          // byte offset of (layer l, net) in the packed stream
          size_t off = 0;
          for (int i = 0; i < l; ++i) off += (size_t)2 * sh.K_of(i) * sh.J_of(i) * 2;
          const int K = sh.K_of(l), J = sh.J_of(l), N = sh.N_of(l);
          const int ks = sh.stage_k(K, N);
          const uint32_t bytes = (uint32_t)(ks * N * 2);
          const int n_st = sh.halves(l) * (K / ks);
          const unsigned char* nsrc = a.packed + off + (size_t)net * K * J * 2;
          for (int st = 0; st < n_st; ++st) {
            mbar_wait(&ct.w_empty[slot], phase ^ 1);
            mbar_arrive_expect_tx(&ct.w_full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * stage_bytes, nsrc + (size_t)st * bytes, bytes, &ct.w_full[slot]);
            if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kEpiWarps) {
    // =============================== MMA issuer (warp-uniform, elected lane issues) ===============================
    const bool leader = elect_one();
    uint32_t slot = 0, phase = 0, a1_phase = 0, e_phase = 0;
    const bool diag = a.dbg != nullptr;   // diagnostics: cycles waiting on epilogue / on weights
    long long t_dep = 0, t_w = 0, t_all = diag ? clock64() : 0;
    const uint32_t ring_addr = smem_u32(ring);
    const uint32_t a1_addr = smem_u32(sA1), act_addr = smem_u32(sAct);
    const uint32_t a1_sz = (uint32_t)sh.a1_bytes(), act_sz = (uint32_t)sh.act_bytes();
    const int JT = 2 * (sh.L + 1);
    const int64_t n_jobs = iters * JT;
    const int shift = (G > 1) ? sh.L + 1 : 0;   // see the producer: group 1 trails group 0 by half a tile
    for (int64_t n = 0; n < n_jobs + shift; ++n) {
      for (int g = 0; g < G; ++g) {
        const int64_t j = n - (g ? shift : 0);
        if (j < 0 || j >= n_jobs) continue;
        const int jj = (int)(j % JT);
        const int net = jj / (sh.L + 1), l = jj % (sh.L + 1);
        const int K = sh.K_of(l), N = sh.N_of(l), halves = sh.halves(l);
        const int ks = sh.stage_k(K, N);
        const uint32_t idesc = make_idesc(N);
        const long long c0 = diag ? clock64() : 0;
        if (net == 0 && l == 0) {
          mbar_wait(&ct.a1_ready[g], (a1_phase >> g) & 1);
          a1_phase ^= 1u << g;
        } else {
          mbar_wait(&ct.e_done[g], (e_phase >> g) & 1);
          e_phase ^= 1u << g;
        }
        tc_fence_after();
        if (diag) t_dep += clock64() - c0;
        const uint32_t a_addr = (l == 0) ? a1_addr + g * a1_sz : act_addr + g * act_sz;
        const uint64_t a_base = make_desc(a_addr, kTileM);
        const uint64_t b_base = make_desc(0u, N);                 // + (stage address >> 4)
        const uint32_t a_step = (2u * kTileM * 16u) >> 4, b_step = (2u * (uint32_t)N * 16u) >> 4;
        for (int hb = 0; hb < halves; ++hb) {
          const uint32_t d_tmem = tmem + (uint32_t)g * 256u + (uint32_t)(hb * sh.Nh);
          for (int k0 = 0; k0 < K; k0 += ks) {
            const long long c1 = diag ? clock64() : 0;
            mbar_wait(&ct.w_full[slot], phase);
            tc_fence_after();
            if (diag) t_w += clock64() - c1;
            // descriptors advance by constants: one K=16 step = two 8-column K groups = 2*rows*16 bytes
            const uint64_t ad = a_base + (uint64_t)(((uint32_t)k0 >> 4) * a_step);
            const uint64_t bd = b_base + (uint64_t)((ring_addr + slot * stage_bytes) >> 4);
            if (leader) {
              const uint32_t acc0 = k0 > 0 ? 1u : 0u;
              if (ks == 32) {
                umma_ss(d_tmem, ad, bd, idesc, acc0);
                umma_ss(d_tmem, ad + a_step, bd + b_step, idesc, 1u);
              } else {
#pragma unroll 4
                for (int kq = 0; kq < ks / 16; ++kq)
                  umma_ss(d_tmem, ad + (uint64_t)kq * a_step, bd + (uint64_t)kq * b_step, idesc, (kq > 0) ? 1u : acc0);
              }
              tc_commit(&ct.w_empty[slot]);
            }
            __syncwarp();
            if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
          }
          if (leader) tc_commit(&ct.h_ready[g][hb]);
          __syncwarp();
        }
      }
    }
    if (a.dbg != nullptr && blockIdx.x == 0 && leader) {
      a.dbg[2040] = t_dep; a.dbg[2041] = t_w; a.dbg[2042] = clock64() - t_all;
    }
  } else if ((warp >> 2) < G) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, g = warp >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r_tile = q * 32 + lane;            // row inside the tile
    const uint32_t hcol = tmem + lane_addr + (uint32_t)g * 256u;
    unsigned char* myA1 = sA1 + (size_t)g * sh.a1_bytes();
    unsigned char* myAct = sAct + (size_t)g * sh.act_bytes();
    uint64_t* my_h = &ct.h_ready[g][0];
    uint32_t h_phase = 0;
    const int n_pairs = sh.U / (2 * kChunk);     // accumulator chunks are processed two at a time
    const float kLog2e = 1.4426950408889634f;
    constexpr bool kRegs = DH <= 32;             // small D: keep z rows / t outputs in registers
    constexpr int kR = kRegs ? DH : 4;

    int dbg_n = 0;
    const bool dbg_on = a.dbg != nullptr && blockIdx.x == 0 && q == 0 && lane == 0;
    long long* dbg = a.dbg + g * 1024;
#define TNF_STAMP(tag)                                                                     \
  do {                                                                                     \
    if (dbg_on && dbg_n < 500) { dbg[2 * dbg_n] = (tag); dbg[2 * dbg_n + 1] = clock64(); ++dbg_n; } \
  } while (0)

    // Software-pipelined epilogue step.  `cur` holds chunk c's pre-activations (accumulator + bias), `acc` the raw
    // accumulator of chunk c+1 (landed).  First half of the step: 16 MUFU.TANH of chunk c interleaved with all 32
    // bias adds that turn `acc` into `nxt`; then the tcgen05.ld of chunk c+2 is issued into the now-free `acc`
    // (the __syncwarp pins it there: it may not sink below the later st.shared) so that its latency hides under
    // the second half's 16 MUFU.TANH; pack + st.shared of each group trail its tanh by one group.
    auto epi_step = [&](float (&cur)[32], float (&nxt)[32], uint32_t (&acc)[32], const float* bias_next, int c,
                        uint32_t next_ld_col) {
      const float4* b4 = reinterpret_cast<const float4*>(bias_next);
      unsigned char* dst = myAct + img_off(r_tile, c * kChunk, kTileM);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cur[j + e] = tanh_fast(cur[j + e]);
        if (j < 16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = b4[j / 2 + q];
            const int o = 2 * j + 4 * q;
            nxt[o] = __uint_as_float(acc[o]) + b.x;         nxt[o + 1] = __uint_as_float(acc[o + 1]) + b.y;
            nxt[o + 2] = __uint_as_float(acc[o + 2]) + b.z; nxt[o + 3] = __uint_as_float(acc[o + 3]) + b.w;
          }
        }
        if (j > 0) {
          const int k = j - 8;
          *reinterpret_cast<uint4*>(dst + (k >> 3) * (kTileM * 16)) =
              make_uint4(pack_bf16(cur[k], cur[k + 1]), pack_bf16(cur[k + 2], cur[k + 3]),
                         pack_bf16(cur[k + 4], cur[k + 5]), pack_bf16(cur[k + 6], cur[k + 7]));
        }
        if (j == 8) {
          tmem_ld32(next_ld_col, acc);
          __syncwarp();
        }
      }
      *reinterpret_cast<uint4*>(dst + 3 * (kTileM * 16)) =
          make_uint4(pack_bf16(cur[24], cur[25]), pack_bf16(cur[26], cur[27]), pack_bf16(cur[28], cur[29]),
                     pack_bf16(cur[30], cur[31]));
    };
    // conditioning half -> pre-affine -> bf16 A1 image -> publish; returns the pre-affined values in v
    auto publish_a1 = [&](const float* zin_regs, const float* zrow, bool valid, float (&v)[DH]) {
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        float4 in;
        if (kRegs) in = make_float4(zin_regs[kRegs ? j : 0], zin_regs[kRegs ? j + 1 : 0], zin_regs[kRegs ? j + 2 : 0],
                                    zin_regs[kRegs ? j + 3 : 0]);
        else in = valid ? *reinterpret_cast<const float4*>(zrow + sh.c_off + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[j] = fmaf(in.x, s_pscale[sh.c_off + j], s_pshift[sh.c_off + j]);
        v[j + 1] = fmaf(in.y, s_pscale[sh.c_off + j + 1], s_pshift[sh.c_off + j + 1]);
        v[j + 2] = fmaf(in.z, s_pscale[sh.c_off + j + 2], s_pshift[sh.c_off + j + 2]);
        v[j + 3] = fmaf(in.w, s_pscale[sh.c_off + j + 3], s_pshift[sh.c_off + j + 3]);
      }
#pragma unroll
      for (int j = 0; j < DH; j += 8) {
        uint4 p = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                             pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
        *reinterpret_cast<uint4*>(myA1 + img_off(r_tile, j, kTileM)) = p;
      }
      fence_async_smem();
      tc_fence_before();   // also orders this thread's tcgen05.ld of the previous tile's outputs
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.a1_ready[g]);
    };

    float st_y = 0.f, st_y2 = 0.f, st_v = 0.f, st_v2 = 0.f;   // per-lane column sums (lane = column of the half)
    const bool want_stats = kRegs && a.stat_partials != nullptr;
    // ---- first tile of this group: load and publish A1, store the pass-through half
    {
      const int64_t tile = (int64_t)g * gridDim.x + blockIdx.x;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = tile < n_tiles && row < a.rows;
      float zc[kR];
      if (kRegs) {
#pragma unroll
        for (int j = 0; j < kR; j += 4) {
          float4 t4 = valid ? *reinterpret_cast<const float4*>(a.z_in + row * sh.D + sh.c_off + j) : make_float4(0.f, 0.f, 0.f, 0.f);
          zc[j] = t4.x; zc[j + 1] = t4.y; zc[j + 2] = t4.z; zc[j + 3] = t4.w;
        }
      }
      float v[DH];
      publish_a1(zc, a.z_in + row * sh.D, valid, v);
      if (valid) {
#pragma unroll
        for (int j = 0; j < DH; j += 4)
          *reinterpret_cast<float4*>(a.z_out + row * sh.D + sh.c_off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
      if (want_stats) {
        if (DH == 32) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] = valid ? v[DH == 32 ? j : 0] : 0.f; s2[j] = s1[j] * s1[j]; }
          st_v += warp_transpose_sum(s1, lane);
          st_v2 += warp_transpose_sum(s2, lane);
        }
      }
    }
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t tile = (it * G + g) * (int64_t)gridDim.x + blockIdx.x;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = tile < n_tiles && row < a.rows;
      const int64_t nrow = row + G * (int64_t)gridDim.x * kTileM;      // this thread's row in the next iteration
      const bool has_next = it + 1 < iters;
      const bool nvalid = has_next && (tile + G * (int64_t)gridDim.x) < n_tiles && nrow < a.rows;
      const float* zrow = a.z_in + row * sh.D;
      float* orow = a.z_out + row * sh.D;
      TNF_STAMP(100);
      if (nvalid) {   // pull the next tile's row and log-det into L2 a whole tile ahead of their use
        for (int b = 0; b < sh.D * 4; b += 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(a.z_in + nrow * sh.D) + b));
        if (a.accum != TNF_LD_WRITE) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.log_det + nrow));
      }
      float tv[kR];   // t-net output (registers when it fits; else parked in z_out)
      float zt[kR];   // transformed half of this tile
      float zn[kR];   // conditioning half of the next tile
      float ld_old = 0.f, ld_sum = 0.f;
#pragma unroll
      for (int net = 0; net < 2; ++net) {
        const float* bias = s_bias + net * nb;
        // ---- hidden layers: accumulator -> tanh -> A image
#pragma unroll 1
        for (int l = 0; l < sh.L; ++l) {
          const float* bl = bias + l * sh.U;
          TNF_STAMP(200 + net * 10 + l);
          mbar_wait(my_h, h_phase);
          h_phase ^= 1;
          tc_fence_after();
          TNF_STAMP(300 + net * 10 + l);
          const int n_chunks = 2 * n_pairs;
          uint32_t acc[32];
          float xa[32], xb[32];
          tmem_ld32(hcol, acc);
          tc_wait_ld();
          {
            const float4* b4 = reinterpret_cast<const float4*>(bl);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = b4[j / 4];
              xa[j] = __uint_as_float(acc[j]) + b.x;         xa[j + 1] = __uint_as_float(acc[j + 1]) + b.y;
              xa[j + 2] = __uint_as_float(acc[j + 2]) + b.z; xa[j + 3] = __uint_as_float(acc[j + 3]) + b.w;
            }
          }
          tmem_ld32(hcol + (uint32_t)kChunk, acc);
#pragma unroll 1
          for (int c = 0; c < n_chunks; c += 2) {
            // chunk c (xa) while chunk c+1 is prepared into xb and chunk c+2 is fetched
            // (the last iteration re-reads its own last chunk - harmless - so that the loop stays branch-free)
            const int c2 = (c + 2 < n_chunks) ? c + 2 : c + 1;
            const int c3 = (c + 3 < n_chunks) ? c + 3 : c + 1;
            tc_wait_ld();
            epi_step(xa, xb, acc, bl + (c + 1) * kChunk, c, hcol + (uint32_t)(c2 * kChunk));
            // chunk c+1 (xb) while chunk c+2 is prepared into xa and chunk c+3 is fetched
            tc_wait_ld();
            epi_step(xb, xa, acc, bl + c2 * kChunk, c + 1, hcol + (uint32_t)(c3 * kChunk));
          }
          tc_wait_ld();
          fence_async_smem();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ct.e_done[g]);
        }
        // ---- while the final-layer MMAs of the s-net run: issue every global load still needed
        if (net == 1) {
          if (valid && a.accum != TNF_LD_WRITE) ld_old = a.log_det[row];
          if (kRegs) {
#pragma unroll
            for (int j = 0; j < kR; j += 4) {
              float4 t4 = valid ? *reinterpret_cast<const float4*>(zrow + sh.t_off + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              zt[j] = t4.x; zt[j + 1] = t4.y; zt[j + 2] = t4.z; zt[j + 3] = t4.w;
            }
#pragma unroll
            for (int j = 0; j < kR; j += 4) {
              float4 t4 = nvalid ? *reinterpret_cast<const float4*>(a.z_in + nrow * sh.D + sh.c_off + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              zn[j] = t4.x; zn[j + 1] = t4.y; zn[j + 2] = t4.z; zn[j + 3] = t4.w;
            }
          }
        }
        // ---- final layer of this net
        const float* bL = bias + sh.L * sh.U;
        TNF_STAMP(400 + net);
        mbar_wait(my_h, h_phase);
        h_phase ^= 1;
        tc_fence_after();
        TNF_STAMP(500 + net);
        if (net == 0) {
#pragma unroll
          for (int j0 = 0; j0 < DH; j0 += 16) {
            uint32_t o[16];
            tmem_ld16(hcol + (uint32_t)j0, o);
            tc_wait_ld();
            if (kRegs) {
#pragma unroll
              for (int j = 0; j < 16; ++j) tv[kRegs ? j0 + j : 0] = __uint_as_float(o[j]) + bL[j0 + j];
            } else if (valid) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(orow + sh.t_off + j0 + j) =
                    make_float4(__uint_as_float(o[j]) + bL[j0 + j], __uint_as_float(o[j + 1]) + bL[j0 + j + 1],
                                __uint_as_float(o[j + 2]) + bL[j0 + j + 2], __uint_as_float(o[j + 3]) + bL[j0 + j + 3]);
            }
          }
          tc_fence_before();   // t read out: the accumulator may be overwritten by the s-net
          __syncwarp();
          if (lane == 0) mbar_arrive(&ct.e_done[g]);
        } else if (kRegs) {
          // s read out; y kept in registers so that the next tile's A1 can be published BEFORE any global
          // store is issued (the async-proxy fence would otherwise wait for those stores to drain)
          float y[kR];
#pragma unroll
          for (int j0 = 0; j0 < kR; j0 += 16) {
            uint32_t o[16];
            tmem_ld16(hcol + (uint32_t)j0, o);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = sh.t_off + j0 + j;
              const float zin = fmaf(zt[j0 + j], s_pscale[col], s_pshift[col]);
              const float sv = __uint_as_float(o[j]) + bL[j0 + j];
              ld_sum += sv;
              y[j0 + j] = kInverse ? (zin - tv[j0 + j]) * exp2_fast(-sv * kLog2e)
                                   : fmaf(zin, exp2_fast(sv * kLog2e), tv[j0 + j]);
            }
          }
          float v[DH];
          TNF_STAMP(601);
          if (has_next) publish_a1(zn, nullptr, nvalid, v);
          TNF_STAMP(602);
          if (valid) {
#pragma unroll
            for (int j = 0; j < kR; j += 4)
              *reinterpret_cast<float4*>(orow + sh.t_off + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          }
          if (nvalid) {
#pragma unroll
            for (int j = 0; j < DH; j += 4)
              *reinterpret_cast<float4*>(a.z_out + nrow * sh.D + sh.c_off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (want_stats) {
            if (DH == 32) {
              float s1[32], s2[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) { s1[j] = valid ? y[DH == 32 ? j : 0] : 0.f; s2[j] = s1[j] * s1[j]; }
              st_y += warp_transpose_sum(s1, lane);
              st_y2 += warp_transpose_sum(s2, lane);
#pragma unroll
              for (int j = 0; j < 32; ++j) { s1[j] = nvalid ? v[DH == 32 ? j : 0] : 0.f; s2[j] = s1[j] * s1[j]; }
              st_v += warp_transpose_sum(s1, lane);
              st_v2 += warp_transpose_sum(s2, lane);
            }
          }
        } else {
          // large D: stream the transformed half in 16-column pieces (t was parked in z_out)
#pragma unroll 1
          for (int j0 = 0; j0 < DH; j0 += 16) {
            uint32_t o[16];
            tmem_ld16(hcol + (uint32_t)j0, o);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int col = sh.t_off + j0 + j;
              float4 zv = valid ? *reinterpret_cast<const float4*>(zrow + col) : make_float4(0.f, 0.f, 0.f, 0.f);
              float4 tq = valid ? *reinterpret_cast<const float4*>(orow + col) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float zz[4] = {zv.x, zv.y, zv.z, zv.w};
              const float tt[4] = {tq.x, tq.y, tq.z, tq.w};
              float yy[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float zin = fmaf(zz[e], s_pscale[col + e], s_pshift[col + e]);
                const float sv = __uint_as_float(o[j + e]) + bL[j0 + j + e];
                ld_sum += sv;
                yy[e] = kInverse ? (zin - tt[e]) * exp2_fast(-sv * kLog2e) : fmaf(zin, exp2_fast(sv * kLog2e), tt[e]);
              }
              if (valid) *reinterpret_cast<float4*>(orow + col) = make_float4(yy[0], yy[1], yy[2], yy[3]);
            }
          }
          if (has_next) {
            float v[DH];
            publish_a1(nullptr, a.z_in + nrow * sh.D, nvalid, v);
            if (nvalid) {
#pragma unroll
              for (int j = 0; j < DH; j += 4)
                *reinterpret_cast<float4*>(a.z_out + nrow * sh.D + sh.c_off + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          }
        }
      }
      if (valid) {
        float* o = a.log_det + row;
        if (a.accum == TNF_LD_WRITE) *o = ld_sum;
        else if (a.accum == TNF_LD_ADD) *o = ld_old + ld_sum;
        else *o = ld_old - ld_sum;
      }
      TNF_STAMP(600);
    }
#undef TNF_STAMP
    if (want_stats) {   // one [2][D] block of doubles per epilogue warp
      double* out = a.stat_partials + ((size_t)blockIdx.x * kEpiWarps + warp) * 2 * sh.D;
      out[sh.t_off + lane] = (double)st_y;
      out[sh.D + sh.t_off + lane] = (double)st_y2;
      out[sh.c_off + lane] = (double)st_v;
      out[sh.D + sh.c_off + lane] = (double)st_v2;
    }
  } else if (warp < kEpiWarps && a.stat_partials != nullptr) {   // idle epilogue group (solo mode): zero block
    double* out = a.stat_partials + ((size_t)blockIdx.x * kEpiWarps + warp) * 2 * sh.D;
    for (int i = lane; i < 2 * sh.D; i += 32) out[i] = 0.0;
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps) tmem_dealloc(tmem, 512);
}

// ================================================================ building blocks of the two-tile kernel (D <= 128)

// One GEMM job of the MMA warp, fully unrolled over its 32-wide K chunks: the job's bias MMA first (accumulator :=
// ones image . bias image, a stage of its own), then per chunk two K=16 MMAs whose descriptor low words differ from the
// job's base by compile-time constants, an mbarrier wait when a new weight stage begins and a commit when it is used
// up.  Run by the whole warp (uniform control flow), the elected lane issues.  A single warp executes dependent
// instructions only every ~5 cycles, so a generic loop with run-time strides (~80 instructions per chunk) was the
// bottleneck of the first kernel; unrolled with constant offsets a chunk step is ~20 instructions.
template <int K, int N>
__device__ __forceinline__ void mma_job(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo_ring, uint32_t b_hi,
                                        uint32_t ones_lo, uint32_t wfull0, uint32_t wempty0, uint32_t S, uint32_t& slot,
                                        uint32_t& phase, bool leader, long long* t_w = nullptr) {
  constexpr int KS = (kStageElems / N) < K ? (kStageElems / N) : K;   // K rows per weight stage
  constexpr int CPS = KS / kChunk;                                     // chunks per stage
  constexpr uint32_t kStage16 = kStageBytes >> 4;
  const uint32_t idesc = make_idesc(N);
  {
    const long long c0 = t_w ? clock64() : 0;
    mbar_wait_addr(wfull0 + slot * 8u, phase);
    if (t_w) *t_w += clock64() - c0;
    tc_fence_after();
    if (leader) {
      umma_ss2(d_tmem, ones_lo, a_hi, b_lo_ring + slot * kStage16, b_hi, idesc, 0u);
      tc_commit_addr(wempty0 + slot * 8u);
    }
    if (++slot == S) { slot = 0; phase ^= 1; }
  }
#pragma unroll
  for (int c = 0; c < K / kChunk; ++c) {
    if (c % CPS == 0) {
      const long long c0 = t_w ? clock64() : 0;
      mbar_wait_addr(wfull0 + slot * 8u, phase);
      if (t_w) *t_w += clock64() - c0;
      tc_fence_after();
    }
    if (leader) {
      const uint32_t b_lo = b_lo_ring + slot * kStage16 + (uint32_t)((c % CPS) * 4 * N);
      umma_ss2(d_tmem, a_lo + 512u * c, a_hi, b_lo, b_hi, idesc, 1u);
      umma_ss2(d_tmem, a_lo + 512u * c + 256u, a_hi, b_lo + 2u * N, b_hi, idesc, 1u);
      if (c % CPS == CPS - 1) tc_commit_addr(wempty0 + slot * 8u);
    }
    if (c % CPS == CPS - 1) {
      if (++slot == S) { slot = 0; phase ^= 1; }
    }
  }
}

// ================================================================ single-tile pipelined kernel (D = 256)
// At D = 256 two A1 images (32 KB each) and two activation images do not fit next to a weight ring, so the two-tile
// kernel is out.  Here ONE tile is in flight per CTA and all 16 epilogue warps work on it (warp w: TMEM lane quadrant
// w%4, accumulator chunks c = w/4 mod 4, final-layer columns [32*(w/4), +32)).  The 512 TMEM columns hold TWO
// 256-column accumulators, so the MMA warp runs layer l+1 into one while the epilogue still drains layer l from the
// other: it issues the two K=16 MMAs of K-chunk c as soon as the four warps owning chunk c have published those 32
// activation columns (act_ready[c]) - the tensor pipe trails the MUFU-bound tanh epilogue by one chunk.  The A1 image
// is double buffered (tile parity) and written, with the pass-through half, by two I/O warps using coalesced
// 16-byte accesses; the t-net output is parked in z_out (L2) until the s-net is done, as in the first kernel.
constexpr int kThreads2 = (kEpiWarps2 + 4) * 32;
constexpr int kMaxJobs = 6;   // L + 1 <= 6

struct __align__(16) Ctrl2 {
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t a1_ready[2];     // 2 I/O warps: A1 image of a tile written
  uint64_t a1_free[2];      // tcgen05.commit: both layer-0 jobs of the tile have read the A1 image
  uint64_t act_ready[8];    // 4 epilogue warps: activation chunk c (32 K-columns, all 128 rows) written
  uint64_t h_ready[2][kMaxJobs];   // tcgen05.commit per (net, layer): accumulator complete
  uint32_t tmem_base;
  uint32_t pad;
};
// dynamic shared memory:
//   [ring: n_stages x 16 KB][A1 x2][Act][ones 4 KB][Ctrl2][pre_scale D][pre_shift D][ld partial 3 x 128]
__host__ __device__ inline size_t smem_bytes2(const Shape& sh, int n_stages) {
  return (size_t)n_stages * kStageBytes + 2 * sh.a1_bytes() + sh.act_bytes() + kOnesBytes + sizeof(Ctrl2) +
         (size_t)(2 * sh.D + 3 * kTileM) * sizeof(float);
}

// mma_job whose K chunks wait for the epilogue of the previous layer chunk by chunk (act_ready[c], parity e_par).
// The bias MMA is issued after the first activation chunk is published: that also tells that the epilogue warps have
// drained whatever the destination accumulator held.
template <int K, int N>
__device__ __forceinline__ void mma_job_trailing(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo_ring,
                                                 uint32_t b_hi, uint32_t ones_lo, uint32_t act_bar0, uint32_t e_par,
                                                 uint32_t wfull0, uint32_t wempty0, uint32_t S, uint32_t& slot,
                                                 uint32_t& phase, bool leader) {
  constexpr int KS = (kStageElems / N) < K ? (kStageElems / N) : K;   // K rows per weight stage
  constexpr int CPS = KS / kChunk;                                     // chunks per stage
  constexpr uint32_t kStage16 = kStageBytes >> 4;
  const uint32_t idesc = make_idesc(N);
#pragma unroll
  for (int c = 0; c < K / kChunk; ++c) {
    mbar_wait_addr(act_bar0 + 8u * c, e_par);
    if (c == 0) {
      mbar_wait_addr(wfull0 + slot * 8u, phase);
      tc_fence_after();
      if (leader) {
        umma_ss2(d_tmem, ones_lo, a_hi, b_lo_ring + slot * kStage16, b_hi, idesc, 0u);
        tc_commit_addr(wempty0 + slot * 8u);
      }
      if (++slot == S) { slot = 0; phase ^= 1; }
    }
    if (c % CPS == 0) mbar_wait_addr(wfull0 + slot * 8u, phase);
    tc_fence_after();
    if (leader) {
      const uint32_t b_lo = b_lo_ring + slot * kStage16 + (uint32_t)((c % CPS) * 4 * N);
      umma_ss2(d_tmem, a_lo + 512u * c, a_hi, b_lo, b_hi, idesc, 1u);
      umma_ss2(d_tmem, a_lo + 512u * c + 256u, a_hi, b_lo + 2u * N, b_hi, idesc, 1u);
      if (c % CPS == CPS - 1) tc_commit_addr(wempty0 + slot * 8u);
    }
    if (c % CPS == CPS - 1) {
      if (++slot == S) { slot = 0; phase ^= 1; }
    }
  }
}

template <bool kInverse, int DH, int U_>   // DH = D/2 = 128; U_ = hidden units
__global__ void __launch_bounds__(kThreads2, 1) coupling_tc2_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Shape sh(a.D, a.U, a.L, a.upper);
  const int S = a.n_stages;
  unsigned char* ring = smem_raw;
  unsigned char* sA1 = ring + (size_t)S * kStageBytes;              // 2 images (tile parity)
  unsigned char* sAct = sA1 + 2 * sh.a1_bytes();                    // 1 image
  unsigned char* sOnes = sAct + sh.act_bytes();                     // constant A image for the bias MMA
  Ctrl2& ct = *reinterpret_cast<Ctrl2*>(sOnes + kOnesBytes);
  float* s_pscale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(Ctrl2));
  float* s_pshift = s_pscale + sh.D;
  float* s_ldp = s_pshift + sh.D;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles blockIdx.x + i*grid
  constexpr int n_chunks = U_ / kChunk;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&ct.a1_ready[b], 2);
      mbar_init(&ct.a1_free[b], 1);
    }
    for (int c = 0; c < 8; ++c) mbar_init(&ct.act_ready[c], 4);
    for (int n = 0; n < 2; ++n)
      for (int l = 0; l < kMaxJobs; ++l) mbar_init(&ct.h_ready[n][l], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps2) tmem_alloc(&ct.tmem_base, 512);
  {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x)   // row r: K columns 0 and 1 are 1.0 (bf16 0x3f80)
      reinterpret_cast<uint32_t*>(sOnes)[i] = (i < kTileM * 4 && (i & 3) == 0) ? 0x3f803f80u : 0u;
    fence_async_smem();
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;

  if (warp == kEpiWarps2 + 1) {
    // =============================== weight producer (one elected lane) ===============================
    // 16 KB stages whatever grouping tnf_tc_pack used: a packed layer is a sequence of 8-row K groups, so any
    // multiple of 8 K rows is a valid stage
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        for (int net = 0; net < 2; ++net) {
          size_t off = 0;
          for (int l = 0; l <= sh.L; ++l) {
            const int K = sh.K_of(l), J = sh.J_of(l);
            const int ks = (kStageElems / J) < K ? (kStageElems / J) : K;
            const uint32_t bytes = (uint32_t)(ks * J * 2);
            const unsigned char* nsrc = a.packed + off + (size_t)net * K * J * 2;
            {   // the job's bias operand image travels as a stage of its own, ahead of the weights
              mbar_wait(&ct.w_empty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&ct.w_full[slot], (uint32_t)J * 32u);
              bulk_g2s(ring + (size_t)slot * kStageBytes, a.packed + sh.bias_img_off(l, net), (uint32_t)J * 32u, &ct.w_full[slot]);
              if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
            }
            for (int st = 0; st < K / ks; ++st) {
              mbar_wait(&ct.w_empty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&ct.w_full[slot], bytes);
              bulk_g2s(ring + (size_t)slot * kStageBytes, nsrc + (size_t)st * bytes, bytes, &ct.w_full[slot]);
              if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
            }
            off += (size_t)2 * K * J * 2;
          }
        }
      }
    }
  } else if (warp == kEpiWarps2) {
    // =============================== MMA issuer (warp-uniform, elected lane issues) ===============================
    const bool leader = elect_one();
    uint32_t slot = 0, phase = 0, e_par = 0;
    int cur = 0;   // accumulator buffer of the most recent hidden-layer job
    const uint32_t a1_sz16 = (uint32_t)sh.a1_bytes() >> 4;
    const uint32_t act_bar0 = smem_u32(&ct.act_ready[0]);
    const uint32_t wfull0 = smem_u32(&ct.w_full[0]), wempty0 = smem_u32(&ct.w_empty[0]);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint64_t a1_desc = make_desc(smem_u32(sA1), kTileM), act_desc = make_desc(smem_u32(sAct), kTileM);
    const uint32_t a_hi = (uint32_t)(a1_desc >> 32);
    const uint64_t bU_desc = make_desc(0u, U_), bF_desc = make_desc(0u, DH);
    const uint32_t bU_lo = (uint32_t)bU_desc + ring16, bU_hi = (uint32_t)(bU_desc >> 32);
    const uint32_t bF_lo = (uint32_t)bF_desc + ring16, bF_hi = (uint32_t)(bF_desc >> 32);
    const uint32_t act_lo = (uint32_t)act_desc;
    const uint32_t ones_lo = (uint32_t)make_desc(smem_u32(sOnes), kTileM);
    for (int64_t it = 0; it < my_tiles; ++it) {
      const uint32_t ab = (uint32_t)(it & 1);
      const uint32_t a1_lo = (uint32_t)a1_desc + ab * a1_sz16;
      mbar_wait(&ct.a1_ready[ab], (uint32_t)((it >> 1) & 1));
#pragma unroll 1
      for (int net = 0; net < 2; ++net) {
        // layer 0: A1 image, accumulator = the buffer the previous epilogue phase has just drained
        mma_job<DH, U_>(tmem + (uint32_t)cur * 256u, a1_lo, a_hi, bU_lo, bU_hi, ones_lo, wfull0, wempty0, (uint32_t)S, slot,
                        phase, leader);
        if (leader) {
          tc_commit(&ct.h_ready[net][0]);
          if (net == 1) tc_commit(&ct.a1_free[ab]);
        }
        // hidden layers 1..L-1, each trailing the epilogue of the layer before it chunk by chunk
#pragma unroll 1
        for (int l = 1; l < sh.L; ++l) {
          cur ^= 1;
          mma_job_trailing<U_, U_>(tmem + (uint32_t)cur * 256u, act_lo, a_hi, bU_lo, bU_hi, ones_lo, act_bar0, e_par, wfull0,
                                   wempty0, (uint32_t)S, slot, phase, leader);
          e_par ^= 1;
          if (leader) tc_commit(&ct.h_ready[net][l]);
        }
        // final layer -> the other buffer's first DH columns
        mma_job_trailing<U_, DH>(tmem + (uint32_t)(cur ^ 1) * 256u, act_lo, a_hi, bF_lo, bF_hi, ones_lo, act_bar0, e_par,
                                 wfull0, wempty0, (uint32_t)S, slot, phase, leader);
        e_par ^= 1;
        if (leader) tc_commit(&ct.h_ready[net][sh.L]);
        __syncwarp();
      }
    }
  } else if (warp < kEpiWarps2) {
    // =============================== epilogue warps ===============================
    const int q = warp & 3, par = warp >> 2;  // par in 0..3
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r_tile = q * 32 + lane;
    constexpr int W = DH / 4;                 // final-layer columns per thread (32)
    const float kLog2e = 1.4426950408889634f;
    int cur = 0;

    // accumulator chunk c (bias already added by the bias MMA) -> MUFU.TANH -> bf16 -> A image, then publish the
    // chunk: generic-proxy writes -> async proxy, one arrival per warp
    auto epi_step = [&](uint32_t hcol, int c) {
      uint32_t x[32];
      tmem_ld32(hcol + (uint32_t)(c * kChunk), x);
      tc_wait_ld();
      unsigned char* dst = sAct + img_off(r_tile, c * kChunk, kTileM);
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) x[j + e] = __float_as_uint(tanh_fast(__uint_as_float(x[j + e])));
        *reinterpret_cast<uint4*>(dst + (j >> 3) * (kTileM * 16)) =
            make_uint4(pack_bf16(__uint_as_float(x[j]), __uint_as_float(x[j + 1])),
                       pack_bf16(__uint_as_float(x[j + 2]), __uint_as_float(x[j + 3])),
                       pack_bf16(__uint_as_float(x[j + 4]), __uint_as_float(x[j + 5])),
                       pack_bf16(__uint_as_float(x[j + 6]), __uint_as_float(x[j + 7])));
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.act_ready[c]);
    };
    auto quad_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory"); };

    for (int64_t it = 0; it < my_tiles; ++it) {
      const int64_t tile = it * (int64_t)gridDim.x + blockIdx.x;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = row < a.rows;
      const uint32_t h_par = (uint32_t)(it & 1);
      const float* zrow = a.z_in + row * sh.D + sh.t_off + par * W;
      float* orow = a.z_out + row * sh.D + sh.t_off + par * W;
      float ld_old = 0.f;
      if (valid) asm volatile("prefetch.global.L2 [%0];" ::"l"(zrow));   // 128 B = this thread's piece of the transformed half
#pragma unroll
      for (int net = 0; net < 2; ++net) {
#pragma unroll 1
        for (int l = 0; l < sh.L; ++l) {
          if (l > 0) cur ^= 1;
          const uint32_t hcol = tmem + lane_addr + (uint32_t)cur * 256u;
          mbar_wait(&ct.h_ready[net][l], h_par);
          tc_fence_after();
#pragma unroll 1
          for (int c = par; c < n_chunks; c += 4) epi_step(hcol, c);
        }
        // ---- final layer of this net: W columns per thread, in pieces of 16
        const uint32_t fcol = tmem + lane_addr + (uint32_t)(cur ^ 1) * 256u + (uint32_t)(par * W);
        if (net == 0 && par == 0 && valid && a.accum != TNF_LD_WRITE) ld_old = a.log_det[row];
        float zin_s[16], tt_s[16];
        auto load_piece = [&](int j0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 z4 = valid ? __ldg(reinterpret_cast<const float4*>(zrow + j0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 t4 = valid ? *reinterpret_cast<const float4*>(orow + j0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            zin_s[j] = z4.x; zin_s[j + 1] = z4.y; zin_s[j + 2] = z4.z; zin_s[j + 3] = z4.w;
            tt_s[j] = t4.x; tt_s[j + 1] = t4.y; tt_s[j + 2] = t4.z; tt_s[j + 3] = t4.w;
          }
        };
        if (net == 1) load_piece(0);   // in flight while the final-layer MMAs of the s-net run
        mbar_wait(&ct.h_ready[net][sh.L], h_par);
        tc_fence_after();
        if (net == 0) {
          // t (bias included) is parked in z_out until the s-net is done (it stays in L2; the same thread reads it back)
#pragma unroll 1
          for (int j0 = 0; j0 < W; j0 += 16) {
            uint32_t o[16];
            tmem_ld16(fcol + (uint32_t)j0, o);
            tc_wait_ld();
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(orow + j0 + j) = make_float4(__uint_as_float(o[j]), __uint_as_float(o[j + 1]),
                                                                        __uint_as_float(o[j + 2]), __uint_as_float(o[j + 3]));
            }
          }
          tc_fence_before();
          quad_sync();   // every column of t is read out before chunk 0 of the s-net lets MMAs overwrite it
        } else {
          float ld_sum = 0.f;
#pragma unroll 1
          for (int j0 = 0; j0 < W; j0 += 16) {
            uint32_t o[16];
            tmem_ld16(fcol + (uint32_t)j0, o);
            tc_wait_ld();
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = sh.t_off + par * W + j0 + j;
              const float zz = fmaf(zin_s[j], s_pscale[col], s_pshift[col]);
              const float sv = __uint_as_float(o[j]);
              ld_sum += sv;
              y[j] = kInverse ? (zz - tt_s[j]) * exp2_fast(-sv * kLog2e) : fmaf(zz, exp2_fast(sv * kLog2e), tt_s[j]);
            }
            if (j0 + 16 < W) load_piece(j0 + 16);
            if (valid) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(orow + j0 + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
            }
          }
          tc_fence_before();
          if (par > 0) s_ldp[(par - 1) * kTileM + r_tile] = ld_sum;
          quad_sync();   // also: every column of s is read out before the next tile's MMAs may overwrite it
          if (par == 0 && valid) {
            const float tot = ld_sum + s_ldp[r_tile] + s_ldp[kTileM + r_tile] + s_ldp[2 * kTileM + r_tile];
            float* op = a.log_det + row;
            if (a.accum == TNF_LD_WRITE) *op = tot;
            else if (a.accum == TNF_LD_ADD) *op = ld_old + tot;
            else *op = ld_old - tot;
          }
        }
      }
    }
  } else {
    // =============================== I/O warps: conditioning half, coalesced ===============================
    const int w2 = warp - (kEpiWarps2 + 2);
    const int row0 = w2 * (kTileM / 2);
    constexpr int LPR = DH / 4 > 32 ? 32 : DH / 4;   // lanes per row piece (16 B each)
    constexpr int RPI = 32 / LPR;                    // rows per warp instruction
    constexpr int PPR = (DH / 4) / LPR;              // 128-float pieces per row
    constexpr int NI = (kTileM / 2) / RPI * PPR;
    constexpr int kBatch = 16;                       // 16-byte loads in flight per lane (8 KB per warp)
    const int rsub = lane / LPR;
    auto load_tile = [&](int64_t it) {
      const int b = (int)(it & 1);
      const int64_t tile = it * (int64_t)gridDim.x + blockIdx.x;
      unsigned char* a1 = sA1 + (size_t)b * sh.a1_bytes();
#pragma unroll 1
      for (int n0 = 0; n0 < NI; n0 += kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int n = n0 + u;
          const int r = row0 + (n / PPR) * RPI + rsub;
          const int hc = 4 * (lane % LPR) + (n % PPR) * 4 * LPR;
          const int64_t grow = tile * kTileM + r;
          v[u] = grow < a.rows ? __ldg(reinterpret_cast<const float4*>(a.z_in + grow * sh.D + sh.c_off + hc))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int n = n0 + u;
          const int r = row0 + (n / PPR) * RPI + rsub;
          const int hc = 4 * (lane % LPR) + (n % PPR) * 4 * LPR;
          const int col = sh.c_off + hc;
          const int64_t grow = tile * kTileM + r;
          const float4 ps = *reinterpret_cast<const float4*>(s_pscale + col);
          const float4 pb = *reinterpret_cast<const float4*>(s_pshift + col);
          float4 x;
          x.x = fmaf(v[u].x, ps.x, pb.x); x.y = fmaf(v[u].y, ps.y, pb.y);
          x.z = fmaf(v[u].z, ps.z, pb.z); x.w = fmaf(v[u].w, ps.w, pb.w);
          *reinterpret_cast<uint2*>(a1 + img_off(r, hc, kTileM)) = make_uint2(pack_bf16(x.x, x.y), pack_bf16(x.z, x.w));
          if (grow < a.rows) *reinterpret_cast<float4*>(a.z_out + grow * sh.D + col) = x;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.a1_ready[b]);
    };
    if (my_tiles > 0) load_tile(0);
    for (int64_t it = 0; it + 1 < my_tiles; ++it) {
      const int64_t t1 = it + 1;
      if (t1 >= 2) mbar_wait(&ct.a1_free[t1 & 1], (uint32_t)(((t1 >> 1) - 1) & 1));   // layer-0 jobs of tile t1-2 done
      load_tile(t1);
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps2) tmem_dealloc(tmem, 512);
}

// ================================================================ two-tile ping-pong kernel (D <= 128)
// Two tiles in flight per CTA.  Epilogue group g (8 warps: TMEM lane quadrant w%4, accumulator chunks of parity
// (w/4)%2) owns tile (2k+g)*grid + cta, the 256 TMEM columns [256g, 256g+256) and its own A1 / activation images.
// A layer's MMAs start when the group's previous epilogue phase is complete (one accumulator per group: the next
// layer would overwrite what the epilogue is still reading), so a single group alternates between the MUFU pipe and
// the tensor pipe - and the other group, half a tile out of phase, fills each pipe in the gaps.  Weights and the
// bias operand images stream through the ring in the static job order of the MMA warp; biases are added by a bias
// MMA (see mma_job), global I/O of the conditioning half is done by two dedicated warps with coalesced accesses.
constexpr int kThreads3 = (kEpiWarps2 + 4) * 32;   // 16 epilogue, MMA, producer, 2 I/O

struct __align__(16) Ctrl3 {
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t a1_ready[2];   // per group, 2 I/O warps: A1 image of the group's next tile written
  uint64_t a1_free[2];    // per group, tcgen05.commit: both layer-0 jobs of the tile have read the A1 image
  uint64_t e_done[2];     // per group, 8 epilogue warps: accumulator drained (and activation image written)
  uint64_t h_ready[2];    // per group, tcgen05.commit: accumulator of the group's current job complete
  uint32_t tmem_base;
  uint32_t pad;
};
// dynamic shared memory:
//   [ring: n_stages x 16 KB][A1 g0][A1 g1][Act g0][Act g1][ones 4 KB][Ctrl3][pre_scale D][pre_shift D][ld partial 2 x 128]
__host__ __device__ inline size_t smem_bytes3(const Shape& sh, int n_stages) {
  return (size_t)n_stages * sh.stage_elems() * 2 + 2 * sh.a1_bytes() + 2 * sh.act_bytes() + kOnesBytes + sizeof(Ctrl3) +
         (size_t)(2 * sh.D + 2 * kTileM) * sizeof(float);
}


template <bool kInverse, int DH, int U_>   // DH = D/2 = d_in = d_out in {32, 64}; U_ = hidden units
__global__ void __launch_bounds__(kThreads3, 1) coupling_tc3_kernel(Args a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const Shape sh(a.D, a.U, a.L, a.upper);
  const int S = a.n_stages;
  unsigned char* ring = smem_raw;
  const uint32_t stage_bytes = (uint32_t)sh.stage_elems() * 2;
  unsigned char* sA1 = ring + (size_t)S * stage_bytes;             // 2 images (group)
  unsigned char* sAct = sA1 + 2 * sh.a1_bytes();                    // 2 images (group)
  unsigned char* sOnes = sAct + 2 * sh.act_bytes();                 // constant A image for the bias MMA
  Ctrl3& ct = *reinterpret_cast<Ctrl3*>(sOnes + kOnesBytes);
  float* s_pscale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(&ct) + sizeof(Ctrl3));
  float* s_pshift = s_pscale + sh.D;
  float* s_ldp = s_pshift + sh.D;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.rows + kTileM - 1) / kTileM;
  const int64_t G = gridDim.x;
  // tiles of group g: (2k + g) * G + blockIdx.x, k = 0 .. cnt[g] - 1
  int64_t cnt[2];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int64_t first = (int64_t)g * G + blockIdx.x;
    cnt[g] = first < n_tiles ? (n_tiles - first + 2 * G - 1) / (2 * G) : 0;
  }
  constexpr int n_chunks = U_ / kChunk;
  const int JT = 2 * (sh.L + 1);     // jobs per tile
  const int shift = sh.L + 1;        // group 1 runs half a tile behind group 0 (other shifts measured no better)

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&ct.w_full[i], 1); mbar_init(&ct.w_empty[i], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&ct.a1_ready[g], 2);
      mbar_init(&ct.a1_free[g], 1);
      mbar_init(&ct.e_done[g], kEpiWarps2 / 2);
      mbar_init(&ct.h_ready[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps2) tmem_alloc(&ct.tmem_base, 512);
  {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x)   // row r: K columns 0 and 1 are 1.0 (bf16 0x3f80)
      reinterpret_cast<uint32_t*>(sOnes)[i] = (i < kTileM * 4 && (i & 3) == 0) ? 0x3f803f80u : 0u;
    fence_async_smem();
    for (int i = threadIdx.x; i < sh.D; i += blockDim.x) {
      s_pscale[i] = a.pre_scale ? a.pre_scale[i] : 1.0f;
      s_pshift[i] = a.pre_shift ? a.pre_shift[i] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ct.tmem_base;
  const bool want_stats = a.stat_partials != nullptr;
  float io_sv1[4] = {0.f, 0.f, 0.f, 0.f}, io_sv2[4] = {0.f, 0.f, 0.f, 0.f};   // I/O warps: sums of columns col..col+3
  // per-warp statistics rows ([2*D] doubles, own columns only) are gathered in the (then dead) activation images
  double* stat_rows = reinterpret_cast<double*>(sAct);
  constexpr int kStatWarps = kEpiWarps2 + 2;

  if (warp == kEpiWarps2 + 1) {
    // =============================== weight producer (one elected lane) ===============================
    // One thread issues a cp.async.bulk every ~330 cycles (profiles/microbench/tma_rate.cu), i.e. ~50 B/cycle with 16 KB
    // stages.  Tried and rejected: a second producer warp (a 21st warp drops every thread to 80 registers), and two
    // diverged lanes of this warp issuing alternate stages (their mbarrier spin loops serialise: slower).
    if (elect_one()) {
      uint32_t slot = 0, phase = 0;
      const int64_t n_steps = cnt[0] * JT + shift;
      // per-group job counters kept incrementally: no 64-bit division in the scheduling loops (a lone warp spends
      // hundreds of cycles on one)
      int jj2[2] = {0, 0};
      const int64_t total[2] = {cnt[0] * JT, cnt[1] * JT};
      int64_t todo[2] = {total[0], total[1]};
      for (int64_t n = 0; n < n_steps; ++n) {
        for (int g = 0; g < 2; ++g) {
          if ((g == 1 && n < shift) || todo[g] == 0) continue;
          const int jj = jj2[g];
          const bool first_job = todo[g] == total[g];
          --todo[g];
          if (++jj2[g] == JT) jj2[g] = 0;
          const int net = jj > sh.L ? 1 : 0, l = jj > sh.L ? jj - (sh.L + 1) : jj;
          (void)first_job; (void)net;
          size_t off = 0;
          for (int i = 0; i < l; ++i) off += (size_t)2 * sh.K_of(i) * sh.J_of(i) * 2;
          const int K = sh.K_of(l), J = sh.J_of(l), N = sh.N_of(l);
          const int ks = sh.stage_k(K, N);
          const uint32_t bytes = (uint32_t)(ks * N * 2);
          const unsigned char* nsrc = a.packed + off + (size_t)net * K * J * 2;
          // the job's bias operand image travels as a stage of its own, ahead of the weights
          {
            mbar_wait(&ct.w_empty[slot], phase ^ 1);
            mbar_arrive_expect_tx(&ct.w_full[slot], (uint32_t)J * 32u);
            bulk_g2s(ring + (size_t)slot * stage_bytes, a.packed + sh.bias_img_off(l, net), (uint32_t)J * 32u, &ct.w_full[slot]);
          }
          if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
          for (int st = 0; st < K / ks; ++st) {
            {
              mbar_wait(&ct.w_empty[slot], phase ^ 1);
              mbar_arrive_expect_tx(&ct.w_full[slot], bytes);
              bulk_g2s(ring + (size_t)slot * stage_bytes, nsrc + (size_t)st * bytes, bytes, &ct.w_full[slot]);
            }
            if (++slot == (uint32_t)S) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == kEpiWarps2) {
    // =============================== MMA issuer (warp-uniform, elected lane issues) ===============================
    const bool leader = elect_one();
    uint32_t slot = 0, phase = 0, e_par = 0, a1_par = 0;   // parities: bit g
    const long long t_all = a.dbg != nullptr ? clock64() : 0;
    long long t_dep = 0, t_w3[3] = {0, 0, 0};
    const bool diag = a.dbg != nullptr;
    const uint32_t wfull0 = smem_u32(&ct.w_full[0]), wempty0 = smem_u32(&ct.w_empty[0]);
    const uint32_t ring16 = smem_u32(ring) >> 4;
    const uint64_t a1_desc = make_desc(smem_u32(sA1), kTileM), act_desc = make_desc(smem_u32(sAct), kTileM);
    const uint32_t a_hi = (uint32_t)(a1_desc >> 32);
    const uint32_t a1_sz16 = (uint32_t)sh.a1_bytes() >> 4, act_sz16 = (uint32_t)sh.act_bytes() >> 4;
    const uint64_t bU_desc = make_desc(0u, U_), bF_desc = make_desc(0u, DH);
    const uint32_t bU_lo = (uint32_t)bU_desc + ring16, bU_hi = (uint32_t)(bU_desc >> 32);
    const uint32_t bF_lo = (uint32_t)bF_desc + ring16, bF_hi = (uint32_t)(bF_desc >> 32);
    const uint32_t ones_lo = (uint32_t)make_desc(smem_u32(sOnes), kTileM);
    const int64_t n_steps = cnt[0] * JT + shift;
    // per-group job counters kept incrementally: no 64-bit division in the scheduling loops (a lone warp spends
    // hundreds of cycles on one)
    int jj2[2] = {0, 0};
    const int64_t total[2] = {cnt[0] * JT, cnt[1] * JT};
    int64_t todo[2] = {total[0], total[1]};
    for (int64_t n = 0; n < n_steps; ++n) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        if ((g == 1 && n < shift) || todo[g] == 0) continue;
        const int jj = jj2[g];
        const bool first_job = todo[g] == total[g];
        --todo[g];
        if (++jj2[g] == JT) jj2[g] = 0;
        const int net = jj > sh.L ? 1 : 0, l = jj > sh.L ? jj - (sh.L + 1) : jj;
        (void)first_job; (void)net;
        const long long c0 = a.dbg != nullptr ? clock64() : 0;
        if (!first_job) {   // the group's previous epilogue phase: accumulator drained, activations written
          mbar_wait(&ct.e_done[g], (e_par >> g) & 1u);
          e_par ^= 1u << g;
        }
        if (jj == 0) {
          mbar_wait(&ct.a1_ready[g], (a1_par >> g) & 1u);
          a1_par ^= 1u << g;
        }
        if (a.dbg != nullptr) t_dep += clock64() - c0;
        const uint32_t d_tmem = tmem + (uint32_t)g * 256u;
        const uint32_t a1_lo = (uint32_t)a1_desc + (uint32_t)g * a1_sz16;
        const uint32_t act_lo = (uint32_t)act_desc + (uint32_t)g * act_sz16;
        if (l == 0)
          mma_job<DH, U_>(d_tmem, a1_lo, a_hi, bU_lo, bU_hi, ones_lo, wfull0, wempty0, (uint32_t)S, slot, phase, leader, diag ? &t_w3[0] : nullptr);
        else if (l < sh.L)
          mma_job<U_, U_>(d_tmem, act_lo, a_hi, bU_lo, bU_hi, ones_lo, wfull0, wempty0, (uint32_t)S, slot, phase, leader, diag ? &t_w3[1] : nullptr);
        else
          mma_job<U_, DH>(d_tmem, act_lo, a_hi, bF_lo, bF_hi, ones_lo, wfull0, wempty0, (uint32_t)S, slot, phase, leader, diag ? &t_w3[2] : nullptr);
        if (leader) {
          tc_commit(&ct.h_ready[g]);
          if (l == 0 && net == 1) tc_commit(&ct.a1_free[g]);
        }
        __syncwarp();
      }
    }
    if (a.dbg != nullptr && blockIdx.x == 0 && leader) {
      a.dbg[2040] = t_dep; a.dbg[2041] = t_w3[0] + t_w3[1] + t_w3[2]; a.dbg[2042] = clock64() - t_all;
      a.dbg[2043] = t_w3[0]; a.dbg[2044] = t_w3[1]; a.dbg[2045] = t_w3[2];
    }
  } else if (warp < kEpiWarps2) {
    // =============================== epilogue warps ===============================
    const int g = warp >> 3, q = warp & 3, par = (warp >> 2) & 1;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int r_tile = q * 32 + lane;
    constexpr int W = DH / 2;                 // final-layer columns per thread (16 or 32)
    const float kLog2e = 1.4426950408889634f;
    const uint32_t hcol = tmem + lane_addr + (uint32_t)g * 256u;
    unsigned char* myAct = sAct + (size_t)g * sh.act_bytes();
    uint32_t h_par = 0;
    float st_y = 0.f, st_y2 = 0.f;            // per-lane column sums of the transformed half (column par*W + lane%W)

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
        for (int e = 0; e < 8; ++e) x[j + e] = __float_as_uint(tanh_fast(__uint_as_float(x[j + e])));
        *reinterpret_cast<uint4*>(dst + (j >> 3) * (kTileM * 16)) =
            make_uint4(pack_bf16(__uint_as_float(x[j]), __uint_as_float(x[j + 1])),
                       pack_bf16(__uint_as_float(x[j + 2]), __uint_as_float(x[j + 3])),
                       pack_bf16(__uint_as_float(x[j + 4]), __uint_as_float(x[j + 5])),
                       pack_bf16(__uint_as_float(x[j + 6]), __uint_as_float(x[j + 7])));
      }
    };
    // end of an epilogue phase: accumulator reads done, activation image visible to the async proxy
    auto phase_done = [&]() {
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.e_done[g]);
    };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + q) : "memory"); };

    for (int64_t k = 0; k < cnt[g]; ++k) {
      const int64_t tile = (2 * k + g) * G + blockIdx.x;
      const int64_t row = tile * kTileM + r_tile;
      const bool valid = row < a.rows;
      const float* zrow = a.z_in + row * sh.D + sh.t_off + par * W;
      TNF_STAMP(100);
      float tv[W];
#pragma unroll
      for (int net = 0; net < 2; ++net) {
#pragma unroll 1
        for (int l = 0; l < sh.L; ++l) {
          TNF_STAMP(200 + net * 10 + l);
          mbar_wait(&ct.h_ready[g], h_par);
          h_par ^= 1;
          tc_fence_after();
          TNF_STAMP(300 + net * 10 + l);
#pragma unroll 1
          for (int c = par; c < n_chunks; c += 2) epi_step(c);
          phase_done();
        }
        // ---- final layer of this net: W columns per thread
        float zin[W];
        float ld_old = 0.f;
        if (net == 1) {   // the transformed half and the old log-det arrive while the final-layer MMAs run
#pragma unroll
          for (int j = 0; j < W; j += 4) {
            const float4 t4 = valid ? __ldg(reinterpret_cast<const float4*>(zrow + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            zin[j] = t4.x; zin[j + 1] = t4.y; zin[j + 2] = t4.z; zin[j + 3] = t4.w;
          }
          if (par == 0 && valid && a.accum != TNF_LD_WRITE) ld_old = a.log_det[row];
        }
        TNF_STAMP(400 + net);
        mbar_wait(&ct.h_ready[g], h_par);
        h_par ^= 1;
        tc_fence_after();
        TNF_STAMP(500 + net);
        uint32_t o[W];
        if (W == 16) tmem_ld16(hcol + (uint32_t)(par * W), reinterpret_cast<uint32_t(&)[16]>(o));
        else tmem_ld32(hcol + (uint32_t)(par * W), reinterpret_cast<uint32_t(&)[32]>(o));
        tc_wait_ld();
        if (net == 0) {
#pragma unroll
          for (int j = 0; j < W; ++j) tv[j] = __uint_as_float(o[j]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ct.e_done[g]);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&ct.e_done[g]);   // accumulator free: the next tile's first job may start
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
          if (valid) {
            float* orow = a.z_out + row * sh.D + sh.t_off + par * W;
#pragma unroll
            for (int j = 0; j < W; j += 4)
              *reinterpret_cast<float4*>(orow + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
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
          if (want_stats) {
            float s1[W], s2[W];
#pragma unroll
            for (int j = 0; j < W; ++j) { s1[j] = valid ? y[j] : 0.f; s2[j] = s1[j] * s1[j]; }
            if (W == 16) {
              st_y += warp_transpose_sum16(reinterpret_cast<float(&)[16]>(s1), lane);
              st_y2 += warp_transpose_sum16(reinterpret_cast<float(&)[16]>(s2), lane);
            } else {
              st_y += warp_transpose_sum(reinterpret_cast<float(&)[32]>(s1), lane);
              st_y2 += warp_transpose_sum(reinterpret_cast<float(&)[32]>(s2), lane);
            }
          }
        }
      }
      TNF_STAMP(600);
    }
#undef TNF_STAMP
    if (want_stats) {
      __syncwarp();
      asm volatile("bar.sync 9, %0;" ::"r"(kEpiWarps2 * 32) : "memory");   // all epilogue warps are past their last phase
      double* rowp = stat_rows + (size_t)warp * 2 * sh.D;
      for (int i = lane; i < 2 * sh.D; i += 32) rowp[i] = 0.0;
      __syncwarp();
      if (lane < W) {
        rowp[sh.t_off + par * W + lane] = (double)st_y;
        rowp[sh.D + sh.t_off + par * W + lane] = (double)st_y2;
      }
    }
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
      const int64_t tile = (2 * k + g) * G + blockIdx.x;
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
          if (grow < a.rows) {
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
      __syncwarp();
      if (lane == 0) mbar_arrive(&ct.a1_ready[g]);
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
    }
    if (want_stats) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1) {
          io_sv1[e] += __shfl_xor_sync(0xffffffffu, io_sv1[e], o);
          io_sv2[e] += __shfl_xor_sync(0xffffffffu, io_sv2[e], o);
        }
      }
    }
  }
  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (want_stats) {
    // gather: one [2*D] row of doubles per CTA = fixed-order sum over the epilogue warps' rows and the I/O warps' sums
    if (warp == kEpiWarps2 + 2 || warp == kEpiWarps2 + 3) {   // I/O warps: expand their sums into rows kEpiWarps2 + w2
      const int w2 = warp - (kEpiWarps2 + 2);
      constexpr int LPR = DH / 4;
      double* rowp = stat_rows + (size_t)(kEpiWarps2 + w2) * 2 * sh.D;
      for (int i = lane; i < 2 * sh.D; i += 32) rowp[i] = 0.0;
      __syncwarp();
      if (lane < LPR) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          rowp[sh.c_off + 4 * lane + e] = (double)io_sv1[e];
          rowp[sh.D + sh.c_off + 4 * lane + e] = (double)io_sv2[e];
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
  if (warp == kEpiWarps2) tmem_dealloc(tmem, 512);
}

// mma_job for a CTA pair: M = 256 MMAs (tcgen05 cta_group::2) issued by the leader CTA; each CTA's ring holds its half of
// the B rows (N/2), so a stage is landed when BOTH this CTA's w_full and the partner's forwarded w_peer completed;
// commits go to both CTAs.

// ---------------------------------------------------------------- diagnostic: one UMMA GEMM
// out[128 x N] = bf16(A[128 x K]) . bf16(W[K x N]) through the same operand images, descriptors and
// TMEM accumulator layout as the fused kernel.  a_in_tmem selects the A source (TMEM or SMEM image).
__global__ void __launch_bounds__(128, 1) selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                           float* __restrict__ out, int K, int N, int a_in_tmem) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* bimg = smem_raw;                       // K x N image, stages of stage_k(K,N) rows
  unsigned char* aimg = smem_raw + (size_t)K * N * 2;   // 128 x K image
  uint64_t* bar = reinterpret_cast<uint64_t*>(aimg + (size_t)kTileM * K * 2);
  uint32_t* tbase = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = (kStageElems / N) < K ? (kStageElems / N) : K;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) tmem_alloc(tbase, 512);
  for (int idx = threadIdx.x; idx < K * N; idx += blockDim.x) {
    const int k = idx / N, n = idx % N;
    unsigned char* dst = bimg + (size_t)(k / ks) * ks * N * 2 + img_off(n, k % ks, N);
    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(W[idx]);
  }
  const int row = warp * 32 + lane;
  for (int k = 0; k < K; ++k)
    *reinterpret_cast<__nv_bfloat16*>(aimg + img_off(row, k, kTileM)) = __float2bfloat16_rn(A[row * K + k]);
  fence_async_smem();   // generic-proxy smem writes -> async proxy (UMMA)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tbase;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  const uint32_t a_col = 256;
  if (a_in_tmem) {
    for (int k = 0; k < K; k += 16) {
      uint32_t p[8];
      for (int j = 0; j < 8; ++j) p[j] = pack_bf16(A[row * K + k + 2 * j], A[row * K + k + 2 * j + 1]);
      tmem_st8(tmem + lane_addr + a_col + (uint32_t)(k / 2), p);
    }
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(N);
    for (int k = 0; k < K; k += 16) {
      const uint32_t b_addr = smem_u32(bimg + (size_t)(k / ks) * ks * N * 2) + (uint32_t)((k % ks) >> 3) * (uint32_t)N * 16u;
      const uint64_t bdesc = make_desc(b_addr, N);
      if (a_in_tmem) {
        umma_ts(tmem, tmem + a_col + (uint32_t)(k >> 1), bdesc, idesc, k > 0 ? 1u : 0u);
      } else {
        const uint64_t adesc = make_desc(smem_u32(aimg) + (uint32_t)(k >> 3) * (kTileM * 16u), kTileM);
        umma_ss(tmem, adesc, bdesc, idesc, k > 0 ? 1u : 0u);
      }
    }
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t o[16];
    tmem_ld16(tmem + lane_addr + (uint32_t)n0, o);
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

int tnf_tc_supported(int D, int U, int L, int precision) {
  if (precision == TNF_TC_FP32) return tc::shape_supported6(D, U, L) ? 1 : 0;
  return precision == TNF_TC_BF16 && tc::shape_supported(D, U, L) ? 1 : 0;
}

// the D = 256 bf16 layer runs on the TMEM-resident single-tile kernel (coupling_tc6.cu, unsplit) with ITS packed format
static bool bf16_on_tc6(int D, int U, int L) { return D == 256 && tc::shape_supported6(D, U, L); }

size_t tnf_tc_packed_bytes(int D, int U, int L, int precision) {
  if (precision == TNF_TC_FP32) return tc::shape_supported6(D, U, L) ? tc::packed_bytes6(D, U, L, 1) : 0;
  if (precision == TNF_TC_BF16 && bf16_on_tc6(D, U, L)) return tc::packed_bytes6(D, U, L, 0);
  if (!tc::shape_supported(D, U, L) || precision != TNF_TC_BF16) return 0;
  return (size_t)tc::Shape(D, U, L, 1).packed_bytes();
}

int tnf_tc_pack(const float* params, void* packed, int D, int U, int L, int transform_upper, int precision,
                tnf_stream_t stream) {
  TNF_REQUIRE(tnf_tc_supported(D, U, L, precision), TNF_ERR_UNSUPPORTED,
              "tnf_tc_pack: shape D=%d U=%d L=%d not supported at precision %d", D, U, L, precision);
  TNF_REQUIRE(params && packed, TNF_ERR_ARG, "tnf_tc_pack: null pointer");
  TNF_REQUIRE(((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN, "tnf_tc_pack: packed buffer must be 16-byte aligned");
  if (precision == TNF_TC_FP32 || bf16_on_tc6(D, U, L)) {
    tc::pack6_launch(params, packed, D, U, L, transform_upper != 0, precision == TNF_TC_FP32, (cudaStream_t)stream);
    return check_launch("tnf_tc_pack");
  }
  tc::Shape sh(D, U, L, transform_upper != 0);
  tc::pack_kernel<<<num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(params, (unsigned char*)packed, sh);
  return check_launch("tnf_tc_pack");
}

int tnf_coupling_tc(const float* z_in, float* z_out, float* log_det, const void* packed, int64_t rows, int D, int U,
                    int L, int transform_upper, int direction, int accum, const float* pre_scale,
                    const float* pre_shift, double* col_stats, void* stats_workspace, int precision, int variant,
                    void* debug, tnf_stream_t stream) {
  return tnf::coupling_tc_impl(z_in, z_out, log_det, packed, rows, D, U, L, transform_upper, direction, accum, pre_scale, pre_shift,
                               col_stats, stats_workspace, precision, variant, debug, nullptr, nullptr, stream, nullptr);
}

}  // extern "C"

namespace tnf {

static bool pairs5_default(int D, int U, int L) {
  return tc::shape_supported2(D, U, L) && tc::shape_supported5(D, U, L) && tc::smem_bytes5(tc::Shape(D, U, L, 1), 4) <= 227 * 1024;
}

bool tc_stats_fusable(int D, int U, int L, int precision) {
  if (!tnf_tc_supported(D, U, L, precision)) return false;
  return precision == TNF_TC_FP32 || bf16_on_tc6(D, U, L) || D <= 128;
}

bool tc_lp_fusable(int D, int U, int L, int precision) {
  if (!tnf_tc_supported(D, U, L, precision)) return false;
  return precision == TNF_TC_FP32 || bf16_on_tc6(D, U, L) || pairs5_default(D, U, L);
}

int coupling_tc_impl(const float* z_in, float* z_out, float* log_det, const void* packed, int64_t rows, int D, int U,
                     int L, int transform_upper, int direction, int accum, const float* pre_scale,
                     const float* pre_shift, double* col_stats, void* stats_workspace, int precision, int variant,
                     void* debug, float* out_lp, const float* lp_scal, tnf_stream_t stream, void* ev_after_kernel) {
  TNF_REQUIRE(out_lp == nullptr || (direction == TNF_INVERSE && accum == TNF_LD_ADD && (variant & 15) == 0 &&
                                    tc_lp_fusable(D, U, L, precision)),
              TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: the fused base density needs the inverse direction, TNF_LD_ADD and the default kernel");
  TNF_REQUIRE(tnf_tc_supported(D, U, L, precision), TNF_ERR_UNSUPPORTED,
              "tnf_coupling_tc: shape D=%d U=%d L=%d not supported at precision %d", D, U, L, precision);
  if (precision == TNF_TC_FP32 || bf16_on_tc6(D, U, L)) {   // TMEM-resident single-tile kernel
    const int split6 = precision == TNF_TC_FP32;
    TNF_REQUIRE((variant & 15) == 0, TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: this shape / precision has one kernel (variant 0)");
    TNF_REQUIRE(rows >= 0, TNF_ERR_ARG, "tnf_coupling_tc: rows < 0");
    if (rows == 0) return 0;
    TNF_REQUIRE(z_in && (z_out || out_lp) && log_det && packed, TNF_ERR_ARG, "tnf_coupling_tc: null pointer");
    TNF_REQUIRE((((uintptr_t)z_in | (uintptr_t)z_out) & 31) == 0 && ((uintptr_t)packed & 15) == 0, TNF_ERR_ALIGN,
                "tnf_coupling_tc: z must be 32-byte aligned (256-bit accesses), the packed weights 16-byte aligned");
    TNF_REQUIRE(col_stats == nullptr || (stats_workspace != nullptr && direction == TNF_FORWARD), TNF_ERR_ARG,
                "tnf_coupling_tc: this kernel takes fused column statistics in the sample direction, with a workspace");
    const int64_t n_super6 = ((rows + tc::kTileM - 1) / tc::kTileM + 1) / 2;
    const int64_t max_pairs6 = num_sms() / 2;
    const int grid6 = 2 * (int)(n_super6 < max_pairs6 ? n_super6 : max_pairs6);
    int ns = 10;
    while (ns > 2 && tc::smem_bytes6(D, U, L, split6, ns) > 227 * 1024) --ns;
    const size_t smem6 = tc::smem_bytes6(D, U, L, split6, ns);
    TNF_REQUIRE(ns >= 4, TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: only %d weight stages fit (a job needs up to 4)", ns);
    TNF_REQUIRE(smem6 <= 227 * 1024, TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: shape needs %zu B shared memory", smem6);
    tc::Args a6{z_in, z_out, log_det, (const unsigned char*)packed, pre_scale, pre_shift, rows,
                D, U, L, transform_upper != 0, direction == TNF_INVERSE, accum, ns, 2, variant >> 8,
                col_stats ? (double*)stats_workspace : nullptr, (long long*)debug, out_lp, lp_scal};
    cudaError_t e6 = (cudaError_t)tc::launch_tc6(a6, grid6, split6, smem6, (cudaStream_t)stream);
    if (e6 != cudaSuccess) {
      set_error("tnf_coupling_tc: cudaFuncSetAttribute(%zu B smem): %s", smem6, cudaGetErrorString(e6));
      return (int)e6;
    }
    int rc6 = check_launch("tnf_coupling_tc");
    if (ev_after_kernel) cudaEventRecord((cudaEvent_t)ev_after_kernel, (cudaStream_t)stream);
    if (rc6 || col_stats == nullptr) return rc6;
    return colstats_reduce_launch((const double*)stats_workspace, grid6, D, col_stats, (double)rows, (cudaStream_t)stream);
  }
  const int g_tc_variant = variant & 15, g_tc_groups = (variant & 16) ? 1 : 2;   // per-call diagnostics, no global state
  long long* const g_tc_debug = (long long*)debug;
  TNF_REQUIRE(rows >= 0, TNF_ERR_ARG, "tnf_coupling_tc: rows < 0");
  if (rows == 0) return 0;
  TNF_REQUIRE(z_in && (z_out || out_lp) && log_det && packed, TNF_ERR_ARG, "tnf_coupling_tc: null pointer");
  TNF_REQUIRE((((uintptr_t)z_in | (uintptr_t)z_out | (uintptr_t)packed) & 15) == 0, TNF_ERR_ALIGN,
              "tnf_coupling_tc: z and packed weights must be 16-byte aligned");
  TNF_REQUIRE(col_stats == nullptr || (D <= 128 && stats_workspace != nullptr), TNF_ERR_UNSUPPORTED,
              "tnf_coupling_tc: fused column statistics need D <= 128 and a workspace");
  TNF_REQUIRE(col_stats == nullptr || D == 64 || (variant & 15) != 1, TNF_ERR_UNSUPPORTED,
              "tnf_coupling_tc: the tile ping-pong kernel accumulates column statistics only at D = 64");
  tc::Shape sh(D, U, L, transform_upper != 0);
  const int64_t n_tiles = (rows + tc::kTileM - 1) / tc::kTileM;
  int grid = (int)(n_tiles < num_sms() ? n_tiles : num_sms());
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  const bool inv = direction == TNF_INVERSE;
  // kernel choice: D <= 128 -> two-tile kernel; D = 256 -> single-tile pipelined kernel; variant 1 -> the first kernel
  const bool pingpong2 = tc::shape_supported2(D, U, L) && g_tc_variant != 1;
  const bool pairs = pingpong2 && g_tc_variant != 2;      // clusters of two CTAs sharing the weight operands
  // N-half / TMEM-fed kernel: the default where it applies (L = 2, U >= 128); variant 4 forces the older pair kernel,
  // variant 3 runs it without the FMA-pipe tanh share (then bit-identical to variant 4)
  const bool pairs5 = pairs && g_tc_variant != 4 && tc::shape_supported5(D, U, L) &&
                      tc::smem_bytes5(tc::Shape(D, U, L, 1), 4) <= 227 * 1024;   // its weight ring is 4 stages, fixed
  const bool pipelined = D == 256 && g_tc_variant != 1;
  if (pairs) {   // one 256-row super tile per CTA pair and group; the grid is a whole number of pairs
    const int64_t n_super = (n_tiles + 1) / 2;
    const int64_t max_pairs = num_sms() / 2;
    grid = 2 * (int)(n_super < max_pairs ? n_super : max_pairs);
  }
  // as many 16 KB weight stages as fit next to the activation images (227 KB per CTA)
  int n_stages = tc::kMaxStages;
  size_t smem;
  if (pairs5) {
    n_stages = 4;
    smem = tc::smem_bytes5(sh, n_stages);
  } else if (pairs) {
    while (n_stages > 2 && tc::smem_bytes4(sh, n_stages) > 227 * 1024) --n_stages;
    smem = tc::smem_bytes4(sh, n_stages);
  } else if (pingpong2) {
    while (n_stages > 2 && tc::smem_bytes3(sh, n_stages) > 227 * 1024) --n_stages;
    smem = tc::smem_bytes3(sh, n_stages);
  } else if (pipelined) {
    while (n_stages > 2 && tc::smem_bytes2(sh, n_stages) > 227 * 1024) --n_stages;
    smem = tc::smem_bytes2(sh, n_stages);
  } else {
    while (n_stages > 2 && tc::smem_bytes(sh, n_stages) > 227 * 1024) --n_stages;
    smem = tc::smem_bytes(sh, n_stages);
  }
  TNF_REQUIRE(smem <= 227 * 1024, TNF_ERR_UNSUPPORTED, "tnf_coupling_tc: shape needs %zu B shared memory", smem);
  tc::Args a{z_in, z_out, log_det, (const unsigned char*)packed, pre_scale, pre_shift, rows,
             D, U, L, transform_upper != 0, direction == TNF_INVERSE, accum, n_stages, g_tc_groups,
             (variant >> 8) | ((variant & 15) == 3 ? 0x100 : 0),   // tune: bits 0-3 of variant >> 8 = diagnostic knob
             col_stats ? (double*)stats_workspace : nullptr, g_tc_debug, out_lp, lp_scal};
#define TNF_TC_LAUNCH(KERNEL, THREADS, INV, DHV)                                                              \
  do {                                                                                                        \
    e = cudaFuncSetAttribute(tc::KERNEL<INV, DHV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    if (e == cudaSuccess) tc::KERNEL<INV, DHV><<<grid, THREADS, smem, st>>>(a);                               \
  } while (0)
#define TNF_TCU_LAUNCH(KERNEL, THREADS, INV, DHV, UV)                                                         \
  do {                                                                                                        \
    e = cudaFuncSetAttribute(tc::KERNEL<INV, DHV, UV>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                             (int)smem);                                                                      \
    if (e == cudaSuccess) tc::KERNEL<INV, DHV, UV><<<grid, THREADS, smem, st>>>(a);                           \
  } while (0)
#define TNF_TCU(KERNEL, THREADS, INV, DHV)                                                                    \
  do {                                                                                                        \
    if (U == 256) TNF_TCU_LAUNCH(KERNEL, THREADS, INV, DHV, 256);                                             \
    else if (U == 128) TNF_TCU_LAUNCH(KERNEL, THREADS, INV, DHV, 128);                                        \
    else TNF_TCU_LAUNCH(KERNEL, THREADS, INV, DHV, 64);                                                       \
  } while (0)
  int stat_blocks = grid * tc::kEpiWarps;
  if (pairs) {
    stat_blocks = grid;
    // coupling_tc5 moves the transformed half with 256-bit accesses (rows are 256 / 512 bytes: only the base matters)
    TNF_REQUIRE(!pairs5 || (((uintptr_t)z_in | (uintptr_t)z_out) & 31) == 0, TNF_ERR_ALIGN,
                "tnf_coupling_tc: z must be 32-byte aligned for this kernel");
    e = (cudaError_t)(pairs5 ? tc::launch_tc5(a, grid, smem, st) : tc::launch_tc4(a, grid, smem, st));
  } else if (pingpong2) {
    stat_blocks = grid;
    if (D == 64) { if (inv) TNF_TCU(coupling_tc3_kernel, tc::kThreads3, true, 32); else TNF_TCU(coupling_tc3_kernel, tc::kThreads3, false, 32); }
    else { if (inv) TNF_TCU(coupling_tc3_kernel, tc::kThreads3, true, 64); else TNF_TCU(coupling_tc3_kernel, tc::kThreads3, false, 64); }
  } else if (pipelined) {
    if (inv) TNF_TCU(coupling_tc2_kernel, tc::kThreads2, true, 128); else TNF_TCU(coupling_tc2_kernel, tc::kThreads2, false, 128);
  } else if (D == 64) { if (inv) TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, true, 32); else TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, false, 32); }
  else if (D == 128) { if (inv) TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, true, 64); else TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, false, 64); }
  else { if (inv) TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, true, 128); else TNF_TC_LAUNCH(coupling_tc_kernel, tc::kThreads, false, 128); }
#undef TNF_TCU
#undef TNF_TCU_LAUNCH
#undef TNF_TC_LAUNCH
  if (e != cudaSuccess) {
    set_error("tnf_coupling_tc: cudaFuncSetAttribute(%zu B smem): %s", smem, cudaGetErrorString(e));
    return (int)e;
  }
  int rc = check_launch("tnf_coupling_tc");
  if (ev_after_kernel) cudaEventRecord((cudaEvent_t)ev_after_kernel, st);   // the coupling kernel alone, not the statistics reduce
  if (rc || col_stats == nullptr) return rc;
  return colstats_reduce_launch((const double*)stats_workspace, stat_blocks, D, col_stats, (double)rows, st);
}

}  // namespace tnf

extern "C" {

int tnf_tc_selftest_gemm(const float* A, const float* W, float* out, int K, int N, int a_in_tmem,
                         tnf_stream_t stream) {
  TNF_REQUIRE(A && W && out, TNF_ERR_ARG, "tnf_tc_selftest_gemm: null pointer");
  TNF_REQUIRE(K >= 16 && K <= 256 && K % 16 == 0 && N >= 16 && N <= 256 && N % 16 == 0, TNF_ERR_ARG,
              "tnf_tc_selftest_gemm: need 16 <= K,N <= 256, multiples of 16");
  const size_t smem = (size_t)K * N * 2 + (size_t)tc::kTileM * K * 2 + 64;
  cudaError_t e = cudaFuncSetAttribute(tc::selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("tnf_tc_selftest_gemm: %s", cudaGetErrorString(e));
    return (int)e;
  }
  tc::selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, W, out, K, N, a_in_tmem);
  return check_launch("tnf_tc_selftest_gemm");
}

}  // extern "C"
