// RealNVP coupling layer, exact-precision CUDA-core path (float / double).
//
// Replaces RealNVP.forward_and_log_det / inverse_and_log_det / _t_s_layer
// (reference torch_nf/bijectors.py:145-242) for every shape the reference
// accepts: any D >= 2 (odd D included), 1 <= L <= 5, U <= 1000, shared weights
// (regime A) or one weight row per m (regime B, the conditional case).
//
// One CTA owns a tile of RB consecutive samples of one parameter row m.
// The tile's activations for both conditioner nets live in shared memory; a
// thread owns one output unit (column) of a layer and keeps RT rows of
// accumulators in registers, so each weight is read once per RT rows, coalesced
// across the CTA (weights are (K, J) row-major: consecutive threads read
// consecutive j).  In regime B (N = 1) this is a pure stream of the parameter
// row at full coalescing: the HBM roofline of that regime.
//
// The tensor-core path for large shared-weight layers is coupling_tc.cu.
#include "common.cuh"

namespace tnf {

struct CouplingShape {
  int D, U, L, upper, maf;
  int h, d_in, d_out, c_off, t_off, W;  // c_off: first conditioning column; t_off: first transformed column
  // maf != 0: masked autoregressive layer (reference bijectors.py:742-796): every column conditions and is
  // transformed (d_in = d_out = D), the nets have no biases and their weights are multiplied by fixed masks.
  __host__ __device__ CouplingShape(int D_, int U_, int L_, int upper_, int maf_ = 0)
      : D(D_), U(U_), L(L_), upper(upper_), maf(maf_) {
    h = D / 2;
    if (maf) { d_in = D; d_out = D; c_off = 0; t_off = 0; }
    else if (upper) { d_in = h; d_out = D - h; c_off = 0; t_off = h; }
    else { d_in = D - h; d_out = h; c_off = h; t_off = 0; }
    W = U > d_out ? U : d_out;
  }
  __host__ __device__ int64_t num_params() const {
    return 2 * ((int64_t)d_in * U + (int64_t)d_out * U + d_out + U + (int64_t)(L - 1) * (U + 1) * U);
  }
};

// 4- / 8-byte asynchronous global -> shared copies: the whole parameter row of a context is requested at once (one
// memory latency) instead of ~K dependent loads per layer and thread
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem, const T* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem), "n"((int)sizeof(T)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T>
struct CouplingArgs {
  const T* z_in; T* z_out; T* log_det; const T* params;
  int64_t pstride, M, N;
  int D, U, L, upper, inverse, accum, RB;
  const float* mask;   // MAF: flat 0/1 mask in the parameter-row layout (NULL = RealNVP)
  int passes;          // MAF forward: D-1 fixed-point passes (reference bijectors.py:751-756); else 1
  int stage_params;    // > 0: per-row weights (regime B), one CTA per row: the row's `stage_params` parameters are copied to
                       // shared memory with cp.async first (one latency for the whole row)
  int64_t n_vblocks;   // tiles in total; with WPB warps per CTA the last CTA may have idle warps
  int vblock_smem;     // shared-memory bytes of one tile
};

// y[r][j] = act(sum_k in[r][k] W[k][j] + b[j]) for both nets; rows 0..RB-1 (RB % RT == 0)
template <typename T, int RT>
__device__ __forceinline__ void mlp_layer(const T* __restrict__ in_t, const T* __restrict__ in_s, int in_stride,
                                          int K, int J, const T* __restrict__ Wt, const T* __restrict__ Ws,
                                          const T* __restrict__ bt, const T* __restrict__ bs, T* __restrict__ out_t,
                                          T* __restrict__ out_s, int out_stride, int RB, bool act, int vt, int vn,
                                          const float* __restrict__ mk = nullptr) {
  for (int c = vt; c < 2 * J; c += vn) {
    const int net = c >= J;
    const int j = c - net * J;
    const T* W = net ? Ws : Wt;
    const T* in = net ? in_s : in_t;
    T* out = net ? out_s : out_t;
    const T bias = bt ? (net ? bs : bt)[j] : T(0);
    for (int r0 = 0; r0 < RB; r0 += RT) {
      T acc[RT];
#pragma unroll
      for (int r = 0; r < RT; ++r) acc[r] = T(0);
      const T* inr = in + (size_t)r0 * in_stride;
#pragma unroll 4
      for (int k = 0; k < K; ++k) {
        T w = W[(size_t)k * J + j];
        if (mk) w *= (T)mk[(size_t)k * J + j];
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] += inr[r * in_stride + k] * w;
      }
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        T v = acc[r] + bias;
        out[(size_t)(r0 + r) * out_stride + j] = act ? t_tanh<T>(v) : v;
      }
    }
  }
}

// WPB > 1 (single-warp tiles, the conditional regime): WPB independent tiles per CTA, one per warp, synchronised with
// __syncwarp - a CTA per 32-thread tile caps the SM at 32 resident warps, and this kernel lives on latency hiding
template <typename T, int RT, int WPB>
__global__ void coupling_generic_kernel(CouplingArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_base[];
  const int vt = WPB > 1 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  const int vn = WPB > 1 ? 32 : (int)blockDim.x;
  const int64_t vb = WPB > 1 ? (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5) : (int64_t)blockIdx.x;
  if (WPB > 1 && vb >= a.n_vblocks) return;      // whole warp: nothing below synchronises across warps
  unsigned char* smem_raw = smem_base + (WPB > 1 ? (size_t)(threadIdx.x >> 5) * a.vblock_smem : 0);
  auto vsync = [&]() { if (WPB > 1) __syncwarp(); else __syncthreads(); };
  const CouplingShape sh(a.D, a.U, a.L, a.upper, a.mask != nullptr);
  const int RB = a.RB, D = a.D, U = a.U;
  T* bufA = reinterpret_cast<T*>(smem_raw);    // [2][RB][W]
  T* bufB = bufA + (size_t)2 * RB * sh.W;      // [2][RB][W]
  T* zt = bufB + (size_t)2 * RB * sh.W;        // [RB][D]
  T* ut = zt + (size_t)RB * D;                 // [RB][D]  MAF forward: the fixed input u
  T* prow = ut + (size_t)RB * D;               // [stage_params]  this row's parameters (regime B)

  const int64_t tiles_per_m = (a.N + RB - 1) / RB;
  const int64_t m = vb / tiles_per_m;
  const int64_t n0 = (vb % tiles_per_m) * RB;
  const int rows = (int)((a.N - n0) < RB ? (a.N - n0) : RB);
  const T* p0 = a.params + m * a.pstride;
  if (a.stage_params > 0) {
    for (int i = vt; i < a.stage_params; i += vn) cp_async_elem(prow + i, p0 + i);
    cp_async_wait_all();
    p0 = prow;      // visible to the other threads after the barrier that follows the z tile load
  }
  const T* zin = a.z_in + (m * a.N + n0) * D;
  T* zout = a.z_out + (m * a.N + n0) * D;

  for (int e = vt; e < RB * D; e += vn) {
    const T v = e < rows * D ? zin[e] : T(0);
    zt[e] = v;
    if (sh.maf) ut[e] = v;
  }
  vsync();

  const T* tt = nullptr;
  const T* ss = nullptr;
  for (int pass = 0; pass < a.passes; ++pass) {
    // layer 0: d_in -> U on the conditioning columns, both nets read the same input
    const T* p = p0;
    const float* mk = a.mask;
    const T* src_t = zt + sh.c_off;
    const T* src_s = zt + sh.c_off;
    int src_stride = D, K = sh.d_in;
    T* cur = bufA;
    T* nxt = bufB;
    for (int l = 0; l <= a.L; ++l) {
      const int J = (l == a.L) ? sh.d_out : U;
      const T* Wt = p;
      const T* Ws = p + (size_t)K * J;
      const T* bt = sh.maf ? nullptr : Ws + (size_t)K * J;
      const T* bs = sh.maf ? nullptr : bt + J;
      mlp_layer<T, RT>(src_t, src_s, src_stride, K, J, Wt, Ws, bt, bs, cur, cur + (size_t)RB * sh.W, sh.W, RB,
                       l < a.L, vt, vn, mk);
      vsync();
      p = sh.maf ? Ws + (size_t)K * J : bs + J;
      if (mk) mk += (size_t)2 * K * J;
      src_t = cur;
      src_s = cur + (size_t)RB * sh.W;
      src_stride = sh.W;
      K = J;
      T* tmp = cur; cur = nxt; nxt = tmp;
    }
    // src_t / src_s now hold t and s ([RB][W], first d_out columns)
    tt = src_t;
    ss = src_s;
    for (int e = vt; e < rows * sh.d_out; e += vn) {
      const int r = e / sh.d_out, j = e - r * sh.d_out;
      const T t = tt[(size_t)r * sh.W + j], sv = ss[(size_t)r * sh.W + j];
      T* zp = &zt[r * D + sh.t_off + j];
      const T z2 = (sh.maf && !a.inverse) ? ut[r * D + j] : *zp;
      *zp = a.inverse ? (z2 - t) / t_exp<T>(sv) : t + z2 * t_exp<T>(sv);
    }
    vsync();
  }
  for (int r = vt; r < rows; r += vn) {
    T ld = T(0);
    for (int j = 0; j < sh.d_out; ++j) ld += ss[(size_t)r * sh.W + j];
    T* o = a.log_det + m * a.N + n0 + r;
    if (a.accum == TNF_LD_WRITE) *o = ld;
    else if (a.accum == TNF_LD_ADD) *o += ld;
    else *o -= ld;
  }
  for (int e = vt; e < rows * D; e += vn) zout[e] = zt[e];
}

// ------------------------------------------------------------------ backward
template <typename T>
struct CouplingBwdArgs {
  const T* z_in; const T* params; const T* g_y; const T* g_ld; T* g_z; T* g_params;
  int64_t pstride, gstride, M, N;
  int D, U, L, upper, inverse, RB, atomic_params;
  const float* mask;   // MAF (inverse direction only); NULL = RealNVP
  int stage_params;    // > 0 (regime B, one CTA per row): the row's parameters are copied to shared memory with cp.async
  int64_t n_vblocks;   // tiles in total
  int vblock_smem;     // shared-memory bytes of one tile
};

// mode 0: accumulate (read-modify-write), 1: atomic accumulate (several CTAs share the row), 2: plain store (this CTA
// is the only writer of the row and writes every element exactly once: no zero-fill by the caller, no load)
template <typename T>
__device__ __forceinline__ void grad_add(T* addr, T v, int mode) {
  if (mode == 1) atomicAdd(addr, v);
  else if (mode == 2) *addr = v;
  else *addr += v;
}

// Shared-memory plan (all [RB][.] row-major):
//   zt   [RB][D]            layer input tile (later reused to assemble g_z)
//   act  [2][L][RB][U]      post-tanh activations of every hidden layer, both nets
//   dl   [2][2][RB][W]      delta ping-pong, both nets
template <typename T, int RT, int WPB>
__global__ void coupling_generic_bwd_kernel(CouplingBwdArgs<T> a) {
  extern __shared__ __align__(16) unsigned char smem_base[];
  const int vt = WPB > 1 ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  const int vn = WPB > 1 ? 32 : (int)blockDim.x;
  const int64_t vb = WPB > 1 ? (int64_t)blockIdx.x * WPB + (threadIdx.x >> 5) : (int64_t)blockIdx.x;
  if (WPB > 1 && vb >= a.n_vblocks) return;
  unsigned char* smem_raw = smem_base + (WPB > 1 ? (size_t)(threadIdx.x >> 5) * a.vblock_smem : 0);
  auto vsync = [&]() { if (WPB > 1) __syncwarp(); else __syncthreads(); };
  const CouplingShape sh(a.D, a.U, a.L, a.upper, a.mask != nullptr);
  const int RB = a.RB, D = a.D, U = a.U, L = a.L, W = sh.W;
  const int nbias = sh.maf ? 0 : 1;   // MAF nets have no biases
  T* zt = reinterpret_cast<T*>(smem_raw);
  T* act = zt + (size_t)RB * D;
  T* dl = act + (size_t)2 * L * RB * U;
  T* gz1 = dl + (size_t)4 * RB * W;  // [RB][d_in]
  T* gdir = gz1 + (size_t)RB * sh.d_in;  // [RB][D]  MAF: direct part of g_z (zt must stay intact: it is the layer-0 input)
  T* prow = gdir + (size_t)RB * D;       // [stage_params] this row's parameters (regime B)
  auto ACT = [&](int net, int l) { return act + ((size_t)(net * L + l) * RB) * U; };  // output of layer l
  auto DL = [&](int buf, int net) { return dl + ((size_t)(buf * 2 + net) * RB) * W; };

  const int64_t tiles_per_m = (a.N + RB - 1) / RB;
  const int64_t m = vb / tiles_per_m;
  const int64_t n0 = (vb % tiles_per_m) * RB;
  const int rows = (int)((a.N - n0) < RB ? (a.N - n0) : RB);
  const T* p0 = a.params + m * a.pstride;
  T* gp0 = a.g_params + (a.gstride ? m * a.gstride : 0);
  if (a.stage_params > 0) {   // read twice below (recompute, input deltas): one asynchronous copy of the whole row
    for (int i = vt; i < a.stage_params; i += vn) cp_async_elem(prow + i, p0 + i);
    cp_async_wait_all();
    p0 = prow;                // visible to the other threads after the barrier that follows the z tile load
  }
  const T* zin = a.z_in + (m * a.N + n0) * D;
  const T* gy = a.g_y ? a.g_y + (m * a.N + n0) * D : nullptr;
  const T* gld = a.g_ld ? a.g_ld + m * a.N + n0 : nullptr;
  T* gz = a.g_z + (m * a.N + n0) * D;

  for (int e = vt; e < RB * D; e += vn) zt[e] = e < rows * D ? zin[e] : T(0);
  vsync();

  // ---- recompute the conditioner, keeping every hidden activation
  {
    const T* p = p0;
    const float* mk = a.mask;
    const T* src_t = zt + sh.c_off;
    const T* src_s = src_t;
    int src_stride = D, K = sh.d_in;
    for (int l = 0; l < L; ++l) {
      const T* Wt = p; const T* Ws = p + (size_t)K * U;
      const T* bt = nbias ? Ws + (size_t)K * U : nullptr; const T* bs = nbias ? bt + U : nullptr;
      mlp_layer<T, RT>(src_t, src_s, src_stride, K, U, Wt, Ws, bt, bs, ACT(0, l), ACT(1, l), U, RB, true, vt, vn, mk);
      vsync();
      p = Ws + (size_t)K * U + (size_t)nbias * 2 * U;
      if (mk) mk += (size_t)2 * K * U;
      src_t = ACT(0, l); src_s = ACT(1, l); src_stride = U; K = U;
    }
    const T* Wt = p; const T* Ws = p + (size_t)K * sh.d_out;
    const T* bt = nbias ? Ws + (size_t)K * sh.d_out : nullptr; const T* bs = nbias ? bt + sh.d_out : nullptr;
    mlp_layer<T, RT>(src_t, src_s, src_stride, K, sh.d_out, Wt, Ws, bt, bs, DL(1, 0), DL(1, 1), W, RB, false, vt, vn, mk);
    vsync();
  }
  // ---- output deltas: DL(0,net) <- d loss / d (t, s); g_z2 into zt
  {
    const T* tt = DL(1, 0); const T* ss = DL(1, 1);
    T* dt = DL(0, 0); T* ds = DL(0, 1);
    for (int e = vt; e < RB * sh.d_out; e += vn) {
      const int r = e / sh.d_out, j = e - r * sh.d_out;
      T g_t = T(0), g_s = T(0);
      if (r < rows) {
        const T t = tt[(size_t)r * W + j], s = ss[(size_t)r * W + j];
        const T z2 = zt[r * D + sh.t_off + j];
        const T g2 = gy ? gy[r * D + sh.t_off + j] : T(0);
        const T gl = gld ? gld[r] : T(0);
        const T es = t_exp<T>(s);
        T* direct = sh.maf ? &gdir[r * D + j] : &zt[r * D + sh.t_off + j];
        if (!a.inverse) {
          g_t = g2; g_s = g2 * z2 * es + gl;
          *direct = g2 * es;
        } else {
          const T y2 = (z2 - t) / es;
          g_t = -g2 / es; g_s = -g2 * y2 + gl;
          *direct = g2 / es;
        }
      }
      dt[(size_t)r * W + j] = g_t;
      ds[(size_t)r * W + j] = g_s;
    }
    vsync();
  }
  // ---- walk the layers backwards
  int cur = 0;
  // parameter offsets of each layer
  for (int l = L; l >= 0; --l) {
    const int K = (l == 0) ? sh.d_in : U;
    const int J = (l == L) ? sh.d_out : U;
    // offset of layer l inside the parameter row (MAF rows carry no biases)
    int64_t off = 0;
    if (l >= 1) off += 2 * ((int64_t)sh.d_in * U + nbias * U);
    if (l >= 2) off += (int64_t)(l - 1) * 2 * ((int64_t)U * U + nbias * U);
    const T* Wt = p0 + off; const T* Ws = Wt + (size_t)K * J;
    const float* mk = a.mask ? a.mask + off : nullptr;
    T* gWt = gp0 + off; T* gWs = gWt + (size_t)K * J; T* gbt = gWs + (size_t)K * J; T* gbs = gbt + J;
    const T* in_t = (l == 0) ? zt + sh.c_off : ACT(0, l - 1);
    const T* in_s = (l == 0) ? zt + sh.c_off : ACT(1, l - 1);
    const int in_stride = (l == 0) ? D : U;
    const T* dt = DL(cur, 0); const T* ds = DL(cur, 1);
    // weight / bias gradients: thread per (k, j), j fastest (coalesced)
    for (int64_t e = vt; e < (int64_t)2 * K * J; e += vn) {
      const int net = e >= (int64_t)K * J;
      const int64_t kj = e - (int64_t)net * K * J;
      const int k = (int)(kj / J), j = (int)(kj - (int64_t)k * J);
      const T* in = net ? in_s : in_t;
      const T* dd = net ? ds : dt;
      T sacc = T(0);
      for (int r = 0; r < rows; ++r) sacc += in[(size_t)r * in_stride + k] * dd[(size_t)r * W + j];
      if (mk) sacc *= (T)mk[kj];
      grad_add<T>((net ? gWs : gWt) + kj, sacc, a.atomic_params);
    }
    if (nbias) {
      for (int c = vt; c < 2 * J; c += vn) {
        const int net = c >= J; const int j = c - net * J;
        const T* dd = net ? ds : dt;
        T sacc = T(0);
        for (int r = 0; r < rows; ++r) sacc += dd[(size_t)r * W + j];
        grad_add<T>((net ? gbs : gbt) + j, sacc, a.atomic_params);
      }
    }
    // input deltas
    if (l > 0) {
      T* nt_ = DL(cur ^ 1, 0); T* ns_ = DL(cur ^ 1, 1);
      for (int c = vt; c < 2 * K; c += vn) {
        const int net = c >= K; const int k = c - net * K;
        const T* Wn = net ? Ws : Wt; const T* dd = net ? ds : dt;
        const T* ain = net ? in_s : in_t;
        T* outp = net ? ns_ : nt_;
        for (int r0 = 0; r0 < RB; r0 += RT) {
          T acc[RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) acc[r] = T(0);
          for (int j = 0; j < J; ++j) {
            T w = Wn[(size_t)k * J + j];
            if (mk) w *= (T)mk[(size_t)k * J + j];
#pragma unroll
            for (int r = 0; r < RT; ++r) acc[r] += dd[(size_t)(r0 + r) * W + j] * w;
          }
#pragma unroll
          for (int r = 0; r < RT; ++r) {
            const T av = ain[(size_t)(r0 + r) * in_stride + k];
            outp[(size_t)(r0 + r) * W + k] = acc[r] * (T(1) - av * av);
          }
        }
      }
    } else {
      for (int k = vt; k < K; k += vn) {
        for (int r0 = 0; r0 < RB; r0 += RT) {
          T acc[RT];
#pragma unroll
          for (int r = 0; r < RT; ++r) acc[r] = T(0);
          for (int j = 0; j < J; ++j) {
            T wt = Wt[(size_t)k * J + j], ws = Ws[(size_t)k * J + j];
            if (mk) { const T mm = (T)mk[(size_t)k * J + j]; wt *= mm; ws *= mm; }
#pragma unroll
            for (int r = 0; r < RT; ++r)
              acc[r] += dt[(size_t)(r0 + r) * W + j] * wt + ds[(size_t)(r0 + r) * W + j] * ws;
          }
#pragma unroll
          for (int r = 0; r < RT; ++r) gz1[(size_t)(r0 + r) * sh.d_in + k] = acc[r];
        }
      }
    }
    vsync();
    cur ^= 1;
  }
  // ---- assemble g_z: conditioning half = g_y1 + MLP input gradient; transformed half is in zt
  for (int e = vt; e < rows * D; e += vn) {
    const int r = e / D, d = e - r * D;
    T v;
    if (sh.maf) v = gdir[e] + gz1[(size_t)r * sh.d_in + d];     // every column is transformed AND conditions
    else if (d >= sh.t_off && d < sh.t_off + sh.d_out) v = zt[e];
    else v = (gy ? gy[e] : T(0)) + gz1[(size_t)r * sh.d_in + (d - sh.c_off)];
    gz[e] = v;
  }
}

template <typename T>
static int pick_rb(int64_t N, size_t bytes_per_row, size_t fixed, size_t budget) {
  int rb = 32;
  while (rb > 1 && (rb / 2 >= N || fixed + (size_t)rb * bytes_per_row > budget)) rb /= 2;
  return rb;
}

static int validate(const char* what, int64_t M, int64_t N, int D, int U, int L) {
  TNF_REQUIRE(M >= 0 && N >= 0, TNF_ERR_ARG, "%s: bad batch M=%lld N=%lld", what, (long long)M, (long long)N);
  TNF_REQUIRE(D >= 2, TNF_ERR_ARG, "%s: D=%d must be >= 2", what, D);
  TNF_REQUIRE(U >= 1 && U <= 1000, TNF_ERR_ARG, "%s: U=%d outside [1,1000]", what, U);
  TNF_REQUIRE(L >= 1 && L <= 5, TNF_ERR_ARG, "%s: L=%d outside [1,5]", what, L);
  return 0;
}

template <typename T>
static int launch_fwd(const void* z_in, void* z_out, void* log_det, const void* params, int64_t pstride, int64_t M,
                      int64_t N, int D, int U, int L, int upper, int direction, int accum, cudaStream_t st,
                      const float* mask = nullptr) {
  CouplingShape sh(D, U, L, upper, mask != nullptr);
  const size_t budget = 200 * 1024;
  size_t per_row = ((size_t)4 * sh.W + 2 * D) * sizeof(T);
  int RB = pick_rb<T>(N, per_row, 0, budget);
  TNF_REQUIRE(per_row * RB <= budget, TNF_ERR_UNSUPPORTED, "tnf_coupling: shape needs %zu B smem", per_row * RB);
  size_t smem = per_row * RB;
  // regime B with one CTA per parameter row and a small net: stage the row in shared memory
  const int64_t np = sh.num_params();
  const int stage = (pstride != 0 && M > 1 && mask == nullptr && (N + RB - 1) / RB == 1 && np * (int64_t)sizeof(T) <= 6 * 1024) ? (int)np : 0;
  smem += (size_t)stage * sizeof(T);
  CouplingArgs<T> a{(const T*)z_in, (T*)z_out, (T*)log_det, (const T*)params, pstride, M, N,
                    D, U, L, upper, direction == TNF_INVERSE, accum, RB, mask,
                    (mask != nullptr && direction != TNF_INVERSE) ? (D - 1 > 0 ? D - 1 : 1) : 1, stage, 0, 0};
  int64_t tiles = M * ((N + RB - 1) / RB);
  TNF_REQUIRE(tiles < (int64_t)1 << 31, TNF_ERR_UNSUPPORTED, "tnf_coupling: too many tiles");
  int nt = 2 * sh.W;
  nt = nt < 32 ? 32 : (nt > 256 ? 256 : (nt + 31) / 32 * 32);
  smem = (smem + 15) & ~(size_t)15;
  a.n_vblocks = tiles; a.vblock_smem = (int)smem;
#define TNF_LAUNCH_FWD(RT)                                                                                    \
  do {                                                                                                        \
    cudaFuncSetAttribute(coupling_generic_kernel<T, RT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    coupling_generic_kernel<T, RT, 1><<<(unsigned)tiles, nt, smem, st>>>(a);                                  \
  } while (0)
  if (nt == 32 && RB == 1 && tiles >= 4096 && 2 * smem <= 64 * 1024) {   // single-warp tiles: two per CTA (64 resident warps per SM)
    cudaFuncSetAttribute(coupling_generic_kernel<T, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));
    coupling_generic_kernel<T, 1, 2><<<(unsigned)((tiles + 1) / 2), 64, 2 * smem, st>>>(a);
    return check_launch("tnf_coupling");
  }
  if (RB >= 8) TNF_LAUNCH_FWD(8);
  else if (RB == 4) TNF_LAUNCH_FWD(4);
  else if (RB == 2) TNF_LAUNCH_FWD(2);
  else TNF_LAUNCH_FWD(1);
#undef TNF_LAUNCH_FWD
  return check_launch("tnf_coupling");
}

template <typename T>
static int launch_bwd(const void* z_in, const void* params, int64_t pstride, const void* g_y, const void* g_ld,
                      void* g_z, void* g_params, int64_t gstride, int64_t M, int64_t N, int D, int U, int L,
                      int upper, int direction, cudaStream_t st, const float* mask = nullptr, bool overwrite = false) {
  CouplingShape sh(D, U, L, upper, mask != nullptr);
  const size_t budget = 200 * 1024;
  size_t per_row = ((size_t)2 * D + (size_t)2 * L * U + (size_t)4 * sh.W + sh.d_in) * sizeof(T);
  int RB = pick_rb<T>(N, per_row, 0, budget);
  TNF_REQUIRE(per_row * RB <= budget, TNF_ERR_UNSUPPORTED, "tnf_coupling_bwd: shape needs %zu B smem", per_row * RB);
  size_t smem = per_row * RB;
  int64_t tiles_per_m = (N + RB - 1) / RB;
  int atomic_params = (gstride == 0 && M * tiles_per_m > 1) || tiles_per_m > 1;
  if (overwrite) {
    TNF_REQUIRE(!atomic_params && gstride != 0, TNF_ERR_UNSUPPORTED,
                "tnf_coupling_bwd_overwrite: needs one parameter row per m and one tile per row (N <= 32)");
    atomic_params = 2;
  }
  // the parameter row staged in shared memory as in the forward (staging the GRADIENT row too was measured slower)
  const int64_t np = sh.num_params();
  const int stage = (pstride != 0 && M > 1 && tiles_per_m == 1 && mask == nullptr && np * (int64_t)sizeof(T) <= 6 * 1024) ? (int)np : 0;
  smem += (size_t)stage * sizeof(T);
  CouplingBwdArgs<T> a{(const T*)z_in, (const T*)params, (const T*)g_y, (const T*)g_ld, (T*)g_z, (T*)g_params,
                       pstride, gstride, M, N, D, U, L, upper, direction == TNF_INVERSE, RB, atomic_params, mask, stage, 0, 0};
  int64_t tiles = M * tiles_per_m;
  TNF_REQUIRE(tiles < (int64_t)1 << 31, TNF_ERR_UNSUPPORTED, "tnf_coupling_bwd: too many tiles");
  int nt = 2 * sh.W;
  nt = nt < 32 ? 32 : (nt > 256 ? 256 : (nt + 31) / 32 * 32);
  smem = (smem + 15) & ~(size_t)15;
  a.n_vblocks = tiles; a.vblock_smem = (int)smem;
#define TNF_LAUNCH_BWD(RT)                                                                                        \
  do {                                                                                                            \
    cudaFuncSetAttribute(coupling_generic_bwd_kernel<T, RT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    coupling_generic_bwd_kernel<T, RT, 1><<<(unsigned)tiles, nt, smem, st>>>(a);                                  \
  } while (0)
  if (nt == 32 && RB == 1 && tiles >= 4096 && 2 * smem <= 64 * 1024) {   // single-warp tiles: two per CTA
    cudaFuncSetAttribute(coupling_generic_bwd_kernel<T, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * smem));
    coupling_generic_bwd_kernel<T, 1, 2><<<(unsigned)((tiles + 1) / 2), 64, 2 * smem, st>>>(a);
    return check_launch("tnf_coupling_bwd");
  }
  if (RB >= 8) TNF_LAUNCH_BWD(8);
  else if (RB == 4) TNF_LAUNCH_BWD(4);
  else if (RB == 2) TNF_LAUNCH_BWD(2);
  else TNF_LAUNCH_BWD(1);
#undef TNF_LAUNCH_BWD
  return check_launch("tnf_coupling_bwd");
}

}  // namespace tnf

using namespace tnf;

extern "C" {

int tnf_coupling(const void* z_in, void* z_out, void* log_det, const void* params, int64_t pstride, int64_t M,
                 int64_t N, int D, int U, int L, int transform_upper, int direction, int accum, int dtype,
                 tnf_stream_t stream) {
  int rc = validate("tnf_coupling", M, N, D, U, L);
  if (rc) return rc;
  if (M == 0 || N == 0) return 0;  // empty batch: nothing to do, pointers may be null
  TNF_REQUIRE(z_in && z_out && log_det && params, TNF_ERR_ARG, "tnf_coupling: null pointer");
  TNF_DISPATCH(dtype, return launch_fwd<T>(z_in, z_out, log_det, params, pstride, M, N, D, U, L, transform_upper != 0,
                                           direction, accum, (cudaStream_t)stream));
  return 0;
}

int tnf_coupling_bwd(const void* z_in, const void* params, int64_t pstride, const void* g_z_out,
                     const void* g_log_det, void* g_z_in, void* g_params, int64_t gstride, int64_t M, int64_t N, int D,
                     int U, int L, int transform_upper, int direction, int dtype, tnf_stream_t stream) {
  int rc = validate("tnf_coupling_bwd", M, N, D, U, L);
  if (rc) return rc;
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && params && g_z_in && g_params, TNF_ERR_ARG, "tnf_coupling_bwd: null pointer");
  TNF_DISPATCH(dtype, return launch_bwd<T>(z_in, params, pstride, g_z_out, g_log_det, g_z_in, g_params, gstride, M, N,
                                           D, U, L, transform_upper != 0, direction, (cudaStream_t)stream));
  return 0;
}

int tnf_coupling_bwd_overwrite(const void* z_in, const void* params, int64_t pstride, const void* g_z_out,
                               const void* g_log_det, void* g_z_in, void* g_params, int64_t gstride, int64_t M, int64_t N,
                               int D, int U, int L, int transform_upper, int direction, int dtype, tnf_stream_t stream) {
  int rc = validate("tnf_coupling_bwd_overwrite", M, N, D, U, L);
  if (rc) return rc;
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && params && g_z_in && g_params, TNF_ERR_ARG, "tnf_coupling_bwd_overwrite: null pointer");
  TNF_DISPATCH(dtype, return launch_bwd<T>(z_in, params, pstride, g_z_out, g_log_det, g_z_in, g_params, gstride, M, N,
                                           D, U, L, transform_upper != 0, direction, (cudaStream_t)stream, nullptr, true));
  return 0;
}

int tnf_maf(const void* z_in, void* z_out, void* log_det, const void* params, int64_t pstride, const float* mask,
            int64_t M, int64_t N, int D, int U, int L, int direction, int accum, int dtype, tnf_stream_t stream) {
  int rc = validate("tnf_maf", M, N, D, U, L);
  if (rc) return rc;
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && z_out && log_det && params && mask, TNF_ERR_ARG, "tnf_maf: null pointer");
  TNF_DISPATCH(dtype, return launch_fwd<T>(z_in, z_out, log_det, params, pstride, M, N, D, U, L, 1, direction, accum,
                                           (cudaStream_t)stream, mask));
  return 0;
}

int tnf_maf_bwd(const void* z_in, const void* params, int64_t pstride, const float* mask, const void* g_z_out,
                const void* g_log_det, void* g_z_in, void* g_params, int64_t gstride, int64_t M, int64_t N, int D,
                int U, int L, int direction, int dtype, tnf_stream_t stream) {
  int rc = validate("tnf_maf_bwd", M, N, D, U, L);
  if (rc) return rc;
  TNF_REQUIRE(direction == TNF_INVERSE, TNF_ERR_UNSUPPORTED,
              "tnf_maf_bwd: only the inverse (log_prob) direction has a backward");
  if (M == 0 || N == 0) return 0;
  TNF_REQUIRE(z_in && params && mask && g_z_in && g_params, TNF_ERR_ARG, "tnf_maf_bwd: null pointer");
  TNF_DISPATCH(dtype, return launch_bwd<T>(z_in, params, pstride, g_z_out, g_log_det, g_z_in, g_params, gstride, M, N,
                                           D, U, L, 1, direction, (cudaStream_t)stream, mask));
  return 0;
}

}  // extern "C"
