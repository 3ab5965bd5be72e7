/*
 * tnf.h -- C ABI of the B200-native torch_nf bijector-chain hot path.
 *
 * The reference (srbittner/torch_nf) is pure Python: it has no FFI layer, its
 * "plugin" boundary is the duck-typed Bijector protocol
 *     bijector(z, params) -> (z, log_det)            torch_nf/bijectors.py:30-63
 * called from NormFlow.forward / NormFlow.inverse_and_log_det
 *                                    torch_nf/density_estimator.py:375-387,395-405.
 * Each entry point below is what a binding for one of those protocol methods
 * calls; the reference method it replaces is cited on the declaration.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by
 *     the caller; the library never allocates, frees or keeps device memory.
 *   - `z` tensors are row-major contiguous (M, N, D); `rows` = M*N.
 *   - `params` is the reference's flat parameter matrix: the pointer addresses
 *     the FIRST parameter of this bijector inside row 0 and `param_row_stride`
 *     is the element distance between rows (so a column slice
 *     params[:, idx:idx+n] of a wider matrix is passed without a copy;
 *     trailing extra columns are legal, tests/test_bijectors.py:89-93).
 *     `param_row_stride == 0` broadcasts row 0 to every m (shared weights).
 *   - `dtype`: TNF_F32 or TNF_F64, the element type of z / params / log_det.
 *   - `accum`: TNF_LD_WRITE stores the log-det, TNF_LD_ADD / TNF_LD_SUB
 *     accumulate it into a running per-sample buffer (log_q_z -= log_det,
 *     density_estimator.py:387; sum_log_det += log_det, :405).
 *   - work is enqueued on `stream` (a cudaStream_t); no implicit sync.
 *   - return 0 on success, >0 a cudaError_t from the launch, <0 an argument
 *     error; tnf_last_error() gives a thread-local message.
 */
#ifndef TNF_H_
#define TNF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TNF_ABI_VERSION 3

enum { TNF_F32 = 0, TNF_F64 = 1 };
enum { TNF_FORWARD = 0, TNF_INVERSE = 1 };
enum { TNF_LD_WRITE = 0, TNF_LD_ADD = 1, TNF_LD_SUB = -1 };
enum { TNF_TC_BF16 = 0, TNF_TC_FP32 = 1 };
enum { TNF_ERR_ARG = -1, TNF_ERR_UNSUPPORTED = -2, TNF_ERR_ALIGN = -3 };

typedef void* tnf_stream_t; /* cudaStream_t */

int tnf_abi_version(void);
const char* tnf_last_error(void);
/* number of kernel launches issued through this library by the calling
 * process since load (bench.py's gpu_launches counter). */
int64_t tnf_launch_count(void);

/* ---- RealNVP coupling layer, exact-precision CUDA-core path -------------
 * replaces RealNVP.forward_and_log_det (bijectors.py:145-181) and
 * RealNVP.inverse_and_log_det (:183-206) incl. _t_s_layer (:208-242).
 * Any D >= 2 (odd D splits as :157-165), 1 <= L <= 5, 1 <= U <= 1000.
 * log_det is (M, N). z_out may alias z_in. */
int tnf_coupling(const void* z_in, void* z_out, void* log_det, const void* params,
                 int64_t param_row_stride, int64_t M, int64_t N, int D, int U, int L,
                 int transform_upper, int direction, int accum, int dtype, tnf_stream_t stream);

/* backward of tnf_coupling.  g_z_out (M,N,D) and g_log_det (M,N) (either may
 * be NULL = zero) are the gradients of the outputs; g_log_det is the gradient
 * w.r.t. the PLAIN log_det (sum s).  g_z_in (M,N,D) is written.  g_params
 * (row stride g_param_row_stride, 0 = one shared row) is ACCUMULATED into, so
 * the caller zero-fills it first.  Activations are recomputed from z_in. */
int tnf_coupling_bwd(const void* z_in, const void* params, int64_t param_row_stride,
                     const void* g_z_out, const void* g_log_det, void* g_z_in, void* g_params,
                     int64_t g_param_row_stride, int64_t M, int64_t N, int D, int U, int L,
                     int transform_upper, int direction, int dtype, tnf_stream_t stream);

/* The same backward for the conditional regime (one parameter row per m, N <= 32 samples per row): every element of the
 * row's slice of g_params is WRITTEN exactly once instead of accumulated - the caller need not zero-fill the (M,
 * D_params) gradient matrix and the kernel does not read it.  TNF_ERR_UNSUPPORTED when rows share parameters or a row
 * spans several tiles (the caller then zero-fills and uses tnf_coupling_bwd). */
int tnf_coupling_bwd_overwrite(const void* z_in, const void* params, int64_t param_row_stride,
                               const void* g_z_out, const void* g_log_det, void* g_z_in, void* g_params,
                               int64_t g_param_row_stride, int64_t M, int64_t N, int D, int U, int L,
                               int transform_upper, int direction, int dtype, tnf_stream_t stream);

/* ---- MAF: replaces MAF.forward_and_log_det / inverse_and_log_det (bijectors.py:742-796).
 * params row = per layer [W_mu (K*J), W_alpha (K*J)], no biases (:698-738); `mask` = the binary masks
 * Ms (:663-696) flattened in the same layout (each layer's mask twice), D_params floats, shared by all m.
 * TNF_INVERSE: one pass z' = (z - mu(z))/exp(alpha(z)).  TNF_FORWARD: D-1 passes z <- u*exp(alpha(z)) + mu(z)
 * with u the input (:751-756); log_det = sum alpha of the last pass.  The backward exists for the inverse
 * (log_prob / training) direction only. */
int tnf_maf(const void* z_in, void* z_out, void* log_det, const void* params, int64_t param_row_stride,
            const float* mask, int64_t M, int64_t N, int D, int U, int L, int direction, int accum,
            int dtype, tnf_stream_t stream);
int tnf_maf_bwd(const void* z_in, const void* params, int64_t param_row_stride, const float* mask,
                const void* g_z_out, const void* g_log_det, void* g_z_in, void* g_params,
                int64_t g_param_row_stride, int64_t M, int64_t N, int D, int U, int L, int direction,
                int dtype, tnf_stream_t stream);

/* ---- RealNVP coupling layer, tcgen05 tensor-core path (bf16 conditioner) --
 * Same math as tnf_coupling for shared weights (regime A: one parameter row),
 * fp32 z / log_det, conditioner GEMMs in bf16 on tcgen05 with fp32 TMEM
 * accumulation, affine transform and log-det in fp32.  D in {64,128,256},
 * U in {64,128,256}, 1 <= L <= 5 (tnf_tc_supported).
 * Weights are repacked once per parameter update by tnf_tc_pack (fp32 flat
 * params -> bf16 UMMA operand images + fp32 biases).
 * pre_scale/pre_shift (D floats each, or NULL): per-column affine
 * z <- z*pre_scale + pre_shift applied on load, which is how BatchNorm /
 * Affine neighbours (bijectors.py:292,399,423-424) are folded in.
 * col_stats (or NULL): receives [sum(D) | sumsq(D) | rows] (float64, the layout of
 * tnf_colstats) of the OUTPUT columns = the next BatchNorm's batch statistics
 * (:401-410), accumulated inside the kernel (fp32 per warp, float64 across warps);
 * needs stats_workspace of tnf_colstats_workspace_bytes(D) bytes; D <= 128 (variant 1: D = 64 only).
 * z_out must not alias z_in. */
int tnf_tc_supported(int D, int U, int L, int precision);
/* diagnostic: out[128 x N] = bf16(A[128 x K]) . bf16(W[K x N]) through the same
 * operand images, UMMA descriptors and TMEM accumulator layout as the fused
 * kernel; a_in_tmem selects the A operand source (1: TMEM, 0: SMEM image). */
int tnf_tc_selftest_gemm(const float* A, const float* W, float* out, int K, int N, int a_in_tmem,
                         tnf_stream_t stream);
/* precision: TNF_TC_BF16 = bf16 operands (stated tolerance max|dz| <= 5e-2);
 *            TNF_TC_FP32 = fp32-parity mode: operands split into fp16 hi + lo parts, three MMAs per
 *            product into the same fp32 accumulator, tanh / exp to fp32 accuracy.
 * The packed image of a layer depends on the precision. */
size_t tnf_tc_packed_bytes(int D, int U, int L, int precision);
int tnf_tc_pack(const float* params, void* packed, int D, int U, int L, int transform_upper,
                int precision, tnf_stream_t stream);
/* variant (diagnostic; no process-global state): 0 = kernel chosen by shape (D <= 128: CTA-pair two-tile kernel,
 *   for L = 2 and U >= 128 its N-half / TMEM-fed form coupling_tc5_kernel; D = 256: single-tile pipelined
 *   kernel), 1 = the first (8 epilogue warp) kernel, 2 = two-tile kernel without CTA pairs, 3 = coupling_tc5_kernel
 *   without its FMA-pipe tanh share, 4 = coupling_tc4_kernel; +16 = first kernel with one epilogue group only.
 * debug (diagnostic, or NULL): device buffer of 4096 int64 receiving (tag, clock64) stamps of CTA 0. */
int tnf_coupling_tc(const float* z_in, float* z_out, float* log_det, const void* packed,
                    int64_t rows, int D, int U, int L, int transform_upper, int direction,
                    int accum, const float* pre_scale, const float* pre_shift,
                    double* col_stats, void* stats_workspace, int precision, int variant, void* debug,
                    tnf_stream_t stream);

/* ---- backward of the tensor-core coupling layer (shared weights, bf16 conditioner, TNF_TC_BF16) ----
 * autograd through RealNVP.forward / inverse_and_log_det (bijectors.py:145-242) for D in {64,128}, U in {128,256},
 * L = 2 (tnf_tc_bwd_supported).  tnf_tc_bwd_pack re-lays one fp32 parameter row as the forward AND transposed bf16
 * UMMA operand images (+ fp32 biases) the kernel streams; tnf_coupling_tc_bwd recomputes the conditioner of every
 * 128-row tile from z_in on tcgen05, turns g_z_out (rows, D) and g_log_det (rows; gradient w.r.t. the plain sum s;
 * either may be NULL = zero) into g_z_in (rows, D), and stores the bf16 matrices the weight gradients are made of
 * into `workspace` (tnf_tc_bwd_workspace_bytes): per net n in {t, s}, in this order,
 *     h1, h2 (tanh outputs), d1, d2 (gradients of the hidden pre-activations): [n][4][rows][U + 16]
 *     d3 (gradient of the net's output):                                      [n][rows][D/2]  after the 8 matrices,
 *     xa = (conditioning half as the layer saw it | 1 | 0 ..):               [rows][64 (D = 64) or 128 (D = 128)]  last.
 * (row pitch U + 16: the pad columns of h1 / h2 are written as [1, 0, ..., 0], so that (h | 1)^T d yields the weight
 * gradient and, as its last row, the bias gradient in one GEMM; the pads of d1 / d2 are not written).
 * The weight gradients are then GEMMs over the batch (dW_l = a_{l-1}^T d_l, db_l = column sums of d_l, a_0 = the
 * conditioning half of z_in), left to the caller: a plain reduction with K = rows.
 * pre_scale / pre_shift (D floats each, or NULL): the layer was evaluated on z_in * pre_scale + pre_shift (a folded
 * BatchNorm with remembered statistics, as in tnf_coupling_tc); g_z_in is the gradient w.r.t. z_in itself. */
int tnf_tc_bwd_supported(int D, int U, int L);
size_t tnf_tc_bwd_packed_bytes(int D, int U, int L);
size_t tnf_tc_bwd_workspace_bytes(int64_t rows, int D, int U, int L);
int tnf_tc_bwd_pack(const float* params, void* packed, int D, int U, int L, int transform_upper,
                    tnf_stream_t stream);
int tnf_coupling_tc_bwd(const float* z_in, const void* packed, const float* g_z_out, const float* g_log_det,
                        float* g_z_in, void* workspace, int64_t rows, int D, int U, int L, int transform_upper,
                        int direction, const float* pre_scale, const float* pre_shift, tnf_stream_t stream);

/* ---- Affine: replaces Affine.forward_and_log_det / inverse_and_log_det
 * (bijectors.py:277-315).  params row = [alpha(D), shift(D)].
 * log_det (M) = sum alpha, written when non-NULL. */
int tnf_affine(const void* z_in, void* z_out, void* log_det, const void* params,
               int64_t param_row_stride, int64_t M, int64_t N, int D, int direction, int dtype,
               tnf_stream_t stream);
int tnf_affine_bwd(const void* z_in, const void* params, int64_t param_row_stride,
                   const void* g_z_out, const void* g_log_det, void* g_z_in, void* g_params,
                   int64_t g_param_row_stride, int64_t M, int64_t N, int D, int direction,
                   int dtype, tnf_stream_t stream);

/* ---- BatchNorm: replaces BatchNorm.forward_and_log_det / inverse_and_log_det
 * (bijectors.py:389-426).
 * tnf_colstats: deterministic two-stage reduction of per-column sum and sum
 *   of squares over `rows` rows -> sums[2*D+1] doubles: [sum | sumsq | rows]
 *   (this rank's partial; a data-parallel caller all-reduces the whole buffer,
 *   so the global row count travels with the sums and no host sync is needed).
 *   `workspace` must hold tnf_colstats_workspace_bytes(D) bytes.
 * tnf_bn_finalize: mean, alpha = sqrt(biased var + eps), log_det = -sum log alpha
 *   (all in `dtype`) from the (reduced) sums; the row count is sums[2*D].
 * tnf_bn_apply: TNF_FORWARD (z-mean)/alpha ; TNF_INVERSE z*alpha+mean. */
size_t tnf_colstats_workspace_bytes(int D);
int tnf_colstats(const void* z, int64_t rows, int D, double* sums, void* workspace, int dtype,
                 tnf_stream_t stream);
int tnf_bn_finalize(const double* sums, int D, double eps, void* mean, void* alpha,
                    void* log_det, int dtype, tnf_stream_t stream);
int tnf_bn_apply(const void* z_in, void* z_out, const void* mean, const void* alpha,
                 int64_t rows, int D, int direction, int dtype, tnf_stream_t stream);
/* backward of batch-statistics normalisation y=(z-mean)/alpha, incl. the
 * log-det term: g_z = (g_y - mean_r(g_y) - y*mean_r(g_y*y))/alpha - g_ld*(y/alpha)/R,
 * given the (all-reduced) column sums gsums = [sum g_y (D), sum g_y*y (D), -]
 * and `count` = pointer to the global row count R (device double). */
int tnf_bn_bwd_sums(const void* g_y, const void* y, int64_t rows, int D, double* gsums,
                    void* workspace, int dtype, tnf_stream_t stream);
int tnf_bn_bwd_apply(const void* g_y, const void* y, const void* alpha, const double* gsums,
                     const void* g_log_det, const double* count, void* g_z, int64_t rows, int D,
                     int dtype, tnf_stream_t stream);

/* ---- per-column affine maps: how BatchNorm / Affine (shared weights) are folded into the
 * neighbouring coupling kernel instead of costing an HBM pass each.
 * tnf_fold_colaffine composes a pending map z -> z*ps_in + pb_in (NULL = identity) with one more
 * bijector and writes the composed (ps_out, pb_out), D floats each:
 *   TNF_FOLD_BN_FWD   (z - a)/b        a = mean, b = alpha        bijectors.py:399
 *   TNF_FOLD_BN_INV   z*b + a                                     bijectors.py:423-424
 *   TNF_FOLD_AFF_FWD  exp(a)*z + b     a = alpha(D), b = shift(D)  bijectors.py:292
 *   TNF_FOLD_AFF_INV  (z - b)/exp(a)                              bijectors.py:312
 * For the Affine kinds ld_accum[0] += sum(alpha) when ld_accum is non-NULL (its log-det, :293,313).
 * tnf_colaffine applies a map to z (rows, D) when no coupling kernel follows. */
enum { TNF_FOLD_BN_FWD = 0, TNF_FOLD_BN_INV = 1, TNF_FOLD_AFF_FWD = 2, TNF_FOLD_AFF_INV = 3 };
int tnf_fold_colaffine(const float* ps_in, const float* pb_in, int kind, const float* a, const float* b,
                       float* ps_out, float* pb_out, float* ld_accum, int D, tnf_stream_t stream);
int tnf_colaffine(const float* z_in, float* z_out, const float* scale, const float* shift, int64_t rows,
                  int D, tnf_stream_t stream);

/* ---- ToInterval: replaces ToInterval.forward_and_log_det /
 * inverse_and_log_det (bijectors.py:509-557).  consts = 7*D floats
 * [tanh_flg | softplus_flg | tanh_m | tanh_c | softplus_m | softplus_c | log(tanh_m)]
 * (the float32 constants the reference builds at :475-480, plus the float32
 * log(tanh_m) it evaluates at :515). log_det (rows). */
int tnf_tointerval(const void* z_in, void* z_out, void* log_det, const float* consts,
                   int64_t rows, int D, int direction, int accum, int dtype, tnf_stream_t stream);
int tnf_tointerval_bwd(const void* z_in, const float* consts, const void* g_z_out,
                       const void* g_log_det, void* g_z_in, int64_t rows, int D, int direction,
                       int dtype, tnf_stream_t stream);

/* ---- ToSimplex: replaces ToSimplex.forward_and_log_det (bijectors.py:574-591).
 * z_in (rows, D_in) -> z_out (rows, D_in+1); D_attr is the bijector's D used in
 * the log-det.  No inverse exists in the reference. */
int tnf_tosimplex(const void* z_in, void* z_out, void* log_det, int64_t rows, int D_in,
                  int D_attr, int accum, int dtype, tnf_stream_t stream);
int tnf_tosimplex_bwd(const void* z_in, const void* g_z_out, const void* g_log_det, void* g_z_in,
                      int64_t rows, int D_in, int D_attr, int dtype, tnf_stream_t stream);

/* ---- base density of NormFlow
 * Log-dets come in two shapes: per sample (M,N) from
 * RealNVP / ToInterval / ToSimplex, and per parameter row from Affine (M,1)
 * and BatchNorm (scalar).  The chain keeps a per-sample accumulator `sub`
 * and a per-row accumulator `scal`; `scal_div` maps a sample row r to its
 * entry scal[r / scal_div] (N for per-m weights, rows for shared weights).
 * tnf_accum_bcast: dst[i] += src[i / div] for i < n_dst.
 * tnf_base_logprob: out(rows) = -sum(z^2)/2 - D*log(sqrt(2*pi)) - sub(rows) - scal
 *   (density_estimator.py:413-416; `sub`, `scal` may be NULL).
 * tnf_base_sample: omega ~ N(0,1) by Philox4x32-10 + Box-Muller into z (fp32)
 *   and the float64 base log-density of the same draw (:366-372).
 * tnf_base_logq: log_q(rows, f64) = -sum(omega^2)/2 - D*log(sqrt(2*pi)) for
 *   an injected fp32 omega (parity runs).
 * tnf_finish_logq: log_q(rows, f64) -= ld_acc(rows, dtype) + scal[r / scal_div]. */
int tnf_accum_bcast(void* dst, const void* src, int64_t n_dst, int64_t div, int dtype,
                    tnf_stream_t stream);
int tnf_base_logprob(const void* z, const void* sub, const void* scal, int64_t scal_div, void* out,
                     int64_t rows, int D, int dtype, tnf_stream_t stream);
int tnf_base_logprob_bwd(const void* z, const void* g_out, void* g_z, int64_t rows, int D,
                         int dtype, tnf_stream_t stream);
int tnf_base_sample(float* z, double* log_q, int64_t rows, int D, uint64_t seed,
                    uint64_t offset, tnf_stream_t stream);
int tnf_base_logq(const float* omega, double* log_q, int64_t rows, int D, tnf_stream_t stream);
int tnf_finish_logq(double* log_q, const void* ld_acc, const void* scal, int64_t scal_div,
                    int64_t rows, int dtype, tnf_stream_t stream);

/* ---- whole chains: replace the dispatch loops of NormFlow.forward
 * (density_estimator.py:374-388) and NormFlow.inverse_and_log_det + log_prob (:393-416).
 * The caller describes the chain once, in chain order, as an array of PODs; `params` is the flat
 * (M, D_params) float32 matrix (row stride 0 = one shared row) and every bijector reads its slice at
 * `param_offset` (:379-384, :398-402).  float32 z only.
 *   kind              TNF_BIJ_*
 *   num_layers, num_units, transform_upper   RealNVP arguments (ToSimplex: num_units = its D attribute)
 *   packed            RealNVP: tnf_tc_pack image at `tc_precision` -> tensor-core kernel; NULL -> exact CUDA-core
 *                     kernel
 *   bn_mean, bn_alpha, bn_log_det   BatchNorm state (D, D, 1 device floats): READ by tnf_chain_logprob and by
 *                     tnf_chain_sample(freeze_bn != 0) (remembered statistics, bijectors.py:397-399,420-426), WRITTEN
 *                     by tnf_chain_sample(freeze_bn == 0) (batch statistics, :401-415)
 *   consts            ToInterval: the 7 x D constant table of tnf_tointerval
 * Small shared-weight chains of RealNVP / BatchNorm / Affine (D <= 32, U <= 64, <= 32768 parameters; BatchNorm with
 * remembered statistics) run as ONE kernel with z in registers for the whole chain; all other chains run the
 * per-bijector kernels back to back (BatchNorm / Affine folded into the next tensor-core coupling layer).
 * tnf_chain_logprob:  log_prob (M, N) float32 = log N(z0; 0, I) - sum of log-dets.
 * tnf_chain_sample:   z_out (M, N, D [+1 after ToSimplex]) float32 and log_q (M, N) float64; `omega` (device,
 *   (M, N, D) float32) injects the base noise, NULL draws it on the device (Philox stream seed / offset).
 *   `allreduce` (or NULL): data-parallel hook, called on the host after each BatchNorm's statistics kernel has been
 *   enqueued; it must enqueue, on the same stream, an in-place sum over ranks of the 2*D+1 doubles in `stats_buf`
 *   (caller-allocated device buffer, required with the hook) and return 0.  `peer` (or NULL): exchange them over
 *   NVLink peer memory instead (tnf_peer_t below) wherever the BatchNorm is folded into a tensor-core coupling layer;
 *   the hook stays the path of every other chain.
 * `workspace`: tnf_chain_workspace_bytes(M, N, D) bytes of device memory, caller-owned. */
enum { TNF_BIJ_REALNVP = 0, TNF_BIJ_BATCHNORM = 1, TNF_BIJ_AFFINE = 2, TNF_BIJ_TOINTERVAL = 3, TNF_BIJ_TOSIMPLEX = 4 };
typedef struct tnf_bijector {
  int kind;
  int num_layers, num_units, transform_upper;
  int64_t param_offset;
  const void* packed;
  float* bn_mean;
  float* bn_alpha;
  float* bn_log_det;
  double bn_eps;
  const float* consts;
  void* ev_start; /* optional cudaEvent_t pair recorded on `stream` around this bijector's main kernel */
  void* ev_stop;  /* (profiling: per-kernel device time inside a chain call); NULL = off */
} tnf_bijector_t;
typedef int (*tnf_allreduce_fn)(double* stats_buf, int count, void* user);
/* Data-parallel BatchNorm statistics over NVLink peer memory (one process per GPU of one node; replaces the all-reduce
 * hook for chains whose BatchNorms are folded into tensor-core coupling layers): every rank holds a SYMMETRIC buffer
 * pair that all ranks can address,
 *   stats[r]  rank r's buffer of 2 x world x TNF_PEER_SLOT doubles ([sequence parity][source rank][sum | sumsq | rows])
 *   flags[r]  rank r's world sequence counters (unsigned 64-bit, zero before the first exchange)
 * The fold kernel of a BatchNorm stores this rank's 2 D + 1 sums into slot [seq & 1][rank] of EVERY rank's buffer
 * (peer stores), publishes `seq` in flags[r][rank] with a system-scope release, waits until its own counters all reached
 * `seq`, and adds the world's contributions in rank order - bit-identical on every rank, no host round trip, no
 * collective launch.  `seq` is the number of the call's first exchange: the caller keeps a counter that starts at 1
 * and advances by the number of BatchNorms per call, identically on every rank. */
#define TNF_PEER_MAX 8
#define TNF_PEER_SLOT 520
typedef struct tnf_peer {
  int rank, world;
  double* stats[TNF_PEER_MAX];
  unsigned long long* flags[TNF_PEER_MAX];
  unsigned long long seq;
} tnf_peer_t;
size_t tnf_chain_workspace_bytes(int64_t M, int64_t N, int D);
int tnf_chain_logprob(const tnf_bijector_t* chain, int n_bij, const float* z, const float* params,
                      int64_t param_row_stride, int64_t M, int64_t N, int D, int tc_precision,
                      float* log_prob, void* workspace, size_t workspace_bytes, tnf_stream_t stream);
int tnf_chain_sample(const tnf_bijector_t* chain, int n_bij, const float* params,
                     int64_t param_row_stride, int64_t M, int64_t N, int D, int tc_precision,
                     const float* omega, uint64_t seed, uint64_t offset, int freeze_bn,
                     tnf_allreduce_fn allreduce, void* allreduce_user, double* stats_buf, const tnf_peer_t* peer,
                     float* z_out, double* log_q, void* workspace, size_t workspace_bytes,
                     tnf_stream_t stream);

/* ---- hyper-network fusion: replaces ConditionalDensityEstimator.log_prob
 * (conditional_density_estimator.py:101-104: params = param_net(x); density_estimator.log_prob(z, params)) for one
 * sample per context (N = 1).  The hyper-network's LAST Linear (H -> D_params, :34-37) is evaluated inside the flow
 * kernel, per sample and in registers, so the (M, D_params) parameter matrix is never written to or read from HBM.
 *   chain / n_bij   the flow, as for tnf_chain_logprob: STAGES x [RealNVP(upper), BatchNorm, RealNVP(lower), BatchNorm,
 *                   Affine] (+ ToInterval), what NormFlow(arch_type='coupling') builds (density_estimator.py:260-282);
 *                   BatchNorm with its remembered statistics (bijectors.py:420-426)
 *   h (M, H)        output of the hyper-network's last hidden activation (the input of its last Linear), float32
 *   weight, bias    the last Linear's parameters in torch layout: weight (D_params, H) row-major, bias (D_params)
 *   packed          tnf_cde_packed_bytes(D_params, H) bytes, written by tnf_cde_pack once per parameter update: the
 *                   last Linear re-laid in the order the inverse chain consumes the flow parameters (blocks of 32)
 *   z (M, D)        the samples, log_prob (M) float32 = log N(z0; 0, I) - sum of log-dets
 *   variant         TNF_CDE_TC: the parameter rows h . W + b are produced by tcgen05 MMAs into tensor memory (fp32
 *                   parity by an fp16 hi / lo operand split, three MMAs per product) and read by the thread that owns
 *                   the sample's TMEM lane; TNF_CDE_CC: by fp32 FMAs in the consumer threads.  The packed layouts differ.
 * tnf_cde_supported: 1 when the chain has that structure and a compiled shape ((D, U) in {2, 4, 6, 8} x {15} and (8, 16), L = 2,
 * one stage), else 0 - the caller then materialises params and uses tnf_chain_logprob. */
enum { TNF_CDE_TC = 0, TNF_CDE_CC = 1 };
int tnf_cde_supported(const tnf_bijector_t* chain, int n_bij, int D, int H);
size_t tnf_cde_packed_bytes(int64_t D_params, int H, int variant);
int tnf_cde_pack(const tnf_bijector_t* chain, int n_bij, int D, const float* weight, const float* bias, int H,
                 void* packed, int variant, tnf_stream_t stream);
int tnf_cde_logprob(const tnf_bijector_t* chain, int n_bij, int D, const float* h, int H, const void* packed,
                    const float* z, int64_t M, float* log_prob, int variant, tnf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TNF_H_ */
