// 2-CTA UMMA check (sm_100a): a cluster of two CTAs computes D[256 x 256] = A[256 x K] . B[256 x K]^T with
// tcgen05.mma.cta_group::2 (M = 256: 128 rows per CTA).  Each CTA holds ITS 128 rows of A and ITS half (128 of the 256
// N rows) of B in shared memory, K-major SWIZZLE_NONE images as in the coupling kernels; the leader CTA issues the MMAs,
// the commit is multicast to both CTAs, every CTA reads its 128 x 256 accumulator from its own TMEM.
// Answers: does each CTA need only half of B (yes if PASS), what goes into the descriptors, how alloc / commit look.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_2cta umma_2cta.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int rows) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(((uint32_t)rows * 16u >> 4) & 0x3FFF) << 16;   // LBO: bytes between 8-column K groups
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;                    // SBO: bytes between 8-row groups
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t img_off(int row, int k, int rows) { return (uint32_t)((k >> 3) * rows * 16 + row * 16 + (k & 7) * 2); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int K = 64, N = 256, MH = 128, NH = 128;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k2cta(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* aimg = smem;                    // 128 x K bf16
  unsigned char* bimg = smem + MH * K * 2;       // 128 (this CTA's half of N) x K bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(bimg + NH * K * 2);
  uint32_t* tb = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.x / 2;
  (void)pair;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = threadIdx.x; idx < MH * K; idx += blockDim.x) {
    const int r = idx / K, k = idx % K;
    *reinterpret_cast<__nv_bfloat16*>(aimg + img_off(r, k, MH)) = __float2bfloat16_rn(A[(rank * MH + r) * K + k]);
    *reinterpret_cast<__nv_bfloat16*>(bimg + img_off(r, k, NH)) = __float2bfloat16_rn(B[(rank * NH + r) * K + k]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tb)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();   // both CTAs' operand images and barriers are ready
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tb;
  if (rank == 0 && threadIdx.x == 0) {
    // instruction descriptor: D=f32, A=B=bf16, K-major both, N, M = 256
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    for (int k = 0; k < K; k += 16) {
      const uint64_t adesc = make_desc(smem_u32(aimg) + (uint32_t)(k >> 3) * (MH * 16u), MH);
      const uint64_t bdesc = make_desc(smem_u32(bimg) + (uint32_t)(k >> 3) * (NH * 16u), NH);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(adesc), "l"(bdesc),
                   "r"(idesc), "r"(k > 0 ? 1u : 0u) : "memory");
    }
    const uint16_t mask = 0x3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)), "h"(mask) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = warp * 32 + lane;
  for (int n0 = 0; n0 < N; n0 += 16) {
    uint32_t o[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]), "=r"(o[8]),
                   "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]), "=r"(o[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(size_t)(rank * MH + row) * N + n0 + j] = __uint_as_float(o[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

static float bf16r(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  const int M = 256;
  float *hA = (float*)malloc(M * K * 4), *hB = (float*)malloc(N * K * 4), *hO = (float*)malloc(M * N * 4);
  srand(1);
  for (int i = 0; i < M * K; ++i) hA[i] = (rand() % 17 - 8) / 8.0f;
  for (int i = 0; i < N * K; ++i) hB[i] = (rand() % 13 - 6) / 4.0f;
  float *dA, *dB, *dO;
  cudaMalloc(&dA, M * K * 4); cudaMalloc(&dB, N * K * 4); cudaMalloc(&dO, M * N * 4);
  cudaMemcpy(dA, hA, M * K * 4, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, N * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0, M * N * 4);
  const size_t smem = MH * K * 2 + NH * K * 2 + 64;
  k2cta<<<2, 128, smem>>>(dA, dB, dO);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  cudaMemcpy(hO, dO, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)bf16r(hA[m * K + k]) * bf16r(hB[n * K + k]);
      double d = fabs(s - hO[m * N + n]);
      if (d > maxerr) maxerr = d;
      if (d > 1e-3 && bad < 5) { printf("mismatch m=%d n=%d got %f want %f\n", m, n, hO[m * N + n], s); ++bad; }
    }
  printf("2-CTA UMMA M=256 N=%d K=%d: max |err| = %g -> %s\n", N, K, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  return 0;
}
