// Issue-to-completion rate of tcgen05.mma (kind::f16, M=128, K=16) on sm_100a, operands as in the coupling kernel:
// A from SMEM (SS) or TMEM (TS), B from SMEM, SWIZZLE_NONE K-major images.  One CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, int rows) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(((uint32_t)rows * 16u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int TS>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int N, int n_mma, int K_img) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* A = smem;                       // 128 x K_img bf16
  unsigned char* B = smem + 128 * K_img * 2;     // N x K_img bf16
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + (size_t)N * K_img * 2);
  uint32_t* tb = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (128 + N) * K_img / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tb;
  long long t0 = 0, t1 = 0;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(N);
    const int ksteps = K_img / 16;
    const uint64_t a0 = make_desc(smem_u32(A), 128), b0 = make_desc(smem_u32(B), N);
    const uint64_t a_step = (2u * 2048u) >> 4, b_step = (2u * (uint32_t)N * 16u) >> 4;
    (void)ksteps;
    t0 = clock64();
    for (int i = 0; i < n_mma; i += 16) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint64_t bdesc = b0 + (uint64_t)k * b_step;
        if (TS) {
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                       "r"(tmem + 256u + (uint32_t)(8 * k)), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        } else {
          const uint64_t adesc = a0 + (uint64_t)k * a_step;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                       "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                       "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
    t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int K_img = 256, n_mma = 1024;
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {256, 128, 64, 32}) {
      size_t smem = (size_t)(128 + N) * K_img * 2 + 64;
      if (ts) { cudaFuncSetAttribute(bench<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<1><<<148, 128, smem>>>(d, N, n_mma, K_img); }
      else { cudaFuncSetAttribute(bench<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); bench<0><<<148, 128, smem>>>(d, N, n_mma, K_img); }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("%s N=%3d: %.1f cycles per MMA (M=128,K=16)  [%s]\n", ts ? "TS" : "SS", N, c / n_mma, cudaGetErrorString(e));
    }
  return 0;
}
