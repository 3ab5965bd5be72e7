// What does one step of the MMA-issuer loop cost?  tcgen05.mma pairs (M=128, K=16, SS, N given) issued
//   mode 0: back to back by one thread, one commit at the end
//   mode 1: + tcgen05.commit to an mbarrier after every pair
//   mode 2: + mbarrier.try_wait on an already-completed barrier before every pair
//   mode 3: + tcgen05.fence::after_thread_sync before every pair
//   mode 4: mode 3 run warp-uniformly (32 lanes wait, the elected lane issues)
//   mode 5: mode 0 but each pair's commit targets a DIFFERENT barrier that a second warp waits on (consumer wake-up)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_loop mma_loop.cu
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
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(1u) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128, 1) bench(long long* out, int N, int n_pairs, int mode) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int K_img = 256;
  unsigned char* A = smem;
  unsigned char* B = smem + 128 * K_img * 2;
  uint64_t* bar = reinterpret_cast<uint64_t*>(B + (size_t)N * K_img * 2);   // bar[0] final, bar[1] dummy commits, bar[2] pre-completed
  uint32_t* tb = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < (128 + N) * K_img / 2; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 2)) : "memory");   // phase 0 of bar[2] complete
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
  const uint32_t idesc = make_idesc(N);
  const uint64_t a0 = make_desc(smem_u32(A), 128), b0 = make_desc(smem_u32(B), N);
  const uint64_t a_step = (2u * 2048u) >> 4, b_step = (2u * (uint32_t)N * 16u) >> 4;
  const uint32_t bar_final = smem_u32(bar), bar_dummy = smem_u32(bar + 1), bar_done = smem_u32(bar + 2);
  const int warp = threadIdx.x / 32;
  if (warp == 0 && (mode == 4 || threadIdx.x == 0)) {
    uint32_t leader = 1;
    if (mode == 4) asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    const long long t0 = clock64();
    for (int i = 0; i < n_pairs; ++i) {
      const int k = (i & 7) * 2;
      if (mode >= 2) wait(bar_done, 0);
      if (mode >= 3) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (leader) {
        mma(tmem, a0 + (uint64_t)k * a_step, b0 + (uint64_t)k * b_step, idesc);
        mma(tmem, a0 + (uint64_t)(k + 1) * a_step, b0 + (uint64_t)(k + 1) * b_step, idesc);
        if (mode >= 1 && mode <= 4) commit(bar_dummy);
      }
    }
    if (leader) {
      commit(bar_final);
      wait(bar_final, 0);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int n_pairs = 512;
  for (int N : {256, 32})
    for (int mode = 0; mode <= 4; ++mode) {
      size_t smem = (size_t)(128 + N) * 256 * 2 + 64;
      cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      bench<<<148, 128, smem>>>(d, N, n_pairs, mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
      printf("N=%3d mode %d: %.1f cycles per MMA pair  [%s]\n", N, mode, c / n_pairs, cudaGetErrorString(e));
    }
  return 0;
}
