// cp.async.bulk global->shared throughput per SM from an L2-resident buffer (the packed weights: 320 KB read by
// every SM), as a function of copy size and the number of copies in flight.  One CTA per SM, one issuing thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) bench(const unsigned char* src, size_t src_bytes, int copy_bytes, int depth,
                                                int n_copies, long long* out, int n_thr) {
  extern __shared__ __align__(1024) unsigned char smem_all[];
  __shared__ uint64_t bar_all[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar_all + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool lanes_mode = n_thr < 0;
  const int n_iss = lanes_mode ? -n_thr : n_thr;
  if (lanes_mode ? (threadIdx.x < n_iss) : ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < n_iss)) {
    const int tw = lanes_mode ? threadIdx.x : (threadIdx.x >> 5);
    uint64_t* bar = bar_all + tw * 8;
    unsigned char* smem = smem_all + (size_t)tw * depth * copy_bytes;
    size_t off = ((size_t)(blockIdx.x * 4 + tw) * 16384) % src_bytes;
    const long long t0 = clock64();
    for (int i = 0; i < n_copies + depth; ++i) {
      const int slot = i % depth;
      if (i >= depth) {   // wait for the copy issued `depth` iterations ago in this slot
        const uint32_t parity = (uint32_t)(((i / depth) - 1) & 1);
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(
                         smem_u32(bar + slot)), "r"(parity) : "memory");
      }
      if (i < n_copies) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + slot)), "r"(copy_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + (size_t)slot * copy_bytes)), "l"(src + off), "r"(copy_bytes), "r"(smem_u32(bar + slot)) : "memory");
        off += copy_bytes;
        if (off + copy_bytes > src_bytes) off = 0;
      }
    }
    if (tw == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t src_bytes = 320 * 1024;
  unsigned char* src; cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
  long long* d; cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n_bytes_total = 8 << 20;   // per SM
  for (int n_thr : {1, 2, -2, -4})
  for (int copy_kb : {16})
    for (int depth : {2, 4}) {
      if (copy_kb * depth * (n_thr < 0 ? -n_thr : n_thr) > 192) continue;
      const int copy_bytes = copy_kb * 1024, n_copies = n_bytes_total / copy_bytes;
      for (int grid : {148}) {
        bench<<<grid, 128, 200 * 1024>>>(src, src_bytes, copy_bytes, depth, n_copies, d, n_thr);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < grid; ++i) c += h[i]; c /= grid;
        printf("%d issuing threads, copy %2d KB x %d in flight each, %3d SMs: %.1f B/cycle/SM, %.0f cycles per copy per thread  [%s]\n", n_thr, copy_kb, depth, grid,
               (double)n_bytes_total * (n_thr < 0 ? -n_thr : n_thr) / c, c / n_copies, cudaGetErrorString(e));
      }
    }
  return 0;
}
