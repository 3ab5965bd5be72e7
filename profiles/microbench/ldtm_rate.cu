// tcgen05.ld throughput on sm_100a: cycles per LDTM.32x32b.x32 (4 KB per warp) with 1, 2, 4, 8 warps issuing
// back to back (warps 0-3 cover the four lane quadrants = one per SM sub-partition; 4-7 are second warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_rate ldtm_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(int active_warps, int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x / 32;
  if (warp == 0) {
    uint32_t a = (uint32_t)__cvta_generic_to_shared(&tmem_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(a));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint32_t r[32];
      const uint32_t col = base + (uint32_t)((i & 7) * 32 + (warp >> 2) * 256);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
          "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(col));
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int e = 0; e < 32; ++e) acc ^= r[e];
    }
    t1 = clock64();
  }
  __syncthreads();
  if (threadIdx.x % 32 == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

int main() {
  long long* out; uint32_t* sink;
  cudaMallocManaged(&out, 8 * sizeof(long long)); cudaMalloc(&sink, 256 * 4);
  const int iters = 4096;
  for (int w : {1, 2, 4, 8}) {
    k<<<1, 256>>>(w, iters, out, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    long long mx = 0; for (int i = 0; i < w; ++i) mx = out[i] > mx ? out[i] : mx;
    printf("warps %d: %.1f cycles per LDTM.x32 per warp (with wait), aggregate %.1f B/cycle/SM\n", w, (double)mx / iters,
           (double)w * 4096.0 * iters / mx);
  }
  return 0;
}
