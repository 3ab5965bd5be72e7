// The tanh epilogue step in isolation: W warps per SM sub-partition, each looping over 32-column accumulator
// chunks: tcgen05.ld -> +bias -> MUFU.TANH -> bf16 pack -> st.shared (UMMA K-major image) [-> proxy fence + mbarrier
// arrive].  No MMA, no TMA: what does the MUFU-bound loop achieve on its own, and what does each ingredient cost?
//   flags: 1 = no tcgen05.ld, 2 = no st.shared, 4 = no publish (fence + arrive), 8 = no bias ld.shared
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o epi_step epi_step.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int FLAGS>
__global__ void __launch_bounds__(512, 1) bench(long long* out, int wps, int n_steps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* act = smem;                                  // 128 x 256 bf16 image (64 KB)
  float* s_bias = reinterpret_cast<float*>(smem + 65536);     // 256 floats
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 1024);
  uint32_t* tb = reinterpret_cast<uint32_t*>(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_bias[i] = 0.01f * i;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tb)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tb;
  const int q = warp & 3, par = warp >> 2;
  const int r_tile = q * 32 + lane;
  const uint32_t hcol = tmem + ((uint32_t)(q * 32) << 16);
  long long t0 = 0, t1 = 0;
  float sink = 0.f;
  if (par < wps) {
    uint32_t acc[32];
    float xa[32], xb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { xa[j] = 0.001f * (j + lane); acc[j] = 0x3c000000u + j; }
    if (!(FLAGS & 1)) tmem_ld32(hcol + (uint32_t)(par * 32), acc);
    t0 = clock64();
    auto step = [&](float (&cur)[32], float (&nxt)[32], int c) {
      const float4* b4 = reinterpret_cast<const float4*>(s_bias + ((c + 2) & 7) * 32);
      unsigned char* dst = act + (size_t)((c & 7) * 4) * 2048 + r_tile * 16;
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) cur[j + e] = tanh_fast(cur[j + e]);
        if (j < 16) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            float4 b = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
            if (!(FLAGS & 8)) b = b4[j / 2 + qq];
            const int o = 2 * j + 4 * qq;
            nxt[o] = __uint_as_float(acc[o]) + b.x;         nxt[o + 1] = __uint_as_float(acc[o + 1]) + b.y;
            nxt[o + 2] = __uint_as_float(acc[o + 2]) + b.z; nxt[o + 3] = __uint_as_float(acc[o + 3]) + b.w;
          }
        }
        if (j > 0 && !(FLAGS & 2)) {
          const int k = j - 8;
          *reinterpret_cast<uint4*>(dst + (k >> 3) * 2048) =
              make_uint4(pack_bf16(cur[k], cur[k + 1]), pack_bf16(cur[k + 2], cur[k + 3]),
                         pack_bf16(cur[k + 4], cur[k + 5]), pack_bf16(cur[k + 6], cur[k + 7]));
        }
        if (j == 8) {
          if (!(FLAGS & 1)) tmem_ld32(hcol + (uint32_t)(((c + 4) & 7) * 32), acc);
          __syncwarp();
        }
      }
      if (!(FLAGS & 2))
        *reinterpret_cast<uint4*>(dst + 3 * 2048) =
            make_uint4(pack_bf16(cur[24], cur[25]), pack_bf16(cur[26], cur[27]), pack_bf16(cur[28], cur[29]),
                       pack_bf16(cur[30], cur[31]));
      else sink += cur[24] + cur[31] + cur[0] + cur[9] + cur[17];
      if (!(FLAGS & 4)) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + (c & 7))) : "memory");
      }
    };
#pragma unroll 1
    for (int i = 0; i < n_steps; i += 2) {
      if (!(FLAGS & 1)) tc_wait_ld();
      step(xa, xb, par + 2 * i);
      if (!(FLAGS & 1)) tc_wait_ld();
      step(xb, xa, par + 2 * i + 2);
    }
    if (!(FLAGS & 1)) tc_wait_ld();
    t1 = clock64();
    sink += xa[3] + xb[5];
  }
  if (lane == 0 && par < wps) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (sink == 123.456f) out[0] = 0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int FLAGS>
void run(long long* d, const char* what) {
  const int n_steps = 256;
  cudaFuncSetAttribute(bench<FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  for (int wps : {1, 2, 4}) {
    cudaMemset(d, 0, 148 * 16 * 8);
    bench<FLAGS><<<148, 512, 70000>>>(d, wps, n_steps);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 16]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; int n = 0;
    for (int i = 0; i < 148 * 16; ++i) if (h[i] > 0) { c += h[i]; ++n; }
    c /= n;
    printf("%-40s warps/SMSP=%d: %.0f cycles per chunk step per warp -> MUFU pipe busy %.0f%%  [%s]\n", what, wps,
           c / n_steps, 100.0 * 256.0 * wps / (c / n_steps), cudaGetErrorString(e));
  }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16 * 8);
  run<0>(d, "full step");
  run<4>(d, "no publish");
  run<1>(d, "no tcgen05.ld");
  run<2>(d, "no st.shared");
  run<8>(d, "no bias ld.shared");
  run<15>(d, "tanh + bias add only");
  return 0;
}
