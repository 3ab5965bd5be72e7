// Throughput of MUFU-class approximations on sm_100a: cycles per warp-instruction per SM sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu ; run: ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ float op(float x) {
  float y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  else if (OP == 3) {  // packed bf16x2 tanh
    unsigned u = __float_as_uint(x), r;
    asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(r) : "r"(u));
    y = __uint_as_float(r);
  } else if (OP == 4) {  // packed f16x2 tanh
    unsigned u = __float_as_uint(x), r;
    asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(u));
    y = __uint_as_float(r);
  } else {  // packed f16x2 ex2
    unsigned u = __float_as_uint(x), r;
    asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(u));
    y = __uint_as_float(r);
  }
  return y;
}

template <int OP>
__global__ void bench(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = 0.01f * (threadIdx.x + i);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = op<OP>(v[i]);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1024 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16, 32}) {   // warps per CTA, one CTA per SM -> warps/4 per sub-partition
    bench<OP><<<148, warps * 32>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double per_warp_inst_per_smsp = c / (double(iters) * 16 * (warps / 4));
    printf("%-22s warps/SMSP=%d  cycles per warp-instruction per SMSP = %.2f\n", name, warps / 4, per_warp_inst_per_smsp);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tanh.approx.f32");
  run<1>("ex2.approx.ftz.f32");
  run<2>("rcp.approx.ftz.f32");
  run<3>("tanh.approx.bf16x2");
  run<4>("tanh.approx.f16x2");
  run<5>("ex2.approx.f16x2");
  return 0;
}
