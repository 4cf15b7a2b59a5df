// Instruction-throughput probe for B200 (sm_100a): warp-instructions per cycle per SM sub-partition for scalar FFMA,
// packed FFMA2, MUFU.EX2, F2FP and a SiLU-like mix, at 1..8 warps per sub-partition.  nvcc -arch=sm_100a; run on the GPU.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ex2(float a) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ float rcp(float a) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
__device__ __forceinline__ unsigned f2fp(float a, float b) { unsigned r; asm volatile("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }

template <int WHAT>
__global__ void probe(float* out, long long* cycles, int iters) {
  float x[8]; u64 y[8];
  for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = ((u64)__float_as_uint(x[i]) << 32) | __float_as_uint(x[i] + 0.5f); }
  const float b = 1.0001f, c = 1e-4f;
  const u64 b2 = ((u64)__float_as_uint(b) << 32) | __float_as_uint(b), c2 = ((u64)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (WHAT == 0) x[i] = ffma(x[i], b, c);
      if (WHAT == 1) y[i] = fma2(y[i], b2, c2);
      if (WHAT == 2) x[i] = ex2(x[i]);
      if (WHAT == 3) x[i] = __uint_as_float(f2fp(x[i], x[(i + 1) & 7]));
      if (WHAT == 4) { x[i] = ex2(x[i]); y[i] = fma2(y[i], b2, c2); y[i] = fma2(y[i], b2, c2); y[i] = fma2(y[i], b2, c2); y[i] = fma2(y[i], b2, c2); }
      if (WHAT == 5) { x[i] = ex2(x[i]); x[i] = ffma(x[i], b, c); x[i] = ffma(x[i], b, c); x[i] = ffma(x[i], b, c); x[i] = ffma(x[i], b, c); }
      if (WHAT == 6) x[i] = rcp(x[i]);
      if (WHAT == 7) x[i] = fminf(x[i], 60.f + i);
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float((unsigned)y[i]) + __uint_as_float((unsigned)(y[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int WHAT> void run(const char* name, int per_iter) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  for (int warps = 4; warps <= 32; warps *= 2) {
    probe<WHAT><<<148, warps * 32>>>(out, cyc, iters);
    probe<WHAT><<<148, warps * 32>>>(out, cyc, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double winstr_per_smsp = (double)iters * 8 * per_iter * (warps / 4.0);
    printf("%-28s warps/SMSP %d: %.3f warp-instr/clk/SMSP (%lld cycles)\n", name, warps / 4, winstr_per_smsp / h, h);
  }
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("FFMA scalar", 1); run<1>("FFMA2 packed", 1); run<2>("MUFU.EX2", 1); run<6>("MUFU.RCP", 1); run<3>("F2FP f16x2", 1); run<7>("FMNMX", 1);
  run<4>("EX2 + 4 FFMA2", 5); run<5>("EX2 + 4 FFMA", 5);
  cudaError_t e = cudaDeviceSynchronize(); printf("%s\n", cudaGetErrorString(e));
  return 0;
}
