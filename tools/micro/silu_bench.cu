// Micro-benchmark: issue/MUFU cost of SiLU formulations on 16 warps per SM (the epilogue's warp count).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tanh_a(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE> __device__ __forceinline__ void silu_group(float (&x)[4]) {
  if (MODE == 0) {  // shared reciprocal of 4 (current)
    float e[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) e[i] = 1.0f + ex2(fminf(x[i] * -1.4426950408889634f, 30.0f));
    const float ab = e[0] * e[1], cd = e[2] * e[3];
    const float r = rcp(ab * cd);
    const float rab = r * cd, rcd = r * ab;
    x[0] *= rab * e[1]; x[1] *= rab * e[0]; x[2] *= rcd * e[3]; x[3] *= rcd * e[2];
  } else if (MODE == 1) {  // direct
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] *= rcp(1.0f + ex2(x[i] * -1.4426950408889634f));
  } else if (MODE == 2) {  // pairs
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      const float a = 1.0f + ex2(fminf(x[i] * -1.4426950408889634f, 60.0f)), b = 1.0f + ex2(fminf(x[i + 1] * -1.4426950408889634f, 60.0f));
      const float r = rcp(a * b);
      x[i] *= r * b; x[i + 1] *= r * a;
    }
  } else if (MODE == 3) {  // tanh
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float h = 0.5f * x[i]; x[i] = fmaf(h, tanh_a(h), h); }
  } else if (MODE == 4) {  // ex2 only (MUFU rate probe)
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = ex2(x[i]);
  } else if (MODE == 5) {  // FMUL only (issue rate probe)
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x[i] * 1.0001f;
  }
}
template <int MODE> __global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = (threadIdx.x * 16 + i) * 1e-4f - 0.4f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 16; g += 4) { float y[4] = {x[g], x[g + 1], x[g + 2], x[g + 3]}; silu_group<MODE>(y); x[g] = y[0] + 0.3f; x[g + 1] = y[1] + 0.3f; x[g + 2] = y[2] + 0.3f; x[g + 3] = y[3] + 0.3f; }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, float* d) {
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148, 512>>>(d, 100);
  cudaEventRecord(a); k<MODE><<<148, 512>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double elems = 148.0 * 512 * 16 * iters;
  printf("%-8s %8.3f ms  %.2f elem/clk/SM (at %d MHz nominal)  %.1f clk per 16-element chunk per warp-quad(4 warps/SMSP)\n", name, ms, elems / 148 / (ms * 1e-3 * clk * 1e3), clk / 1000,
         (ms * 1e-3 * clk * 1e3) / iters);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 512 * 4);
  run<0>("silu4", d); run<1>("direct", d); run<2>("pairs", d); run<3>("tanh", d); run<4>("ex2", d); run<5>("fmul", d);
  return 0;
}
