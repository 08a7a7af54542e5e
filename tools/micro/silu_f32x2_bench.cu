// Micro-benchmark: the epilogue's bias + SiLU chain with scalar FP32 instructions vs Blackwell's packed f32x2 instructions
// (add/mul.f32x2 -> FADD2 / FMUL2 / FFMA2), 16 warps per SM like the conv kernel's epilogue.  Question: do the packed forms
// free issue slots (the epilogue-bound layers are issue-bound under the power cap)?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
struct f2 { float x, y; };
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  unsigned long long pa, pb, pd; f2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  unsigned long long pa, pb, pd; f2 d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(pa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(pb) : "f"(b.x), "f"(b.y));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(pd) : "l"(pa), "l"(pb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(pd));
  return d;
}
// scalar: what the kernel does today (bias add + shared-reciprocal SiLU of 4)
__device__ __forceinline__ void silu4_scalar(float (&x)[4], const float (&b)[4]) {
  float e[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { x[i] += b[i]; e[i] = 1.0f + ex2(fminf(x[i] * -1.4426950408889634f, 30.0f)); }
  const float ab = e[0] * e[1], cd = e[2] * e[3];
  const float r = rcp(ab * cd);
  const float rab = r * cd, rcd = r * ab;
  x[0] *= rab * e[1]; x[1] *= rab * e[0]; x[2] *= rcd * e[3]; x[3] *= rcd * e[2];
}
// packed: natural pairs (x0,x1), (x2,x3)
__device__ __forceinline__ void silu4_packed(float (&x)[4], const float (&b)[4]) {
  f2 x01 = add2(f2{x[0], x[1]}, f2{b[0], b[1]}), x23 = add2(f2{x[2], x[3]}, f2{b[2], b[3]});
  const f2 k{-1.4426950408889634f, -1.4426950408889634f}, one{1.0f, 1.0f};
  f2 t01 = mul2(x01, k), t23 = mul2(x23, k);
  f2 e01 = add2(f2{ex2(fminf(t01.x, 30.0f)), ex2(fminf(t01.y, 30.0f))}, one);
  f2 e23 = add2(f2{ex2(fminf(t23.x, 30.0f)), ex2(fminf(t23.y, 30.0f))}, one);
  const f2 m = mul2(e01, e23);           // (e0 e2, e1 e3)
  const float r = rcp(m.x * m.y);
  const float r13 = r * m.y, r02 = r * m.x;  // 1/(e0 e2), 1/(e1 e3)
  const f2 rr{r13, r02};
  const f2 i01 = mul2(rr, e23);          // (1/e0, 1/e1)
  const f2 i23 = mul2(rr, e01);          // (1/e2, 1/e3)
  x01 = mul2(x01, i01); x23 = mul2(x23, i23);
  x[0] = x01.x; x[1] = x01.y; x[2] = x23.x; x[3] = x23.y;
}
template <int MODE> __global__ void __launch_bounds__(512, 1) k(float* out, int iters) {
  float x[16], b[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = (threadIdx.x * 16 + i) * 1e-4f - 0.4f; b[i] = 0.01f * i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int g = 0; g < 16; g += 4) {
      float y[4] = {x[g], x[g + 1], x[g + 2], x[g + 3]};
      const float bb[4] = {b[g], b[g + 1], b[g + 2], b[g + 3]};
      if (MODE == 0) silu4_scalar(y, bb); else silu4_packed(y, bb);
      x[g] = y[0] + 0.3f; x[g + 1] = y[1] + 0.3f; x[g + 2] = y[2] + 0.3f; x[g + 3] = y[3] + 0.3f;
    }
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
  float h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-8s %8.3f ms   checksum %.6f %.6f\n", name, ms, h[0], h[3]);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 512 * 4);
  for (int rep = 0; rep < 2; ++rep) { run<0>("scalar", d); run<1>("packed", d); }
  return 0;
}
