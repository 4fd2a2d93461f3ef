// Issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(float* out, int iters, float s) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 m = make_float2(s, s * 1.0001f), b = make_float2(0.25f, 0.125f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { acc[i].x = fmaf(acc[i].x, m.x, b.x); acc[i].y = fmaf(acc[i].y, m.y, b.y); }
        else acc[i] = fma2(acc[i], m, b);
      }
  }
  float r = 0.f;
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 512 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<148, 512>>>(out, iters, 0.999f); else probe<1><<<148, 512>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fma = 148.0 * 512 * (double)iters * 64;     // scalar FMAs
      if (rep) printf("%s: %.3f ms, %.2f TFLOP/s fp32, %.1f scalar-FMA/clk/SM @1.9GHz\n", mode ? "FFMA2" : "FFMA ", ms,
                      2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.9e9);
    }
  }
  return 0;
}
