// fp32_pipe_probe.cu — what the CUDA-core FP32 pipe of one B200 SM sustains for the instruction forms the depthwise kernels
// are made of: scalar FFMA (three register operands), packed FFMA2 (fma.rn.f32x2) with all-distinct operands, and FFMA2 with
// one operand shared by consecutive instructions (the operand-reuse form ptxas gives the depthwise taps).  The depthwise
// convolutions need 9 (forward) / 18 (backward) FMAs per element, so this rate — not the HBM rate — bounds them once the
// channel count per byte moved is high enough.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipe_probe tools/fp32_pipe_probe.cu && ./fp32_pipe_probe
// Prints FMA lanes per clock per SM and cycles per instruction per scheduler for each form at 1, 2, 3, 4 warps per scheduler.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int kIters = 2048, kChains = 12;

// MODE 0: FFMA, acc[i] = a[i] * b[i] + acc[i]          (three distinct registers per instruction)
// MODE 1: FFMA2, acc[i] = a[i] * b[i] + acc[i]         (three distinct register pairs)
// MODE 2: FFMA2, acc[i] = a[i] * s + acc[i]            (s shared by all: one operand can sit in the reuse cache)
// MODE 3: FFMA, acc[i] = a[i] * s + acc[i]
// MODE 4: FFMA, acc[i] = a[i] * c + acc[i], c a kernel parameter (constant-bank operand: only two register reads)
template <int MODE>
__global__ void probe(float* out, long long* cycles, const float* __restrict__ in, float cparam) {
  // operands come from memory, so that ptxas cannot fold any of them into immediate-form instructions
  float2 acc[kChains], a[kChains], b[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) {
    const float* q = in + (i * 6) * blockDim.x + threadIdx.x;
    acc[i] = make_float2(q[0], q[blockDim.x]);
    a[i] = make_float2(q[2 * blockDim.x], q[3 * blockDim.x]);
    b[i] = make_float2(q[4 * blockDim.x], q[5 * blockDim.x]);
  }
  const float2 s = b[kChains - 1];
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(a[i].x, b[i].x, acc[i].x); acc[i].y = fmaf(a[i].y, b[i].y, acc[i].y); }
      if (MODE == 1) acc[i] = fma2(a[i], b[i], acc[i]);
      if (MODE == 2) acc[i] = fma2(a[i], s, acc[i]);
      if (MODE == 3) { acc[i].x = fmaf(a[i].x, s.x, acc[i].x); acc[i].y = fmaf(a[i].y, s.y, acc[i].y); }
      if (MODE == 4) { acc[i].x = fmaf(a[i].x, cparam, acc[i].x); acc[i].y = fmaf(a[i].y, cparam, acc[i].y); }
    }
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> static void run(const char* name, int sms) {
  float *out, *in; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024); cudaMalloc(&cyc, sizeof(long long) * sms);
  cudaMalloc(&in, sizeof(float) * 6 * kChains * 1024);
  float h_in[6 * kChains * 1024];
  for (int i = 0; i < 6 * kChains * 1024; ++i) h_in[i] = 1e-3f * (float)((i * 2654435761u) >> 22) - 0.5f;
  cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
  for (int threads : {128, 256, 384, 512}) {
    long long mx = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(cyc, 0, sizeof(long long) * sms);
      probe<MODE><<<sms, threads>>>(out, cyc, in, 0.999f);
      if (cudaDeviceSynchronize() != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed\n", name); return; }
      long long h[256]; cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
      mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    }
    const double fmas = 2.0 * kChains * (double)kIters * threads;        // per block = per SM (one block per SM)
    const double instr_per_smsp = (double)kChains * kIters * (threads / 128) * ((MODE == 0 || MODE == 3 || MODE == 4) ? 2 : 1);
    printf("%-36s %4d threads/SM: %6.1f FMA lanes/clk/SM, %5.2f cycles per instruction per scheduler\n", name, threads,
           fmas / (double)mx, (double)mx / instr_per_smsp);
  }
  cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  printf("%s, %d SMs\n", p.name, sms);
  run<4>("FFMA  a*const+c (constant bank)", sms);
  run<0>("FFMA  a*b+c, all distinct", sms);
  run<3>("FFMA  a*s+c, s shared", sms);
  run<1>("FFMA2 a*b+c, all distinct", sms);
  run<2>("FFMA2 a*s+c, s shared", sms);
  return 0;
}
