// Shared device/host helpers for libunet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <utility>
#include "../../include/unet_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libunet_b200 is written for sm_100a only"
#endif

namespace unet {

// ---------------------------------------------------------------- errors
int set_error(int code, const char* fmt, ...);
int set_cuda_error(cudaError_t e, const char* where);

#define UNET_REQUIRE(cond, code, ...)                          \
  do { if (!(cond)) return ::unet::set_error((code), __VA_ARGS__); } while (0)

#define UNET_LAUNCH_CHECK(where)                               \
  do { cudaError_t e__ = cudaGetLastError();                   \
       if (e__ != cudaSuccess) return ::unet::set_cuda_error(e__, where); } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t i64min(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t i64max(int64_t a, int64_t b) { return a > b ? a : b; }
int sm_count();

// ---------------------------------------------------------------- programmatic dependent launch (opt-in: UNET_B200_PDL=1)
// Every kernel begins with pdl_enter() (or its two halves around a prologue that touches only shared memory / TMEM / kernel
// parameters): `griddepcontrol.wait` holds the first global-memory access until the PREVIOUS kernel of the stream has completed
// and flushed, so a kernel launched with the programmatic-stream-serialization attribute may be scheduled before its
// predecessor has drained; `griddepcontrol.launch_dependents` lets the NEXT kernel be scheduled as soon as every CTA of this one
// has started.  Without the attribute both instructions are no-ops.
//   UNET_B200_PDL=0 (default)  plain launches.  MEASURED on B200, train512 CUDA-graph replay, same box, alternating runs:
//                              55.11 / 55.02 ms per step.
//   UNET_B200_PDL=1            attribute + early trigger: 56.03 / 55.95 ms per step (-1.7 %): programmatic edges cost more per
//                              graph node than the ~195 kernel boundaries of the step give back; inference 512x512 is unchanged
//                              within run-to-run noise (11.45 / 11.63 vs 11.88 / 11.60 ms).
//                              Eager launches of the same step: 55.79 / 56.02 ms without, 56.45 / 56.50 ms with (-1 %): the
//                              dependents' early CTAs take SM slots from the tail of the kernel they wait for.
// Kept as a switch (every kernel has the wait in place, so a future longer-tailed schedule can turn it on); off where it lost.
bool pdl_enabled();
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(std::forward<Args>(args))...);
}
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: remember per (kernel instantiation, device)
struct SmemAttrOnce { bool done[64] = {}; };
template <typename F>
inline cudaError_t ensure_dynamic_smem(SmemAttrOnce& once, F func, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && once.done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) once.done[dev] = true;
  return e;
}

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct DT;
template <> struct DT<float>         { static constexpr int id = UNET_F32;  };
template <> struct DT<__nv_bfloat16> { static constexpr int id = UNET_BF16; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats.  Pointers must be 16-byte aligned (bf16) / 16-byte aligned (fp32).
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t u[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {           // bf16 -> fp32 is a 16-bit shift
    v[2 * i]     = __uint_as_float(u[i] << 16);
    v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 a;
  a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]);
  a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = a;
}
// bf16 -> fp32 of the low half is `u << 16`; written as a byte permute so that it stays on the ALU pipe: ptxas turns the shift
// into IMAD.U32 (x * 65536), which lands on the FMA pipe — the one these kernels saturate (ncu r02: IMAD.U32 carried 11 % of
// the stall samples of the fused depthwise backward, almost all of them math-pipe throttle)
__device__ __forceinline__ float bf16lo_to_f32(uint32_t u) {
  uint32_t d;
  asm("prmt.b32 %0, %1, 0, 0x1044;" : "=r"(d) : "r"(u));     // bytes (lsb first): 0, 0, u.b0, u.b1
  return __uint_as_float(d);
}
// 8-byte shared-memory load through an explicit shared-window address (pointer arithmetic on the dynamic-smem base
// otherwise degrades to generic LD)
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(saddr));
  return r;
}
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t r;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(saddr));
  return r;
}
__device__ __forceinline__ void sts64(uint32_t saddr, const uint2& v) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds32f(uint32_t saddr) { float r; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(saddr)); return r; }
__device__ __forceinline__ void red_shared_add(uint32_t saddr, float v) {
  asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
// Per-thread asynchronous global -> shared copies (LDGSTS): bytes in flight without holding registers
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void* g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}
// Packed fp32x2 arithmetic (sm_100a FFMA2 / FMUL2): two IEEE round-to-nearest FMAs per issued instruction, bit-identical
// to two scalar fmaf() calls.  The depthwise kernels are issue-bound on the FMA pipe without it.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// Order-pinned forms: ptxas keeps volatile asm statements in source order, so a run of these that share one operand in the
// same position is issued back to back and the shared operand comes from the operand-reuse cache (2.25 instead of 3.0 FP32-pipe
// cycles per FFMA2, tools/fp32_pipe_probe.cu); left to itself ptxas flags ~30 % of a depthwise kernel's FFMA2s.
__device__ __forceinline__ float2 fma2v(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2v(float2 a, float2 b) {
  float2 d;
  asm volatile("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// value a store8/load8 round trip would produce (so statistics match what is stored)
template <typename T> __device__ __forceinline__ float round_to(float v) { return to_f32(from_f32<T>(v)); }

// ---------------------------------------------------------------- dropout mask (stateless, reproducible on host)
// Elements are taken in GROUPS OF FOUR consecutive linear offsets (4k .. 4k+3; one 8-byte bf16 vector).  A group draws two
// 32-bit words from its index and the seed,
//     a = mix(k ^ seedmix),   b = a * 0xC2B2AE3D; b ^= b >> 16,        mix(x): x *= 0x9E3779B1; x ^= x >> 15; x *= 0x85EBCA77; x ^= x >> 13
// and element 4k+j keeps iff its 16-bit field (a.lo, a.hi, b.lo, b.hi for j = 0..3) is below floor(keep_prob * 65536)
// (|P(keep) - keep_prob| < 2^-16).  10 integer instructions per four elements instead of one full avalanche hash per pair:
// the mask is recomputed by every producer and consumer of a dropped tensor (never stored), so it sits on the critical path
// of issue-bound kernels.  Group indices are 32-bit: tensors of up to 2^34 elements (the mask repeats beyond).
// oracle/unet_ref.py::dropout_multiplier restates this on the host.
__host__ __device__ __forceinline__ uint32_t dropout_seedmix(uint32_t seed) { return seed * 0x85EBCA6Bu + 0xC2B2AE35u; }
__host__ __device__ __forceinline__ void dropout_words(uint32_t group, uint32_t seedmix, uint32_t& a, uint32_t& b) {
  uint32_t x = (group ^ seedmix) * 0x9E3779B1u;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
  a = x;
  uint32_t y = x * 0xC2B2AE3Du;
  y ^= y >> 16;
  b = y;
}
// threshold in the HIGH half of a word: field < thr  <=>  (field << 16) < (thr << 16), so the high field compares in one instruction
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float keep_prob) {
  uint32_t t = static_cast<uint32_t>(keep_prob * 65536.0f);
  return (t > 65535u ? 65535u : t) << 16;
}
// the 32-bit word holding element idx's field (host helper / tests)
__host__ __device__ __forceinline__ uint32_t dropout_hash(uint64_t idx, uint32_t seed) {
  uint32_t a, b;
  dropout_words(static_cast<uint32_t>(idx >> 2), dropout_seedmix(seed), a, b);
  return (idx & 2) ? b : a;
}
// multiplier for element idx: 0 if dropped, 1/(1-rate) if kept
__host__ __device__ __forceinline__ float dropout_mult(uint64_t idx, uint32_t seed, float keep_prob, float inv_keep) {
  const uint32_t h = dropout_hash(idx, seed);
  const uint32_t f = (idx & 1) ? h : (h << 16);
  return f < dropout_threshold(keep_prob) ? inv_keep : 0.0f;
}
// n consecutive elements starting at an EVEN index (n even); runs that start on a multiple of 4 with n % 4 == 0 (every bf16
// kernel: a thread owns 4, 8 or 32 channels) draw one pair of words per four elements
// PRESCALED: the caller has already multiplied the kept values by 1/keep_prob (folded into an FMA it does anyway); only zero the rest
// ALIGNED4: the caller guarantees even_base % 4 == 0 (no code for the generic path: the tensor-core epilogues are large enough to
// feel every duplicated loop in the instruction cache)
template <int N, bool PRESCALED = false, bool ALIGNED4 = false>
__device__ __forceinline__ void dropout_apply(float (&v)[N], uint64_t even_base, uint32_t seed, float keep_prob, float inv_keep) {
  const uint32_t thr = dropout_threshold(keep_prob);
  const uint32_t sm = dropout_seedmix(seed);
  const float m = PRESCALED ? 1.0f : inv_keep;
  if (N % 4 == 0 && (ALIGNED4 || (even_base & 2) == 0)) {
    const uint32_t g0 = static_cast<uint32_t>(even_base >> 2);
#pragma unroll
    for (int j = 0; j < N / 4; ++j) {
      uint32_t a, b;
      dropout_words(g0 + j, sm, a, b);
      v[4 * j]     = (a << 16) < thr ? (PRESCALED ? v[4 * j] : v[4 * j] * m) : 0.0f;
      v[4 * j + 1] = a < thr ? (PRESCALED ? v[4 * j + 1] : v[4 * j + 1] * m) : 0.0f;
      v[4 * j + 2] = (b << 16) < thr ? (PRESCALED ? v[4 * j + 2] : v[4 * j + 2] * m) : 0.0f;
      v[4 * j + 3] = b < thr ? (PRESCALED ? v[4 * j + 3] : v[4 * j + 3] * m) : 0.0f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      const uint32_t h = dropout_hash(even_base + j, seed);
      v[j]     = (h << 16) < thr ? (PRESCALED ? v[j] : v[j] * m) : 0.0f;
      v[j + 1] = h < thr ? (PRESCALED ? v[j + 1] : v[j + 1] * m) : 0.0f;
    }
  }
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace unet
