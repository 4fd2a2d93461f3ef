// The first conv_block of the U-Net (enc1_block1: SeparableConv2D(64, 3) on the RGB image, reference
// model/u_net.py:14-20,63-66) as ONE kernel per direction.  With 3 input channels the depthwise half has almost no
// data and the pointwise half is a K = 3 contraction (1 flop/B): both are bound by the 64-channel tensor on the
// other side, so the fused kernels touch that tensor exactly once and everything else stays on chip.
//
//   forward : x[N,H,W,3] -> dw 3x3 (smem) -> pw 3->64 (registers) -> z (pre-BN) + batch statistics, or BN+ReLU applied
//   backward: x, dz[N,H,W,64] -> d(pointwise_kernel), d(depthwise_kernel)   (no input gradient: x is the image)
//
// A CTA walks 8x32-pixel tiles (grid-stride); 256 threads; a thread owns 8 output channels of 8 pixels of the tile,
// so a warp's 16-byte stores / loads cover 4 pixels x 128 B contiguous.
#include "common.cuh"
#include "ptx.cuh"

namespace unet {

constexpr int kStemCin = 3, kStemCout = 64, kTH = 8, kTW = 32;

template <typename T>
__device__ __forceinline__ void stem_load_tile(const T* __restrict__ x, int n, int h0, int w0, int H, int W, float* s_x) {
  // (kTH+2) x (kTW+2) x 3 halo tile, zero padded ('same'), as fp32
  constexpr int kRowElems = (kTW + 2) * kStemCin;
  for (int i = threadIdx.x; i < (kTH + 2) * kRowElems; i += blockDim.x) {
    const int r = i / kRowElems, e = i - r * kRowElems;
    const int hh = h0 - 1 + r, ww = w0 - 1 + e / kStemCin;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f32(x[(((int64_t)n * H + hh) * W + w0 - 1) * kStemCin + e]);
    s_x[i] = v;
  }
}

__device__ __forceinline__ void stem_depthwise(const float* s_x, const float* s_wd, float* s_d) {
  // one pixel per thread: d[px][ci] = sum_taps x * wd
  const int pr = threadIdx.x >> 5, pc = threadIdx.x & 31;
  float d[kStemCin] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
      for (int ci = 0; ci < kStemCin; ++ci)
        d[ci] = fmaf(s_x[((pr + a) * (kTW + 2) + pc + b) * kStemCin + ci], s_wd[(a * 3 + b) * kStemCin + ci], d[ci]);
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci) s_d[threadIdx.x * kStemCin + ci] = d[ci];
}

// STATS: out = z and batch statistics;  otherwise: out = act(z*scale+shift)
template <typename T, bool STATS>
__global__ void __launch_bounds__(256, 4)
stem_fwd_kernel(const T* __restrict__ x, const float* __restrict__ wd9c, const float* __restrict__ wp, T* __restrict__ out,
                int64_t ldo, int N, int H, int W, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                double* __restrict__ colsum, double* __restrict__ colsq, int tiles_h, int tiles_w, float* __restrict__ d_out) {
  pdl_enter();
  __shared__ float s_x[(kTH + 2) * (kTW + 2) * kStemCin];
  __shared__ float s_d[kTH * kTW * kStemCin];
  __shared__ float s_wd[9 * kStemCin];
  __shared__ float s_stat[2 * kStemCout];
  if (threadIdx.x < 9 * kStemCin) s_wd[threadIdx.x] = wd9c[threadIdx.x];
  if (threadIdx.x < 2 * kStemCout) s_stat[threadIdx.x] = 0.f;
  const int cg = threadIdx.x & 7, pslot = threadIdx.x >> 3;       // 8 channel groups x 32 pixel slots
  float w[kStemCin][8], sa[8], sb[8];            // STATS: running sum / sum of squares;  else: scale / shift
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci) load8(wp + ci * kStemCout + cg * 8, w[ci]);
#pragma unroll
  for (int j = 0; j < 8; ++j) { sa[j] = STATS ? 0.f : 1.f; sb[j] = 0.f; }
  if (!STATS && scale) load8(scale + cg * 8, sa);
  if (!STATS && shift) load8(shift + cg * 8, sb);
  const int64_t total = (int64_t)N * tiles_h * tiles_w;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int tw = (int)(t % tiles_w); const int64_t q = t / tiles_w;
    const int th = (int)(q % tiles_h); const int n = (int)(q / tiles_h);
    const int h0 = th * kTH, w0 = tw * kTW;
    __syncthreads();                                 // previous tile's s_d readers are done
    stem_load_tile<T>(x, n, h0, w0, H, W, s_x);
    __syncthreads();
    stem_depthwise(s_x, s_wd, s_d);
    if (d_out) {                                     // depthwise output of this thread's pixel, kept for the streaming backward
      const int hh = h0 + (threadIdx.x >> 5), ww = w0 + (threadIdx.x & 31);
      if (hh < H && ww < W) {
        float* dp = d_out + (((int64_t)n * H + hh) * W + ww) * 3;
        dp[0] = s_d[threadIdx.x * 3]; dp[1] = s_d[threadIdx.x * 3 + 1]; dp[2] = s_d[threadIdx.x * 3 + 2];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = pslot + 32 * i;                  // pixel of the tile: row p/32, column p%32
      const int hh = h0 + (p >> 5), ww = w0 + (p & 31);
      if (hh >= H || ww >= W) continue;
      const float d0 = s_d[p * 3], d1 = s_d[p * 3 + 1], d2 = s_d[p * 3 + 2];
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = fmaf(d2, w[2][j], fmaf(d1, w[1][j], d0 * w[0][j]));
        if (!STATS) { v = fmaf(v, sa[j], sb[j]); if (relu) v = fmaxf(v, 0.f); }
        o[j] = v;
      }
      store8(out + (((int64_t)n * H + hh) * W + ww) * ldo + cg * 8, o);
      if (STATS) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float r = round_to<T>(o[j]); sa[j] += r; sb[j] = fmaf(r, r, sb[j]); }
      }
    }
  }
  if (STATS) {
    // lanes with equal cg (lane & 7) hold the same channels: fold 4 pixel slots per warp, then shared, then global
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], 8); sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], 16);
      sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], 8); sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], 16);
    }
    __syncthreads();
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&s_stat[cg * 8 + j], sa[j]); atomicAdd(&s_stat[kStemCout + cg * 8 + j], sb[j]); }
    }
    __syncthreads();
    if (threadIdx.x < kStemCout) {
      atomicAdd(&colsum[threadIdx.x], (double)s_stat[threadIdx.x]);
      atomicAdd(&colsq[threadIdx.x], (double)s_stat[kStemCout + threadIdx.x]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ streaming forward
// The same first block as two barrier-free kernels: (1) depthwise 3x3 on the 3-channel image, one pixel per thread, fp32
// result d3 (12 B per pixel: 2 % of the 64-channel output), (2) pointwise 3 -> 64 as a pure stream over d3 whose only real
// traffic is the 128 B/pixel output; 8 lanes x 8 channels per pixel, a warp stores 4 pixels x 128 B contiguous.
template <typename T>
__global__ void __launch_bounds__(256)
stem_dw_kernel(const T* __restrict__ x, const float* __restrict__ wd9c, float* __restrict__ d3, int N, int H, int W) {
  pdl_enter();
  __shared__ float s_wd[9 * kStemCin];
  if (threadIdx.x < 9 * kStemCin) s_wd[threadIdx.x] = wd9c[threadIdx.x];
  __syncthreads();
  const int64_t M = (int64_t)N * H * W;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(m % W); const int64_t q = m / W; const int i = (int)(q % H);
    float d[kStemCin] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = i + a - 1;
      if (ii < 0 || ii >= H) continue;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int jj = j + b - 1;
        if (jj < 0 || jj >= W) continue;
        const T* px = x + (m + (int64_t)(a - 1) * W + (b - 1)) * kStemCin;
#pragma unroll
        for (int ci = 0; ci < kStemCin; ++ci) d[ci] = fmaf(to_f32(px[ci]), s_wd[(a * 3 + b) * kStemCin + ci], d[ci]);
      }
    }
    d3[m * 3] = d[0]; d3[m * 3 + 1] = d[1]; d3[m * 3 + 2] = d[2];
  }
}

// bf16, W % 4 == 0: four adjacent pixels per thread.  The 6 x 3 input columns of a row are 18 bf16 that start 6 bytes before a
// 24-byte-aligned offset: ten aligned 32-bit words cover them, so a thread issues 30 word loads for 4 pixels instead of 108
// two-byte loads (the one-pixel kernel is LSU-bound: 0.20 ms at 64 x 512 x 512 for 0.3 GB of traffic).  Same FMA order per
// pixel as stem_dw_kernel (out-of-image taps contribute fmaf(0, w, d) = d), so the results are bit-identical.
__global__ void __launch_bounds__(256)
stem_dw4_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wd9c, float* __restrict__ d3, int N, int H, int W) {
  pdl_enter();
  __shared__ float s_wd[9 * kStemCin];
  if (threadIdx.x < 9 * kStemCin) s_wd[threadIdx.x] = wd9c[threadIdx.x];
  __syncthreads();
  const int W4 = W >> 2;
  const int64_t Q = (int64_t)N * H * W4;
  const int row_elems = 3 * W;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < Q; t += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(t % W4); const int64_t r = t / W4; const int i = (int)(r % H);
    float d[4][kStemCin];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int ci = 0; ci < kStemCin; ++ci) d[p][ci] = 0.f;
    const int e_first = 12 * k - 4;                 // element (column * 3 + channel) of the first word of the window
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int ii = i + a - 1;
      if (ii < 0 || ii >= H) continue;
      const uint32_t* row = reinterpret_cast<const uint32_t*>(x + (r + (a - 1)) * row_elems);
      float f[20];                                  // window elements e_first .. e_first + 19; f[1 + 3 * c + ch] = column 4k-1+c
#pragma unroll
      for (int w = 0; w < 10; ++w) {
        const int e0 = e_first + 2 * w;
        const uint32_t u = (e0 >= 0 && e0 < row_elems) ? __ldg(row + (e0 >> 1)) : 0u;
        f[2 * w] = __uint_as_float(u << 16); f[2 * w + 1] = __uint_as_float(u & 0xffff0000u);
      }
#pragma unroll
      for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int ci = 0; ci < kStemCin; ++ci)
            d[p][ci] = fmaf(f[1 + 3 * (p + b) + ci], s_wd[(a * 3 + b) * kStemCin + ci], d[p][ci]);
    }
    float4* dst = reinterpret_cast<float4*>(d3 + t * 12);
    dst[0] = make_float4(d[0][0], d[0][1], d[0][2], d[1][0]);
    dst[1] = make_float4(d[1][1], d[1][2], d[2][0], d[2][1]);
    dst[2] = make_float4(d[2][2], d[3][0], d[3][1], d[3][2]);
  }
}

template <typename T> static bool stem_dw4_ok(const void*, int) { return false; }
template <> bool stem_dw4_ok<__nv_bfloat16>(const void* x, int W) { return W % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0; }
template <typename T> static void stem_dw_launch(const T* x, const float* wd9c, float* d3, int N, int H, int W, unsigned g1, cudaStream_t st) {
  launch_pdl(stem_dw_kernel<T>, g1, 256, 0, st, x, wd9c, d3, N, H, W);
}
template <> void stem_dw_launch<__nv_bfloat16>(const __nv_bfloat16* x, const float* wd9c, float* d3, int N, int H, int W, unsigned g1, cudaStream_t st) {
  if (stem_dw4_ok<__nv_bfloat16>(x, W) && aligned16(d3)) {
    const int64_t Q = (int64_t)N * H * (W / 4);
    launch_pdl(stem_dw4_kernel, (unsigned)i64min(ceil_div(Q, 256), (int64_t)sm_count() * 32), 256, 0, st, x, wd9c, d3, N, H, W);
  } else {
    launch_pdl(stem_dw_kernel<__nv_bfloat16>, g1, 256, 0, st, x, wd9c, d3, N, H, W);
  }
}

template <typename T, bool STATS>
__global__ void __launch_bounds__(256, 4)
stem_pw_kernel(const float* __restrict__ d3, const float* __restrict__ wp, T* __restrict__ out, int64_t ldo, int64_t M,
               const float* __restrict__ scale, const float* __restrict__ shift, int relu,
               double* __restrict__ colsum, double* __restrict__ colsq) {
  pdl_enter();
  __shared__ float s_stat[2 * kStemCout];
  if (threadIdx.x < 2 * kStemCout) s_stat[threadIdx.x] = 0.f;
  const int cg = threadIdx.x & 7, slot = threadIdx.x >> 3;
  float w[kStemCin][8], sa[8], sb[8];            // STATS: running sum / sum of squares;  else: scale / shift
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci) load8(wp + ci * kStemCout + cg * 8, w[ci]);
#pragma unroll
  for (int j = 0; j < 8; ++j) { sa[j] = STATS ? 0.f : 1.f; sb[j] = 0.f; }
  if (!STATS && scale) load8(scale + cg * 8, sa);
  if (!STATS && shift) load8(shift + cg * 8, sb);
  constexpr int U = 8;
  const int base = threadIdx.x & 24;              // first lane of this pixel's 8-lane group
  const int64_t stride = (int64_t)gridDim.x * 32 * U;
  for (int64_t m0 = (int64_t)blockIdx.x * 32 * U; m0 < M; m0 += stride) {
    float dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t m = m0 + u * 32 + slot;
      dv[u] = (m < M && cg < 3) ? __ldg(d3 + m * 3 + cg) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t m = m0 + u * 32 + slot;
      const float d0 = __shfl_sync(0xffffffffu, dv[u], base), d1 = __shfl_sync(0xffffffffu, dv[u], base + 1),
                  d2 = __shfl_sync(0xffffffffu, dv[u], base + 2);
      if (m >= M) continue;
      float o[8];
      const float2 d0p = make_float2(d0, d0), d1p = make_float2(d1, d1), d2p = make_float2(d2, d2);
#pragma unroll
      for (int j = 0; j < 8; j += 2) {              // packed FFMA2: two channels per instruction, same rounding as fmaf
        float2 v = fma2(d2p, make_float2(w[2][j], w[2][j + 1]), fma2(d1p, make_float2(w[1][j], w[1][j + 1]),
                        mul2(d0p, make_float2(w[0][j], w[0][j + 1]))));
        if (!STATS) {
          v = fma2(v, make_float2(sa[j], sa[j + 1]), make_float2(sb[j], sb[j + 1]));
          if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
        }
        o[j] = v.x; o[j + 1] = v.y;
      }
      store8(out + m * ldo + cg * 8, o);
      if (STATS) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float2 r = make_float2(round_to<T>(o[j]), round_to<T>(o[j + 1]));
          const float2 s2 = add2(make_float2(sa[j], sa[j + 1]), r), q2 = fma2(r, r, make_float2(sb[j], sb[j + 1]));
          sa[j] = s2.x; sa[j + 1] = s2.y; sb[j] = q2.x; sb[j + 1] = q2.y;
        }
      }
    }
  }
  if (STATS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], 8); sa[j] += __shfl_xor_sync(0xffffffffu, sa[j], 16);
      sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], 8); sb[j] += __shfl_xor_sync(0xffffffffu, sb[j], 16);
    }
    __syncthreads();
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(&s_stat[cg * 8 + j], sa[j]); atomicAdd(&s_stat[kStemCout + cg * 8 + j], sb[j]); }
    }
    __syncthreads();
    if (threadIdx.x < kStemCout) {
      atomicAdd(&colsum[threadIdx.x], (double)s_stat[threadIdx.x]);
      atomicAdd(&colsq[threadIdx.x], (double)s_stat[kStemCout + threadIdx.x]);
    }
  }
}

template <typename T> struct StemRaw;
template <> struct StemRaw<__nv_bfloat16> { uint4 a; };
template <> struct StemRaw<float> { float4 a, b; };
__device__ __forceinline__ void ldraw8(const __nv_bfloat16* p, StemRaw<__nv_bfloat16>& r) { r.a = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void ldraw8(const float* p, StemRaw<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void zero8(StemRaw<__nv_bfloat16>& r) { r.a = make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void zero8(StemRaw<float>& r) { r.a = make_float4(0.f, 0.f, 0.f, 0.f); r.b = r.a; }
__device__ __forceinline__ void unraw8(const StemRaw<__nv_bfloat16>& r, float (&v)[8]) {
  const uint32_t u[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = bf16lo_to_f32(u[i]); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}
__device__ __forceinline__ void unraw8(const StemRaw<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
stem_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dz, int64_t lddz, const float* __restrict__ wd9c,
                const float* __restrict__ wp, float* __restrict__ dwd9c, float* __restrict__ dwp, int N, int H, int W,
                int tiles_h, int tiles_w) {
  pdl_enter();
  __shared__ float s_x[(kTH + 2) * (kTW + 2) * kStemCin];
  __shared__ float s_d[kTH * kTW * kStemCin];
  __shared__ float s_dd[kTH * kTW * kStemCin];
  __shared__ float s_wd[9 * kStemCin];
  __shared__ float s_gp[kStemCin * kStemCout];
  __shared__ float s_gd[9 * kStemCin];
  if (threadIdx.x < 9 * kStemCin) { s_wd[threadIdx.x] = wd9c[threadIdx.x]; s_gd[threadIdx.x] = 0.f; }
  if (threadIdx.x < kStemCin * kStemCout) s_gp[threadIdx.x] = 0.f;
  const int cg = threadIdx.x & 7, pslot = threadIdx.x >> 3;
  float w[kStemCin][8], gp[kStemCin][8], gd[9][kStemCin];
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci) {
    load8(wp + ci * kStemCout + cg * 8, w[ci]);
#pragma unroll
    for (int j = 0; j < 8; ++j) gp[ci][j] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int ci = 0; ci < kStemCin; ++ci) gd[i][ci] = 0.f;
  const int64_t total = (int64_t)N * tiles_h * tiles_w;
  for (int64_t t = blockIdx.x; t < total; t += gridDim.x) {
    const int tw = (int)(t % tiles_w); const int64_t q = t / tiles_w;
    const int th = (int)(q % tiles_h); const int n = (int)(q / tiles_h);
    const int h0 = th * kTH, w0 = tw * kTW;
    StemRaw<T> graw[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = pslot + 32 * i;
      const int hh = h0 + (p >> 5), ww = w0 + (p & 31);
      if (hh < H && ww < W) ldraw8(dz + (((int64_t)n * H + hh) * W + ww) * lddz + cg * 8, graw[i]);
      else zero8(graw[i]);
    }
    __syncthreads();
    stem_load_tile<T>(x, n, h0, w0, H, W, s_x);
    __syncthreads();
    stem_depthwise(s_x, s_wd, s_d);
    __syncthreads();
    // pointwise gradient and dd = dz . Wp^T  (the tile's eight 16-byte dz loads are issued before the first use; they
    // were started before the depthwise phase so that their latency overlaps the two barriers above)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = pslot + 32 * i;
      float g[8];
      unraw8(graw[i], g);
      const float d0 = s_d[p * 3], d1 = s_d[p * 3 + 1], d2 = s_d[p * 3 + 2];
      float dd0 = 0.f, dd1 = 0.f, dd2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gp[0][j] = fmaf(d0, g[j], gp[0][j]); gp[1][j] = fmaf(d1, g[j], gp[1][j]); gp[2][j] = fmaf(d2, g[j], gp[2][j]);
        dd0 = fmaf(g[j], w[0][j], dd0); dd1 = fmaf(g[j], w[1][j], dd1); dd2 = fmaf(g[j], w[2][j], dd2);
      }
      // the 8 channel groups of a pixel are 8 adjacent lanes
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        dd0 += __shfl_xor_sync(0xffffffffu, dd0, o); dd1 += __shfl_xor_sync(0xffffffffu, dd1, o); dd2 += __shfl_xor_sync(0xffffffffu, dd2, o);
      }
      if (cg == 0) { s_dd[p * 3] = dd0; s_dd[p * 3 + 1] = dd1; s_dd[p * 3 + 2] = dd2; }
    }
    __syncthreads();
    // depthwise gradient: one pixel per thread
    {
      const int pr = threadIdx.x >> 5, pc = threadIdx.x & 31;
      const float e0 = s_dd[threadIdx.x * 3], e1 = s_dd[threadIdx.x * 3 + 1], e2 = s_dd[threadIdx.x * 3 + 2];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          const float* xp = s_x + ((pr + a) * (kTW + 2) + pc + b) * kStemCin;
          gd[a * 3 + b][0] = fmaf(xp[0], e0, gd[a * 3 + b][0]);
          gd[a * 3 + b][1] = fmaf(xp[1], e1, gd[a * 3 + b][1]);
          gd[a * 3 + b][2] = fmaf(xp[2], e2, gd[a * 3 + b][2]);
        }
    }
  }
  // fold and publish
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = gp[ci][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&s_gp[ci * kStemCout + cg * 8 + j], v);
    }
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int ci = 0; ci < kStemCin; ++ci) {
      const float v = warp_sum(gd[i][ci]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_gd[i * kStemCin + ci], v);
    }
  __syncthreads();
  if (threadIdx.x < kStemCin * kStemCout) atomicAdd(&dwp[threadIdx.x], s_gp[threadIdx.x]);
  if (threadIdx.x < 9 * kStemCin) atomicAdd(&dwd9c[threadIdx.x], s_gd[threadIdx.x]);
}

// Streaming backward of the first block with BatchNormalization backward folded in (see unet_bn_bwd_coef): no dz tensor,
// no shared-memory tiles, no block barriers.  Per pixel (8 lanes x 8 channels): dz = A*g + B*z + K in registers,
//   dwp[ci][co] += d[ci] * dz[co]        (d = depthwise output stored by the forward kernel, 3 floats per pixel)
//   dd[ci]       = sum_co dz[co] * wp[ci][co]   (8-lane shuffle reduction) -> stored, 3 values per pixel
// The depthwise weight gradient then is unet_dwconv3x3_bwd_weight(x, dd) on two 3-channel tensors.
// ring depth: 4 (bf16) x (16 B of g + 16 B of z [+ 4 B of d]) per thread in flight; fp32 (parity path) halves it to fit 48 KB
template <typename T>
__global__ void __launch_bounds__(256, 2)
stem_bwd_folded_kernel(const T* __restrict__ g, int64_t ldg, const T* __restrict__ z, const float* __restrict__ coef,
                       const float* __restrict__ d3, const float* __restrict__ wp, float* __restrict__ dwp, T* __restrict__ dd,
                       int64_t M) {
  pdl_enter();
  constexpr int kRaw = 8 * (int)sizeof(T);          // bytes of 8 channels
  constexpr int kSbD = sizeof(T) == 2 ? 4 : 2;
  __shared__ __align__(16) uint8_t ring_g[kSbD][256 * kRaw];
  __shared__ __align__(16) uint8_t ring_z[kSbD][256 * kRaw];
  __shared__ float ring_d[kSbD][256];
  __shared__ float s_gp[kStemCin * kStemCout];
  if (threadIdx.x < kStemCin * kStemCout) s_gp[threadIdx.x] = 0.f;
  __syncthreads();
  const int cg = threadIdx.x & 7, slot = threadIdx.x >> 3;
  float w[kStemCin][8], gp[kStemCin][8], ca[8], cb[8], ck[8];
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci) {
    load8(wp + ci * kStemCout + cg * 8, w[ci]);
#pragma unroll
    for (int j = 0; j < 8; ++j) gp[ci][j] = 0.f;
  }
  load8(coef + cg * 8, ca); load8(coef + kStemCout + cg * 8, cb); load8(coef + 2 * kStemCout + cg * 8, ck);
  const uint32_t sg = smem_u32(&ring_g[0][threadIdx.x * kRaw]), sz = smem_u32(&ring_z[0][threadIdx.x * kRaw]),
                 sd = smem_u32(&ring_d[0][threadIdx.x]);
  const int64_t first = (int64_t)blockIdx.x * 32 + slot;          // this thread's pixels: first + i * gridDim.x * 32
  const int64_t step = (int64_t)gridDim.x * 32;
  const int count = first < M ? (int)((M - first + step - 1) / step) : 0;
  const int base = threadIdx.x & 24;                // first lane of this pixel's 8-lane group (within the warp)
  auto issue = [&](int i) {
    if (i < count) {
      const int64_t m = first + (int64_t)i * step;
      const int sl = i % kSbD;
#pragma unroll
      for (int q = 0; q < kRaw / 16; ++q) {
        cp_async16(sg + sl * (256 * kRaw) + q * 16, reinterpret_cast<const uint8_t*>(g + m * ldg + cg * 8) + q * 16);
        cp_async16(sz + sl * (256 * kRaw) + q * 16, reinterpret_cast<const uint8_t*>(z + m * kStemCout + cg * 8) + q * 16);
      }
      if (cg < 3) cp_async4(sd + sl * (256 * 4), d3 + m * 3 + cg);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < kSbD; ++i) issue(i);
  // every lane of a warp runs the same number of trips only if the pixel counts agree; lanes past their count idle in the shuffles
  const int trips = __reduce_max_sync(0xffffffffu, count);
  for (int i = 0; i < trips; ++i) {
    cp_async_wait<kSbD - 1>();
    const int sl = i % kSbD;
    const bool live = i < count;
    StemRaw<T> graw, zraw;
    zero8(graw); zero8(zraw);
    float dv = 0.f;
    if (live) {
      if (sizeof(T) == 2) {
        const uint4 a = lds128u(sg + sl * (256 * kRaw)), bq = lds128u(sz + sl * (256 * kRaw));
        *reinterpret_cast<uint4*>(&graw) = a; *reinterpret_cast<uint4*>(&zraw) = bq;
      } else {
        uint4* gq = reinterpret_cast<uint4*>(&graw); uint4* zq = reinterpret_cast<uint4*>(&zraw);
        gq[0] = lds128u(sg + sl * (256 * kRaw)); gq[1] = lds128u(sg + sl * (256 * kRaw) + 16);
        zq[0] = lds128u(sz + sl * (256 * kRaw)); zq[1] = lds128u(sz + sl * (256 * kRaw) + 16);
      }
      if (cg < 3) dv = lds32f(sd + sl * (256 * 4));
    }
    float gv[8], zv[8];
    unraw8(graw, gv); unraw8(zraw, zv);
    const float d0 = __shfl_sync(0xffffffffu, dv, base), d1 = __shfl_sync(0xffffffffu, dv, base + 1),
                d2 = __shfl_sync(0xffffffffu, dv, base + 2);
    // packed FFMA2, two channels per instruction.  A pixel past this thread's count has d0 = d1 = d2 = 0 (dv stays 0), so its
    // dz (= K) adds nothing to gp, and its dd is not stored: no select on dz needed.
    float2 dd0p = make_float2(0.f, 0.f), dd1p = dd0p, dd2p = dd0p;
    const float2 d0p = make_float2(d0, d0), d1p = make_float2(d1, d1), d2p = make_float2(d2, d2);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const float2 dz = fma2(make_float2(ca[j], ca[j + 1]), make_float2(gv[j], gv[j + 1]),
                             fma2(make_float2(cb[j], cb[j + 1]), make_float2(zv[j], zv[j + 1]), make_float2(ck[j], ck[j + 1])));
      const float2 g0 = fma2(d0p, dz, make_float2(gp[0][j], gp[0][j + 1])), g1 = fma2(d1p, dz, make_float2(gp[1][j], gp[1][j + 1])),
                   g2 = fma2(d2p, dz, make_float2(gp[2][j], gp[2][j + 1]));
      gp[0][j] = g0.x; gp[0][j + 1] = g0.y; gp[1][j] = g1.x; gp[1][j + 1] = g1.y; gp[2][j] = g2.x; gp[2][j + 1] = g2.y;
      dd0p = fma2(dz, make_float2(w[0][j], w[0][j + 1]), dd0p);
      dd1p = fma2(dz, make_float2(w[1][j], w[1][j + 1]), dd1p);
      dd2p = fma2(dz, make_float2(w[2][j], w[2][j + 1]), dd2p);
    }
    float dd0 = dd0p.x + dd0p.y, dd1 = dd1p.x + dd1p.y, dd2 = dd2p.x + dd2p.y;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      dd0 += __shfl_xor_sync(0xffffffffu, dd0, o); dd1 += __shfl_xor_sync(0xffffffffu, dd1, o); dd2 += __shfl_xor_sync(0xffffffffu, dd2, o);
    }
    if (live && cg < 3) dd[(first + (int64_t)i * step) * 3 + cg] = from_f32<T>(cg == 0 ? dd0 : (cg == 1 ? dd1 : dd2));
    issue(i + kSbD);
  }
  cp_async_wait<0>();
#pragma unroll
  for (int ci = 0; ci < kStemCin; ++ci)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = gp[ci][j];
      v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
      if ((threadIdx.x & 31) < 8) atomicAdd(&s_gp[ci * kStemCout + cg * 8 + j], v);
    }
  __syncthreads();
  if (threadIdx.x < kStemCin * kStemCout) atomicAdd(&dwp[threadIdx.x], s_gp[threadIdx.x]);
}

static int stem_check(const char* who, const void* x, int N, int H, int W, int Cin, int Cout) {
  UNET_REQUIRE(x && N > 0 && H > 0 && W > 0, UNET_EINVAL, "%s: bad argument", who);
  UNET_REQUIRE(Cin == kStemCin && Cout == kStemCout, UNET_EUNSUPPORTED,
               "%s: the fused stem is built for %d -> %d channels (got %d -> %d); use dwconv3x3 + gemm", who, kStemCin,
               kStemCout, Cin, Cout);
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_stem_fwd(const void* x, const float* wd9c, const float* wp, void* out, int64_t ldo,
                             int N, int H, int W, int Cin, int Cout, int dtype,
                             const float* scale, const float* shift, int relu, double* colsum, double* colsq, float* d_out, void* stream) {
  if (int e = stem_check("stem_fwd", x, N, H, W, Cin, Cout)) return e;
  UNET_REQUIRE(wd9c && wp && out && ldo >= Cout, UNET_EINVAL, "stem_fwd: bad argument");
  UNET_REQUIRE((colsum == nullptr) == (colsq == nullptr), UNET_EINVAL, "stem_fwd: colsum/colsq must come together");
  UNET_REQUIRE(ldo % 8 == 0 && aligned16(out) && aligned16(wp) && (!scale || aligned16(scale)) && (!shift || aligned16(shift)),
               UNET_EALIGN, "stem_fwd: out / parameters must be 16B aligned, ldo%%8==0");
  const int th = (int)ceil_div(H, kTH), tw = (int)ceil_div(W, kTW);
  const int64_t tiles = (int64_t)N * th * tw;
  const unsigned grid = (unsigned)i64min(tiles, (int64_t)sm_count() * 8);
  cudaStream_t st = (cudaStream_t)stream;
  UNET_REQUIRE(!(colsum && (scale || shift)), UNET_EINVAL, "stem_fwd: statistics are taken on the raw contraction (no scale/shift)");
  if (d_out) {     // streaming pair: depthwise into the caller's d workspace, then a barrier-free pointwise stream
    const int64_t M = (int64_t)N * H * W;
    const unsigned g1 = (unsigned)i64min(ceil_div(M, 256), (int64_t)sm_count() * 32);
    const unsigned g2 = (unsigned)i64min(ceil_div(M, 256), (int64_t)sm_count() * 16);
#define STEM_STREAM(T) do { stem_dw_launch<T>((const T*)x, wd9c, d_out, N, H, W, g1, st); \
      if (colsum) launch_pdl(stem_pw_kernel<T, true>, g2, 256, 0, st, d_out, wp, (T*)out, ldo, M, scale, shift, relu, colsum, colsq); \
      else launch_pdl(stem_pw_kernel<T, false>, g2, 256, 0, st, d_out, wp, (T*)out, ldo, M, scale, shift, relu, colsum, colsq); } while (0)
    if (dtype == UNET_F32) STEM_STREAM(float);
    else if (dtype == UNET_BF16) STEM_STREAM(__nv_bfloat16);
    else return set_error(UNET_EINVAL, "stem_fwd: bad dtype %d", dtype);
#undef STEM_STREAM
    UNET_LAUNCH_CHECK("stem_fwd(stream)");
    return UNET_OK;
  }
#define STEM_FWD(T, S) launch_pdl(stem_fwd_kernel<T, S>, grid, 256, 0, st, (const T*)x, wd9c, wp, (T*)out, ldo, N, H, W, scale, shift, relu, colsum, colsq, th, tw, d_out)
  if (dtype == UNET_F32) { if (colsum) STEM_FWD(float, true); else STEM_FWD(float, false); }
  else if (dtype == UNET_BF16) { if (colsum) STEM_FWD(__nv_bfloat16, true); else STEM_FWD(__nv_bfloat16, false); }
#undef STEM_FWD
  else return set_error(UNET_EINVAL, "stem_fwd: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("stem_fwd");
  return UNET_OK;
}

extern "C" int unet_stem_bwd(const void* x, const void* dz, int64_t lddz, const float* wd9c, const float* wp,
                             float* dwd9c, float* dwp, int N, int H, int W, int Cin, int Cout, int dtype, void* stream) {
  if (int e = stem_check("stem_bwd", x, N, H, W, Cin, Cout)) return e;
  UNET_REQUIRE(dz && wd9c && wp && dwd9c && dwp && lddz >= Cout, UNET_EINVAL, "stem_bwd: bad argument");
  UNET_REQUIRE(lddz % 8 == 0 && aligned16(dz) && aligned16(wp), UNET_EALIGN, "stem_bwd: dz / wp must be 16B aligned, lddz%%8==0");
  const int th = (int)ceil_div(H, kTH), tw = (int)ceil_div(W, kTW);
  const int64_t tiles = (int64_t)N * th * tw;
  const unsigned grid = (unsigned)i64min(tiles, (int64_t)sm_count() * 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32)
    launch_pdl(stem_bwd_kernel<float>, grid, 256, 0, st, (const float*)x, (const float*)dz, lddz, wd9c, wp, dwd9c, dwp, N, H, W, th, tw);
  else if (dtype == UNET_BF16)
    launch_pdl(stem_bwd_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dz, lddz, wd9c, wp, dwd9c, dwp,
                                                        N, H, W, th, tw);
  else return set_error(UNET_EINVAL, "stem_bwd: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("stem_bwd");
  return UNET_OK;
}

extern "C" int unet_stem_bwd_folded(const void* g, int64_t ldg, const void* z, const float* coef, const float* d3,
                                    const float* wp, float* dwp, void* dd, int64_t M, int dtype, void* stream) {
  UNET_REQUIRE(g && z && coef && d3 && wp && dwp && dd && M > 0 && ldg >= kStemCout, UNET_EINVAL, "stem_bwd_folded: bad argument");
  UNET_REQUIRE(ldg % 8 == 0 && aligned16(g) && aligned16(z) && aligned16(wp) && aligned16(coef), UNET_EALIGN,
               "stem_bwd_folded: g / z / wp / coef must be 16B aligned, ldg%%8==0");
  const unsigned grid = (unsigned)i64min(ceil_div(M, 32 * 64), (int64_t)sm_count() * 12);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32)
    launch_pdl(stem_bwd_folded_kernel<float>, grid, 256, 0, st, (const float*)g, ldg, (const float*)z, coef, d3, wp, dwp, (float*)dd, M);
  else if (dtype == UNET_BF16)
    launch_pdl(stem_bwd_folded_kernel<__nv_bfloat16>, grid, 256, 0, st, (const __nv_bfloat16*)g, ldg, (const __nv_bfloat16*)z, coef, d3, wp, dwp,
                                                               (__nv_bfloat16*)dd, M);
  else return set_error(UNET_EINVAL, "stem_bwd_folded: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("stem_bwd_folded");
  return UNET_OK;
}
