// Depthwise 3x3 convolution (the depthwise half of SeparableConv2D, reference model/u_net.py:14-20).
// NHWC, zero 'same' padding, stride 1, depth multiplier 1, cross-correlation (no kernel flip).
//
// HBM-bound (18 flop per output element, ~4.5 flop/B in bf16): the kernels are organised so that every
// input element is requested from DRAM once.  Each thread owns 8 (forward) or 4 (weight gradient) consecutive
// channels of one image column and slides down a segment of rows keeping the running partial sums in
// registers; the three column taps of a row come from the two neighbouring threads' lines in L1.
#include "common.cuh"

namespace unet {

struct DropArgs { float keep, inv_keep; uint32_t seed; int on; int64_t ctot, c0; const uint32_t* seed_dev; };
__device__ __forceinline__ uint32_t drop_seed(const DropArgs& d) { return d.seed + (d.seed_dev ? __ldg(d.seed_dev) : 0u); }

static DropArgs make_drop(const unet_dropout* d) {
  DropArgs a{1.f, 1.f, 0u, 0, 0, 0, nullptr};
  if (d && d->rate > 0.f) {
    a.on = 1; a.keep = 1.f - d->rate; a.inv_keep = 1.f / (1.f - d->rate);
    a.seed = d->seed; a.ctot = d->ctot; a.c0 = d->c0; a.seed_dev = d->seed_dev;
  }
  return a;
}

// ------------------------------------------------------------------------------------------------ forward
template <typename T, bool AFFINE>
__global__ void __launch_bounds__(256)
dwconv3x3_vec8_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w9c,
                      T* __restrict__ y, int64_t ldy, int N, int H, int W, int C, int R, int nseg, int flip,
                      const float* __restrict__ in_scale, const float* __restrict__ in_shift, DropArgs dp) {
  const int cv = C >> 3;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int c8 = (int)(t % cv); t /= cv;
  const int wq = (int)(t % W);  t /= W;
  const int hs = (int)(t % nseg);
  const int64_t n = t / nseg;
  if (n >= N) return;
  const int c0 = c8 << 3;

  float k[9][8];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int src = flip ? 8 - i : i;
    load8(w9c + (int64_t)src * C + c0, k[i]);
  }
  float sc[8], sh[8];
  if (AFFINE) { load8(in_scale + c0, sc); load8(in_shift + c0, sh); }

  const int h0 = hs * R, h1 = min(H, h0 + R);
  const bool has_l = wq > 0, has_r = wq + 1 < W;
  const T* xcol = x + ((n * H) * (int64_t)W + wq) * ldx + c0;     // row 0 of this column
  T* ycol = y + ((n * H) * (int64_t)W + wq) * ldy + c0;
  const int64_t xrow = (int64_t)W * ldx, yrow = (int64_t)W * ldy;

  float prev[8], cur[8];   // prev: partial sum of output row r-1, cur: of output row r
#pragma unroll
  for (int j = 0; j < 8; ++j) { prev[j] = 0.f; cur[j] = 0.f; }

  for (int r = h0 - 1; r <= h1; ++r) {
    float a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; c[j] = 0.f; }
    if ((r >= 0) && (r < H)) {
      const T* p = xcol + r * xrow;
      load8(p, b);
      if (has_l) load8(p - ldx, a);
      if (has_r) load8(p + ldx, c);
      if (AFFINE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          b[j] = fmaxf(fmaf(b[j], sc[j], sh[j]), 0.f);
          if (has_l) a[j] = fmaxf(fmaf(a[j], sc[j], sh[j]), 0.f);
          if (has_r) c[j] = fmaxf(fmaf(c[j], sc[j], sh[j]), 0.f);
        }
      }
    }
    // output row r-1 is complete once kernel row 2 has seen input row r
    if (r - 1 >= h0) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = fmaf(k[8][j], c[j], fmaf(k[7][j], b[j], fmaf(k[6][j], a[j], prev[j])));
      if (dp.on) {
        const uint64_t base = (uint64_t)((n * H + (r - 1)) * (int64_t)W + wq) * dp.ctot + dp.c0 + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] *= dropout_mult(base + j, drop_seed(dp), dp.keep, dp.inv_keep);
      }
      store8(ycol + (r - 1) * yrow, o);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      prev[j] = fmaf(k[5][j], c[j], fmaf(k[4][j], b[j], fmaf(k[3][j], a[j], cur[j])));
      cur[j]  = fmaf(k[2][j], c[j], fmaf(k[1][j], b[j], k[0][j] * a[j]));
    }
  }
}

// any channel count (used for the 3-channel input image)
template <typename T, bool AFFINE>
__global__ void dwconv3x3_scalar_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w9c,
                                        T* __restrict__ y, int64_t ldy, int N, int H, int W, int C, int flip,
                                        const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                                        DropArgs dp) {
  const int64_t total = (int64_t)N * H * W * C;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C); int64_t p = t / C;
    const int wq = (int)(p % W); const int64_t q = p / W;
    const int h = (int)(q % H); const int64_t n = q / H;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int hh = h + a - 1;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int ww = wq + b - 1;
        if (ww < 0 || ww >= W) continue;
        float v = to_f32(x[((n * H + hh) * (int64_t)W + ww) * ldx + c]);
        if (AFFINE) v = fmaxf(fmaf(v, in_scale[c], in_shift[c]), 0.f);
        const int ki = flip ? 8 - (a * 3 + b) : a * 3 + b;
        acc = fmaf(v, w9c[(int64_t)ki * C + c], acc);
      }
    }
    if (dp.on) acc *= dropout_mult((uint64_t)p * dp.ctot + dp.c0 + c, drop_seed(dp), dp.keep, dp.inv_keep);
    y[p * ldy + c] = from_f32<T>(acc);
  }
}

template <typename T>
static int dw_fwd_launch(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy, int N, int H, int W, int C,
                         int flip, const float* in_scale, const float* in_shift, DropArgs dp, cudaStream_t st) {
  const bool vec = (C % 8 == 0) && (ldx % 8 == 0) && (ldy % 8 == 0) && aligned16(x) && aligned16(y) &&
                   aligned16(w9c) && (!in_scale || (aligned16(in_scale) && aligned16(in_shift)));
  if (vec) {
    // rows per thread: long segments amortise the 2 halo rows; shrink until the grid covers the machine twice
    int R = 32;
    const int64_t cols = (int64_t)N * W * (C / 8);
    const int64_t want = (int64_t)sm_count() * 2048 * 2;
    while (R > 4 && cols * ceil_div(H, R) < want) R >>= 1;
    const int nseg = (int)ceil_div(H, R);
    const int64_t threads = cols * nseg;
    const unsigned grid = (unsigned)ceil_div(threads, 256);
    if (in_scale)
      dwconv3x3_vec8_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, R, nseg, flip,
                                                          in_scale, in_shift, dp);
    else
      dwconv3x3_vec8_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, R, nseg, flip,
                                                           nullptr, nullptr, dp);
  } else {
    const int64_t total = (int64_t)N * H * W * C;
    const unsigned grid = (unsigned)i64min(ceil_div(total, 256), (int64_t)sm_count() * 32);
    if (in_scale)
      dwconv3x3_scalar_kernel<T, true><<<grid, 256, 0, st>>>((const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, flip,
                                                            in_scale, in_shift, dp);
    else
      dwconv3x3_scalar_kernel<T, false><<<grid, 256, 0, st>>>((const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, flip,
                                                             nullptr, nullptr, dp);
  }
  UNET_LAUNCH_CHECK("dwconv3x3_fwd");
  return UNET_OK;
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dw[a,b,c] = sum_{n,q,w} x[n,q,w+b-1,c] * dy[n,q-a+1,w,c]: one pass over x and dy; each thread keeps the 9x4
// partial sums of its 4 channels in registers across a grid-stride loop, then warp shuffle -> shared -> one
// global atomic per (tap, channel) per block.
template <typename T> __device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(a.x << 16); v[1] = __uint_as_float(a.x & 0xffff0000u);
  v[2] = __uint_as_float(a.y << 16); v[3] = __uint_as_float(a.y & 0xffff0000u);
}

template <typename T>
__global__ void __launch_bounds__(256)
dwconv3x3_bwd_weight_vec4_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                                 float* __restrict__ dw9c, int N, int H, int W, int C, int R, int nseg) {
  extern __shared__ float s_acc[];   // [9][C]
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();

  const int cv = C >> 2;
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int c4 = (int)(g % cv);
  const int c0 = c4 << 2;
  const int64_t item_stride = ((int64_t)gridDim.x * blockDim.x) / cv;
  const int64_t n_items = (int64_t)N * nseg * W;

  float acc[9][4];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t item = g / cv; item < n_items; item += item_stride) {
    const int wq = (int)(item % W); const int64_t t = item / W;
    const int hs = (int)(t % nseg); const int64_t n = t / nseg;
    const int h0 = hs * R, h1 = min(H, h0 + R);
    const bool has_l = wq > 0, has_r = wq + 1 < W;
    const T* xcol = x + ((n * H) * (int64_t)W + wq) * ldx + c0;
    const T* dcol = dy + ((n * H) * (int64_t)W + wq) * lddy + c0;
    const int64_t xrow = (int64_t)W * ldx, drow = (int64_t)W * lddy;

    // dy rows q-1, q, q+1 restricted to this segment's output rows [h0,h1)
    float dm[4] = {0, 0, 0, 0}, d0[4] = {0, 0, 0, 0}, dp[4] = {0, 0, 0, 0};
    if (h0 < h1) load4<T>(dcol + (int64_t)h0 * drow, dp);   // row h0 is "q+1" for q = h0-1
    for (int q = h0 - 1; q <= h1; ++q) {
      // at loop entry: dm = dy[q-1], d0 = dy[q], dp = dy[q+1] (zero outside [h0,h1))
      if (q >= 0 && q < H) {
        float a[4] = {0, 0, 0, 0}, b[4], c[4] = {0, 0, 0, 0};
        const T* p = xcol + q * xrow;
        load4<T>(p, b);
        if (has_l) load4<T>(p - ldx, a);
        if (has_r) load4<T>(p + ldx, c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // kernel row a multiplies input row q into output row q-a+1
          acc[0][j] = fmaf(a[j], dp[j], acc[0][j]); acc[1][j] = fmaf(b[j], dp[j], acc[1][j]); acc[2][j] = fmaf(c[j], dp[j], acc[2][j]);
          acc[3][j] = fmaf(a[j], d0[j], acc[3][j]); acc[4][j] = fmaf(b[j], d0[j], acc[4][j]); acc[5][j] = fmaf(c[j], d0[j], acc[5][j]);
          acc[6][j] = fmaf(a[j], dm[j], acc[6][j]); acc[7][j] = fmaf(b[j], dm[j], acc[7][j]); acc[8][j] = fmaf(c[j], dm[j], acc[8][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { dm[j] = d0[j]; d0[j] = dp[j]; dp[j] = 0.f; }
      if (q + 2 < h1) load4<T>(dcol + (int64_t)(q + 2) * drow, dp);
    }
  }

  // lanes that share a channel group differ by multiples of cv (when cv < 32)
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j];
      for (int o = 16; o >= cv && o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[i][j] = v;
    }
  const int lane = threadIdx.x & 31;
  if (cv >= 32 || lane < cv) {
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&s_acc[i * C + c0 + j], acc[i][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * C; i += blockDim.x) {
    const float v = s_acc[i];
    if (v != 0.f) atomicAdd(&dw9c[i], v);
  }
}

template <typename T>
__global__ void dwconv3x3_bwd_weight_scalar_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy,
                                                   int64_t lddy, float* __restrict__ dw9c, int N, int H, int W, int C) {
  // one block per (tap, channel); small C only
  const int tap = blockIdx.x / C, c = blockIdx.x % C;
  const int a = tap / 3, b = tap % 3;
  const int64_t total = (int64_t)N * H * W;
  float acc = 0.f;
  for (int64_t p = threadIdx.x; p < total; p += blockDim.x) {
    const int wq = (int)(p % W); const int64_t q = p / W;
    const int h = (int)(q % H); const int64_t n = q / H;
    const int hh = h + a - 1, ww = wq + b - 1;
    if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
    acc = fmaf(to_f32(x[((n * H + hh) * (int64_t)W + ww) * ldx + c]), to_f32(dy[p * lddy + c]), acc);
  }
  __shared__ float red[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(&dw9c[(int64_t)tap * C + c], v);
  }
}

template <typename T>
static int dw_bwd_weight_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c,
                                int N, int H, int W, int C, cudaStream_t st) {
  const int cv = C / 4;
  const bool vec = (C % 4 == 0) && (ldx % 8 == 0) && (lddy % 8 == 0) && aligned16(x) && aligned16(dy) &&
                   (256 % cv == 0 || cv % 256 == 0) && (9 * C * 4 <= 160 * 1024);
  if (vec) {
    const int R = 32;
    const int nseg = (int)ceil_div(H, R);
    const int64_t threads = (int64_t)N * nseg * W * cv;
    int64_t grid = i64min(ceil_div(threads, 256), (int64_t)sm_count() * 4);
    // the thread->channel mapping needs gridDim*256 to be a multiple of cv
    if (cv > 256) { const int64_t m = cv / 256; grid = ceil_div(grid, m) * m; }
    const size_t smem = (size_t)9 * C * sizeof(float);
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(dwconv3x3_bwd_weight_vec4_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dwconv3x3_bwd_weight_vec4_kernel<T><<<(unsigned)grid, 256, smem, st>>>((const T*)x, ldx, (const T*)dy, lddy, dw9c,
                                                                           N, H, W, C, R, nseg);
  } else {
    UNET_REQUIRE(C <= 64, UNET_EUNSUPPORTED, "dwconv3x3_bwd_weight: C=%d needs C%%4==0 and 16B-aligned views", C);
    dwconv3x3_bwd_weight_scalar_kernel<T><<<9 * C, 256, 0, st>>>((const T*)x, ldx, (const T*)dy, lddy, dw9c, N, H, W, C);
  }
  UNET_LAUNCH_CHECK("dwconv3x3_bwd_weight");
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy,
                                  int N, int H, int W, int C, int dtype, int flip,
                                  const float* in_scale, const float* in_shift,
                                  const unet_dropout* drop, void* stream) {
  UNET_REQUIRE(x && w9c && y, UNET_EINVAL, "dwconv3x3_fwd: null pointer");
  UNET_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, UNET_EINVAL, "dwconv3x3_fwd: bad dims %d %d %d %d", N, H, W, C);
  UNET_REQUIRE(ldx >= C && ldy >= C, UNET_EINVAL, "dwconv3x3_fwd: ld < C");
  UNET_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), UNET_EINVAL, "dwconv3x3_fwd: scale/shift must come together");
  const DropArgs dp = make_drop(drop);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32)  return dw_fwd_launch<float>(x, ldx, w9c, y, ldy, N, H, W, C, flip, in_scale, in_shift, dp, st);
  if (dtype == UNET_BF16) return dw_fwd_launch<__nv_bfloat16>(x, ldx, w9c, y, ldy, N, H, W, C, flip, in_scale, in_shift, dp, st);
  return set_error(UNET_EINVAL, "dwconv3x3_fwd: bad dtype %d", dtype);
}

extern "C" int unet_dwconv3x3_bwd_weight(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c,
                                         int N, int H, int W, int C, int dtype, void* stream) {
  UNET_REQUIRE(x && dy && dw9c, UNET_EINVAL, "dwconv3x3_bwd_weight: null pointer");
  UNET_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, UNET_EINVAL, "dwconv3x3_bwd_weight: bad dims");
  UNET_REQUIRE(ldx >= C && lddy >= C, UNET_EINVAL, "dwconv3x3_bwd_weight: ld < C");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32)  return dw_bwd_weight_launch<float>(x, ldx, dy, lddy, dw9c, N, H, W, C, st);
  if (dtype == UNET_BF16) return dw_bwd_weight_launch<__nv_bfloat16>(x, ldx, dy, lddy, dw9c, N, H, W, C, st);
  return set_error(UNET_EINVAL, "dwconv3x3_bwd_weight: bad dtype %d", dtype);
}
