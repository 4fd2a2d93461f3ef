// HBM-bound elementwise / reduction kernels of the U-Net hot path: BatchNormalization (reference
// model/u_net.py:23), Activation('relu') (:25), MaxPooling2D (:69), Dropout (:78,98), the Conv2DTranspose
// gradient gather, Keras-form AdamW (scripts/train.py:226), MeanIoU confusion counts (scripts/train.py:231,
// scripts/benchmark.py:237-269) and parameter staging.  All are one-pass, 16-byte vectorised over the NHWC channel
// dimension, fp32 arithmetic.
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

static unsigned grid_for(int64_t threads, int block = 256) { return (unsigned)ceil_div(threads, block); }

// ------------------------------------------------------------------------------------------------ BN fold / finalize
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                               float* scale, float* shift, int C) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float s = g / sqrtf(var[c] + eps);
  scale[c] = s;
  shift[c] = b - mean[c] * s;
}

__global__ void bn_finalize_kernel(const double* colsum, const double* colsq, double inv_count,
                                   const float* gamma, const float* beta, float eps, float momentum,
                                   float* moving_mean, float* moving_var, float* scale, float* shift,
                                   float* save_mean, float* save_rstd, int C) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = colsum[c] * inv_count;
  double var = colsq[c] * inv_count - mean * mean;   // biased, as Keras uses for both normalisation and moving_var
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float s = g * rstd;
  scale[c] = s;
  shift[c] = b - (float)mean * s;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_rstd) save_rstd[c] = rstd;
  if (moving_mean) moving_mean[c] = moving_mean[c] * momentum + (float)mean * (1.f - momentum);
  if (moving_var)  moving_var[c]  = moving_var[c]  * momentum + (float)var  * (1.f - momentum);
}

// ------------------------------------------------------------------------------------------------ BN backward folded into the GEMMs
// BatchNormalization backward is affine per channel in (g, z):  dz = A*g + B*z + K  with g = dy*[y>0],
//   c1 = mean(g), c2 = mean(g*xhat), A = gamma*rstd, B = -A*c2*rstd, K = A*(c2*rstd*mean - c1),
// so the pointwise data gradient dz*W^T and weight gradient d^T*dz never need dz in memory:
//   dd = [g | z] * [W diag(A) | W diag(B)]^T + (K W^T)          (one GEMM, K-concatenated A operand, bias epilogue)
//   dW = (d^T g) diag(A) + (d^T z) diag(B) + colsum(d) K^T       (one GEMM, N-concatenated B operand, then bn_bwd_wgrad_combine)
// The two reductions arrive as sums[0] = sum(g), sums[1] = sum(g*y) from the kernel that produced g (y = gamma*xhat + beta
// wherever g != 0, so sum(g*xhat) = (sums[1] - beta*sums[0]) / gamma).  One block per row of W (input channel).
__global__ void __launch_bounds__(256)
bn_bwd_coef_kernel(const float* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const float* __restrict__ mean, const float* __restrict__ rstd, float inv_count,
                   float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                   const float* __restrict__ w, int Cin, int C, __nv_bfloat16* __restrict__ wab, float* __restrict__ bias,
                   int* __restrict__ ill_conditioned) {
  pdl_enter();
  __shared__ float s_red[8];
  const int i = blockIdx.x;
  float part = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gamma[c], b = beta[c];
    const float s1 = sums[c], s2 = sums[C + c];
    const float dg = g != 0.f ? (s2 - b * s1) / g : 0.f;     // sum(g * xhat)
    // (s2 - b*s1)/g is formed from bf16-rounded activations y = g*xhat + b: once |g| << |b| the rounding of y (2^-9 |b|) swamps
    // g*xhat and dividing by g amplifies it (and g == 0 loses dgamma altogether).  Tell the host, which then takes the
    // explicit reduce/apply schedule (sum(g*xhat) straight from z) for the rest of training.
    if (i == 0 && ill_conditioned && fabsf(g) < 0.0625f * fabsf(b)) *ill_conditioned = 1;
    const float c1 = s1 * inv_count, c2 = dg * inv_count;
    const float A = g * rstd[c];
    const float B = -A * c2 * rstd[c];
    const float K = A * (c2 * rstd[c] * mean[c] - c1);
    if (i == 0) {
      if (dgamma) dgamma[c] += dg;
      if (dbeta) dbeta[c] += s1;
      if (coef) { coef[c] = A; coef[C + c] = B; coef[2 * C + c] = K; }
    }
    if (w) {
      const float wv = w[(int64_t)i * C + c];
      wab[(int64_t)i * 2 * C + c] = __float2bfloat16_rn(wv * A);
      wab[(int64_t)i * 2 * C + C + c] = __float2bfloat16_rn(wv * B);
      part = fmaf(wv, K, part);
    }
  }
  if (w) {
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += s_red[k];
      bias[i] = t;
    }
  }
}

// dW[i,c] += G[i,c]*A[c] + G[i,C+c]*B[c] + sd[i]*K[c]     (G = d^T [g | z], sd = column sums of d)
__global__ void bn_bwd_wgrad_combine_kernel(const float* __restrict__ G, const float* __restrict__ coef, const float* __restrict__ sd,
                                            float* __restrict__ dw, int Cin, int C) {
  pdl_enter();
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)Cin * C) return;
  const int i = (int)(idx / C), c = (int)(idx % C);
  dw[idx] += G[(int64_t)i * 2 * C + c] * coef[c] + G[(int64_t)i * 2 * C + C + c] * coef[C + c] + sd[i] * coef[2 * C + c];
}

// ------------------------------------------------------------------------------------------------ BN apply + ReLU (+pool, +dropout)
// thread layout inside a block: (256/cv pixel slots) x cv channel groups; a thread keeps its 8 channels' scale/shift in
// registers and strides over pixels (POOL: over 2x2 windows, four 16-byte loads in flight per window).
template <typename T, bool POOL>
__global__ void __launch_bounds__(256)
bn_act_kernel(const T* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift, int relu,
              T* __restrict__ y, int64_t ldy, T* __restrict__ pooled, int N, int H, int W, int C, DropArgs dp) {
  pdl_enter();
  const int cv = C >> 3;
  const int slots = blockDim.x / cv;
  const int c0 = (threadIdx.x % cv) << 3;
  const int slot = threadIdx.x / cv;
  float sc[8], sh[8];
  load8(scale + c0, sc); load8(shift + c0, sh);
  const uint32_t seed = dp.on ? drop_seed(dp) : 0u;
  const int64_t stride = (int64_t)gridDim.x * slots;
  if (POOL) {
    const int W2 = W >> 1, H2 = H >> 1;
    const int64_t nwin = (int64_t)N * H2 * W2;
    for (int64_t wi = blockIdx.x * (int64_t)slots + slot; wi < nwin; wi += stride) {
      const int j2 = (int)(wi % W2); const int64_t t = wi / W2;
      const int i2 = (int)(t % H2); const int64_t n = t / H2;
      const int64_t pix0 = (n * H + 2 * i2) * (int64_t)W + 2 * j2;
      float v[4][8], m[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) load8(z + (pix0 + (q >> 1) * W + (q & 1)) * C + c0, v[q]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int64_t pix = pix0 + (q >> 1) * W + (q & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = fmaf(v[q][j], sc[j], sh[j]);
          if (relu) a = fmaxf(a, 0.f);
          a = round_to<T>(a);
          v[q][j] = a;
          m[j] = q == 0 ? a : fmaxf(m[j], a);
        }
        if (dp.on) dropout_apply(v[q], (uint64_t)pix * dp.ctot + dp.c0 + c0, seed, dp.keep, dp.inv_keep);
        store8(y + pix * ldy + c0, v[q]);
      }
      store8(pooled + wi * C + c0, m);
    }
  } else {
    const int64_t npix = (int64_t)N * H * W;
    constexpr int U = 4;
    for (int64_t p0 = blockIdx.x * (int64_t)slots + slot; p0 < npix; p0 += stride * U) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) if (p0 + u * stride < npix) load8(z + (p0 + u * stride) * C + c0, v[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t pix = p0 + u * stride;
        if (pix >= npix) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[u][j] = fmaf(v[u][j], sc[j], sh[j]);
          if (relu) v[u][j] = fmaxf(v[u][j], 0.f);
        }
        if (dp.on) dropout_apply(v[u], (uint64_t)pix * dp.ctot + dp.c0 + c0, seed, dp.keep, dp.inv_keep);
        store8(y + pix * ldy + c0, v[u]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ BN backward
// thread layout inside a block: (rows = 256/cv) x cv, each thread owns 8 channels and strides over rows, kUnroll rows
// per trip with all loads issued before the first use (2*kUnroll 16-byte loads in flight per thread).
constexpr int kBnUnroll = 4;

template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 a; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void ldraw(const __nv_bfloat16* p, Raw8<__nv_bfloat16>& r) { r.a = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void ldraw(const float* p, Raw8<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void unraw(const Raw8<__nv_bfloat16>& r, float (&v)[8]) {
  const uint32_t u[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(u[i] << 16); v[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}
__device__ __forceinline__ void unraw(const Raw8<float>& r, float (&v)[8]) {
  v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ z,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t M, int C, int relu, DropArgs dp) {
  pdl_enter();
  extern __shared__ float s_red[];   // [2][C]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int cv = C >> 3;
  const int rows_per_block = blockDim.x / cv;
  const int c0 = (threadIdx.x % cv) << 3;
  const int r_in = threadIdx.x / cv;
  float sc[8], sh[8], mu[8], rs[8];
  const bool norm = mean != nullptr;
  if (norm) { load8(scale + c0, sc); load8(shift + c0, sh); load8(mean + c0, mu); load8(rstd + c0, rs); }
  const uint32_t seed = dp.on ? drop_seed(dp) : 0u;
  float sg[8], sgx[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sg[j] = 0.f; sgx[j] = 0.f; }
  const int64_t stride = (int64_t)gridDim.x * rows_per_block;
  for (int64_t r0 = blockIdx.x * (int64_t)rows_per_block + r_in; r0 < M; r0 += stride * kBnUnroll) {
    Raw8<T> rg[kBnUnroll], rz[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < M) { ldraw(dy + r * lddy + c0, rg[u]); ldraw(z + r * C + c0, rz[u]); }
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= M) break;
      float g[8], zz[8];
      unraw(rg[u], g); unraw(rz[u], zz);
      if (dp.on) {
        const uint64_t base = (uint64_t)r * dp.ctot + dp.c0 + c0;
        dropout_apply(g, base, seed, dp.keep, dp.inv_keep);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float act = norm ? fmaf(zz[j], sc[j], sh[j]) : zz[j];
        const float gj = (relu && !(act > 0.f)) ? 0.f : g[j];
        sg[j] += gj;
        if (norm) sgx[j] = fmaf(gj, (zz[j] - mu[j]) * rs[j], sgx[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&s_red[c0 + j], sg[j]);
    if (norm) atomicAdd(&s_red[C + c0 + j], sgx[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(&dbeta[i], s_red[i]);
    if (norm && dgamma) atomicAdd(&dgamma[i], s_red[C + i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dy, int64_t lddy, const T* __restrict__ z,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, T* __restrict__ dz,
                    int64_t M, int C, int relu, float inv_m, DropArgs dp) {
  pdl_enter();
  const int cv = C >> 3;
  const int rows_per_block = blockDim.x / cv;
  const int c0 = (threadIdx.x % cv) << 3;
  const int r_in = threadIdx.x / cv;
  const bool norm = mean != nullptr;
  // dz = sc*(g - dbeta/M - (zz-mu)*rs*dgamma/M) = sc*g - k1 - zz*k2  with  k2 = sc*rs*dgamma/M,  k1 = sc*dbeta/M - mu*k2
  float sc[8], sh[8], k1[8], k2[8];
  if (norm) {
    float mu[8], rs[8], dg[8], db[8];
    load8(scale + c0, sc); load8(shift + c0, sh); load8(mean + c0, mu); load8(rstd + c0, rs);
    load8(dgamma + c0, dg); load8(dbeta + c0, db);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      k2[j] = sc[j] * rs[j] * dg[j] * inv_m;
      k1[j] = sc[j] * db[j] * inv_m - mu[j] * k2[j];
    }
  }
  const uint32_t seed = dp.on ? drop_seed(dp) : 0u;
  const int64_t stride = (int64_t)gridDim.x * rows_per_block;
  for (int64_t r0 = blockIdx.x * (int64_t)rows_per_block + r_in; r0 < M; r0 += stride * kBnUnroll) {
    Raw8<T> rg[kBnUnroll], rz[kBnUnroll];
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < M) { ldraw(dy + r * lddy + c0, rg[u]); ldraw(z + r * C + c0, rz[u]); }
    }
#pragma unroll
    for (int u = 0; u < kBnUnroll; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= M) break;
      float g[8], zz[8], o[8];
      unraw(rg[u], g); unraw(rz[u], zz);
      if (dp.on) {
        const uint64_t base = (uint64_t)r * dp.ctot + dp.c0 + c0;
        dropout_apply(g, base, seed, dp.keep, dp.inv_keep);
      }
      if (norm) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float act = fmaf(zz[j], sc[j], sh[j]);
          const float gj = (relu && !(act > 0.f)) ? 0.f : g[j];
          o[j] = fmaf(sc[j], gj, -fmaf(zz[j], k2[j], k1[j]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (relu && !(zz[j] > 0.f)) ? 0.f : g[j];
      }
      store8(dz + r * C + c0, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ max pooling
template <typename T>
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const T* __restrict__ x, int64_t ldx, T* __restrict__ y, int N, int H, int W, int C) {
  pdl_enter();
  const int cv = C >> 3, W2 = W >> 1, H2 = H >> 1;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int c0 = (int)(t % cv) << 3; t /= cv;
  const int j2 = (int)(t % W2); t /= W2;
  const int i2 = (int)(t % H2); const int64_t n = t / H2;
  if (n >= N) return;
  float m[8];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t pix = (n * H + (2 * i2 + (q >> 1))) * (int64_t)W + 2 * j2 + (q & 1);
    float v[8];
    load8(x + pix * ldx + c0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = q == 0 ? v[j] : fmaxf(m[j], v[j]);
  }
  store8(y + ((n * H2 + i2) * (int64_t)W2 + j2) * C + c0, m);
}

template <typename T, bool SUMS, bool DROP>
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const T* __restrict__ z, int64_t ldz, const float* __restrict__ scale, const float* __restrict__ shift,
                   const T* __restrict__ dpool, const T* __restrict__ dskip, int64_t lddskip, T* __restrict__ dy,
                   int N, int H, int W, int C, float* __restrict__ bn_sums, DropArgs drp) {
  pdl_enter();
  extern __shared__ float s_bn[];      // [2][C] when bn_sums
  const uint32_t seed = DROP ? drop_seed(drp) : 0u;
  if (SUMS) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_bn[i] = 0.f;
    __syncthreads();
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  const int cv = C >> 3, W2 = W >> 1, H2 = H >> 1;
  const int slots = blockDim.x / cv;
  const int c0 = (threadIdx.x % cv) << 3;
  const int slot = threadIdx.x / cv;
  float sc[8], sh[8];
  const bool aff = scale != nullptr;
  if (aff) { load8(scale + c0, sc); load8(shift + c0, sh); }
  const int64_t nwin = (int64_t)N * H2 * W2;
  const int64_t stride = (int64_t)gridDim.x * slots;
  for (int64_t wi = blockIdx.x * (int64_t)slots + slot; wi < nwin; wi += stride) {
    const int j2 = (int)(wi % W2); const int64_t t = wi / W2;
    const int i2 = (int)(t % H2); const int64_t n = t / H2;
    const int64_t pix0 = (n * H + 2 * i2) * (int64_t)W + 2 * j2;
    float v[4][8], o[4][8], dp[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {                       // nine 16-byte loads in flight
      const int64_t pix = pix0 + (q >> 1) * W + (q & 1);
      load8(z + pix * ldz + c0, v[q]);
      if (dskip) {
        load8(dskip + pix * lddskip + c0, o[q]);
        if (DROP) dropout_apply(o[q], (uint64_t)pix * drp.ctot + drp.c0 + c0, seed, drp.keep, drp.inv_keep);   // skip tensor was stored dropped-out
      }
    }
    load8(dpool + wi * C + c0, dp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // first maximum in window scan order (TF CPU convention)
      float a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = aff ? round_to<T>(fmaxf(fmaf(v[q][j], sc[j], sh[j]), 0.f)) : v[q][j];
      int arg = 0; float m = a[0];
#pragma unroll
      for (int q = 1; q < 4; ++q) if (a[q] > m) { m = a[q]; arg = q; }
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][j] = (dskip ? o[q][j] : 0.f) + (arg == q ? dp[j] : 0.f);
      if (SUMS) {                     // gradient w.r.t. the BatchNormalization output + its two backward reductions
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float g = a[q] > 0.f ? round_to<T>(o[q][j]) : 0.f;
          o[q][j] = g; s1[j] += g; s2[j] = fmaf(g, a[q], s2[j]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) store8(dy + (pix0 + (q >> 1) * W + (q & 1)) * C + c0, o[q]);
  }
  if (SUMS) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&s_bn[c0 + j], s1[j]); atomicAdd(&s_bn[C + c0 + j], s2[j]); }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&bn_sums[i], s_bn[i]);
  }
}

// ------------------------------------------------------------------------------------------------ convT backward gather
// One block item = one input-image row (n, i): the two upsampled rows 2i and 2i+1 are read as contiguous runs and each
// 16-byte channel group goes to column block (a,b) of G row (n,i,j) — no division in the inner loop, 32-bit offsets.
template <typename T>
__global__ void __launch_bounds__(256)
convt_bwd_gather_kernel(const T* __restrict__ du, int64_t lddu, T* __restrict__ g, float* __restrict__ dbias,
                        int N, int H, int W, int Cout, DropArgs dp) {
  pdl_enter();
  extern __shared__ float s_red[];   // [Cout]
  const uint32_t seed = dp.on ? drop_seed(dp) : 0u;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int cv = Cout >> 3;
  const int per_block = blockDim.x / cv;          // upsampled pixels handled per pass
  const int c0 = (threadIdx.x % cv) << 3;
  const int slot = threadIdx.x / cv;
  const int W2 = 2 * W;
  float sb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sb[j] = 0.f;
  const int64_t rows = (int64_t)N * H * 2;        // upsampled rows (n, 2i + a)
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const int a = (int)(r & 1); const int64_t q = r >> 1;          // q = n*H + i
    const T* src_row = du + (r * W2) * lddu + c0;                  // upsampled row index == r because rows are (n, 2i+a) in order
    T* dst_row = g + (q * W) * (int64_t)(4 * Cout) + (int64_t)(a * 2) * Cout + c0;
    const uint64_t drop_row = (uint64_t)(r * W2) * (uint64_t)dp.ctot + dp.c0 + c0;
    for (int x2 = slot; x2 < W2; x2 += per_block) {                // upsampled column 2j + b
      float v[8];
      load8(src_row + (int64_t)x2 * lddu, v);
      if (dp.on) dropout_apply(v, drop_row + (uint64_t)x2 * (uint64_t)dp.ctot, seed, dp.keep, dp.inv_keep);   // du was stored dropped-out
      store8(dst_row + (int64_t)(x2 >> 1) * (4 * Cout) + (x2 & 1) * Cout, v);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) sb[jj] += v[jj];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&s_red[c0 + j], sb[j]);
  __syncthreads();
  if (dbias)
    for (int i = threadIdx.x; i < Cout; i += blockDim.x) atomicAdd(&dbias[i], s_red[i]);
}

// ------------------------------------------------------------------------------------------------ AdamW (Keras form)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             int64_t n, const float* __restrict__ hyper) {
  pdl_enter();
  const float lr = hyper[0], wd = hyper[1], b1 = hyper[2], b2 = hyper[3], eps = hyper[4], t = hyper[5], gs = hyper[6];
  // alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const float alpha = lr * sqrtf(1.f - powf(b2, t)) / (1.f - powf(b1, t));
  const float decay = 1.f - lr * wd;
  const int64_t i4 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 ww = *reinterpret_cast<float4*>(w + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 mm = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float* pw = &ww.x; const float* pg = &gg.x; float* pm = &mm.x; float* pv = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = pg[j] * gs;
      float wj = pw[j] * decay;
      pm[j] = b1 * pm[j] + (1.f - b1) * gj;
      pv[j] = b2 * pv[j] + (1.f - b2) * gj * gj;
      wj -= alpha * pm[j] / (sqrtf(pv[j]) + eps);
      pw[j] = wj;
    }
    *reinterpret_cast<float4*>(w + i4) = ww;
    *reinterpret_cast<float4*>(m + i4) = mm;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (int64_t i = i4; i < n; ++i) {
      const float gj = g[i] * gs;
      float wj = w[i] * decay;
      const float mj = b1 * m[i] + (1.f - b1) * gj;
      const float vj = b2 * v[i] + (1.f - b2) * gj * gj;
      wj -= alpha * mj / (sqrtf(vj) + eps);
      w[i] = wj; m[i] = mj; v[i] = vj;
    }
  }
}

__global__ void step_advance_kernel(float* hyper, uint32_t* counter) {
  pdl_enter();
  if (hyper) hyper[5] += 1.f;
  if (counter) *counter += 1u;
}

// ------------------------------------------------------------------------------------------------ casts
__global__ void cast_transpose_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                           __nv_bfloat16* __restrict__ dst_t, int R, int C, const float* __restrict__ col_scale) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = blockIdx.y * 32 + i;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[(int64_t)r * C + c];
      if (col_scale) v *= col_scale[c];
      if (dst) dst[(int64_t)r * C + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (!dst_t) return;
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < R && c2 < C) dst_t[(int64_t)c2 * R + r2] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// All dense kernels of the model in ONE launch (the per-step weight staging of the bf16 engine was 22 launches of a few
// microseconds each).  `table` (DEVICE, 6 int64 per matrix): {offset of src in `base` (floats), dst address, dst_t address, R, C,
// first tile}; a block finds its matrix by its tile index (tables are short: linear scan) and then works like the kernel above.
__global__ void cast_transpose_bf16_batched_kernel(const float* __restrict__ base, const int64_t* __restrict__ table, int n) {
  pdl_enter();
  __shared__ float tile[32][33];
  int k = 0;
  while (k + 1 < n && (int64_t)blockIdx.x >= table[(k + 1) * 6 + 5]) ++k;
  const int64_t* e = table + k * 6;
  const float* src = base + e[0];
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e[1]);
  __nv_bfloat16* dst_t = reinterpret_cast<__nv_bfloat16*>(e[2]);
  const int R = (int)e[3], C = (int)e[4];
  const int t = (int)((int64_t)blockIdx.x - e[5]);
  const int tiles_c = (C + 31) / 32;
  const int bx = t % tiles_c, by = t / tiles_c;
  const int c = bx * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by * 32 + i;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[(int64_t)r * C + c];
      if (dst) dst[(int64_t)r * C + c] = __float2bfloat16_rn(v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  if (!dst_t) return;
  const int r2 = by * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c2 = bx * 32 + i;
    if (r2 < R && c2 < C) dst_t[(int64_t)c2 * R + r2] = __float2bfloat16_rn(tile[threadIdx.x][i]);
  }
}

// tf32 (hi, lo) split (fp32 mode on the tensor cores).  tcgen05.mma.kind::tf32 reads an fp32 word and IGNORES its low 13 mantissa
// bits (measured: tools/tf32_trunc_probe.py), so the fp32 tensor itself serves as the `hi` operand (hi = trunc13(x)) and only
// lo = x - trunc13(x) has to be materialised; lo is rounded to nearest tf32 here so that the hardware's truncation of it does
// not bias every product toward zero.  `hi` (optional) receives x unchanged — needed only when the split also transposes.
__device__ __forceinline__ float tf32_lo(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);     // exact
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(r));
  return __uint_as_float(u);
}
__global__ void split_tf32_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int64_t cols,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  pdl_enter();
  const int64_t cv = cols >> 2;                   // cols % 4 == 0 (checked on the host)
  const int64_t n4 = rows * cv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cv, c = (i - r * cv) << 2;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + r * ld + c));
    if (hi) *reinterpret_cast<float4*>(hi + r * cols + c) = v;
    *reinterpret_cast<float4*>(lo + r * cols + c) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}
__global__ void split_tf32_t_kernel(const float* __restrict__ src, int64_t ld, int R, int C, float* __restrict__ hi_t, float* __restrict__ lo_t) {
  pdl_enter();
  __shared__ float tile[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = blockIdx.y * 32 + i;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c2 = blockIdx.x * 32 + i;
    if (r2 < R && c2 < C) {
      const float v = tile[threadIdx.x][i];
      if (hi_t) hi_t[(int64_t)c2 * R + r2] = v;
      lo_t[(int64_t)c2 * R + r2] = tf32_lo(v);
    }
  }
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ s, D* __restrict__ d, int64_t n) {
  pdl_enter();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = from_f32<D>(to_f32(s[i]));
}

// ------------------------------------------------------------------------------------------------ MeanIoU confusion counts
__global__ void confusion_kernel(const float* __restrict__ yt, const float* __restrict__ yp, int64_t n, int C,
                                 unsigned long long* __restrict__ counts, int use_thr, float thr) {
  pdl_enter();
  extern __shared__ unsigned int s_cnt[];   // [C*C]
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) s_cnt[i] = 0u;
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)yt[i];                                   // tf.cast(..., int64): truncation toward zero
    const int p = use_thr ? (yp[i] > thr ? 1 : 0) : (int)yp[i];
    if (t >= 0 && t < C && p >= 0 && p < C) atomicAdd(&s_cnt[t * C + p], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
}

// per-sample 2x2 confusion counts of thresholded probabilities (scripts/benchmark.py:159-170,260): counts[nb][t*2+p]
__global__ void sample_confusion_thr_kernel(const float* __restrict__ yt, const float* __restrict__ yp, int64_t per_sample,
                                            unsigned long long* __restrict__ counts, float thr) {
  pdl_enter();
  __shared__ unsigned int s_cnt[4];
  if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t nb = blockIdx.y;
  const float* t = yt + nb * per_sample;
  const float* p = yp + nb * per_sample;
  unsigned c[4] = {0u, 0u, 0u, 0u};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    const int tv = (int)t[i];
    const int pv = p[i] > thr ? 1 : 0;
    if (tv >= 0 && tv < 2) c[tv * 2 + pv] += 1u;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned v = c[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_cnt[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 4 && s_cnt[threadIdx.x]) atomicAdd(&counts[nb * 4 + threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ (I,T,P) sums
// sums[nb][c][0..2] += (sum t*p, sum t, sum p) over hw pixels.  blockDim is a multiple of C so that a thread's class is fixed.
__global__ void seg_sums_kernel(const float* __restrict__ yt, const float* __restrict__ yp, double* __restrict__ sums,
                                int64_t hw, int C) {
  pdl_enter();
  extern __shared__ double s_sum[];   // [C][3]
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) s_sum[i] = 0.0;
  __syncthreads();
  const int64_t nb = blockIdx.y;
  const int64_t per_img = hw * C;
  const float* t = yt + nb * per_img;
  const float* p = yp + nb * per_img;
  const int c = threadIdx.x % C;
  float si = 0.f, st = 0.f, sp = 0.f;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < per_img; e += (int64_t)gridDim.x * blockDim.x) {
    const float tv = t[e], pv = p[e];
    si = fmaf(tv, pv, si); st += tv; sp += pv;
  }
  atomicAdd(&s_sum[c * 3 + 0], (double)si);
  atomicAdd(&s_sum[c * 3 + 1], (double)st);
  atomicAdd(&s_sum[c * 3 + 2], (double)sp);
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) atomicAdd(&sums[nb * 3 * C + i], s_sum[i]);
}

// loss finalize: one block
__global__ void seg_loss_finalize_kernel(const double* __restrict__ sums, int npairs, float smooth, int kind,
                                         float grad_scale, float* __restrict__ out3, float* __restrict__ coef) {
  pdl_enter();
  __shared__ double s_d[32], s_i[32];
  double dsum = 0.0, isum = 0.0;
  const double s = (double)smooth, inv = 1.0 / (double)npairs;
  for (int i = threadIdx.x; i < npairs; i += blockDim.x) {
    const double I = sums[3 * i], T = sums[3 * i + 1], P = sums[3 * i + 2];
    const double D = T + P + s, num = 2.0 * I + s;
    const double U = T + P - I + s;
    dsum += num / D;
    isum += (I + s) / U;
    if (coef) {
      double ca, cb;
      if (kind == 0) { ca = -2.0 * inv / D; cb = num * inv / (D * D); }
      else           { ca = -inv * (U + (I + s)) / (U * U); cb = inv * (I + s) / (U * U); }
      coef[2 * i] = (float)(ca * grad_scale);
      coef[2 * i + 1] = (float)(cb * grad_scale);
    }
  }
  dsum = warp_sum(dsum); isum = warp_sum(isum);
  if ((threadIdx.x & 31) == 0) { s_d[threadIdx.x >> 5] = dsum; s_i[threadIdx.x >> 5] = isum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double d = 0.0, i = 0.0;
    for (int k = 0; k < (blockDim.x >> 5); ++k) { d += s_d[k]; i += s_i[k]; }
    d *= inv; i *= inv;
    out3[0] = (float)(1.0 - (kind == 0 ? d : i));
    out3[1] = (float)d;
    out3[2] = (float)i;
  }
}

}  // namespace unet

using namespace unet;
#define ST ((cudaStream_t)stream)

extern "C" int unet_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                            float* scale, float* shift, int C, void* stream) {
  UNET_REQUIRE(mean && var && scale && shift && C > 0, UNET_EINVAL, "bn_fold: bad argument");
  launch_pdl(bn_fold_kernel, grid_for(C, 128), 128, 0, ST, gamma, beta, mean, var, eps, scale, shift, C);
  UNET_LAUNCH_CHECK("bn_fold");
  return UNET_OK;
}

extern "C" int unet_bn_finalize(const double* colsum, const double* colsq, int64_t count,
                                const float* gamma, const float* beta, float eps, float momentum,
                                float* moving_mean, float* moving_var,
                                float* scale, float* shift, float* save_mean, float* save_rstd, int C, void* stream) {
  UNET_REQUIRE(colsum && colsq && scale && shift && C > 0 && count > 0, UNET_EINVAL, "bn_finalize: bad argument");
  launch_pdl(bn_finalize_kernel, grid_for(C, 128), 128, 0, ST, colsum, colsq, 1.0 / (double)count, gamma, beta, eps, momentum,
                                                      moving_mean, moving_var, scale, shift, save_mean, save_rstd, C);
  UNET_LAUNCH_CHECK("bn_finalize");
  return UNET_OK;
}

template <typename T>
static int bn_act_launch(const void* z, const float* scale, const float* shift, int relu, void* y, int64_t ldy,
                         void* pooled, int N, int H, int W, int C, DropArgs dp, cudaStream_t st) {
  const int cv = C / 8, slots = 256 / cv;
  if (pooled) {
    const int64_t items = (int64_t)N * (H / 2) * (W / 2);
    const unsigned grid = (unsigned)i64min(ceil_div(items, slots), (int64_t)sm_count() * 16);
    launch_pdl(bn_act_kernel<T, true>, grid, 256, 0, st, (const T*)z, scale, shift, relu, (T*)y, ldy, (T*)pooled, N, H, W, C, dp);
  } else {
    const int64_t items = (int64_t)N * H * W;
    const unsigned grid = (unsigned)i64min(ceil_div(items, (int64_t)slots * 4), (int64_t)sm_count() * 16);
    launch_pdl(bn_act_kernel<T, false>, grid, 256, 0, st, (const T*)z, scale, shift, relu, (T*)y, ldy, nullptr, N, H, W, C, dp);
  }
  UNET_LAUNCH_CHECK("bn_act");
  return UNET_OK;
}

extern "C" int unet_bn_act(const void* z, const float* scale, const float* shift, int relu,
                           void* y, int64_t ldy, void* pooled,
                           int N, int H, int W, int C, int dtype, const unet_dropout* drop, void* stream) {
  UNET_REQUIRE(z && scale && shift && y, UNET_EINVAL, "bn_act: null pointer");
  UNET_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ldy >= C, UNET_EINVAL, "bn_act: bad dims");
  UNET_REQUIRE(C % 8 == 0 && ldy % 8 == 0 && aligned16(z) && aligned16(y), UNET_EALIGN, "bn_act: needs C%%8==0, ld%%8==0, 16B pointers");
  UNET_REQUIRE(!pooled || (H % 2 == 0 && W % 2 == 0), UNET_EINVAL, "bn_act: pooling needs even H,W");
  UNET_REQUIRE(256 % (C / 8) == 0, UNET_EUNSUPPORTED, "bn_act: C/8 must divide 256 (C=%d)", C);
  const DropArgs dp = make_drop(drop);
  if (dtype == UNET_F32)  return bn_act_launch<float>(z, scale, shift, relu, y, ldy, pooled, N, H, W, C, dp, ST);
  if (dtype == UNET_BF16) return bn_act_launch<__nv_bfloat16>(z, scale, shift, relu, y, ldy, pooled, N, H, W, C, dp, ST);
  return set_error(UNET_EINVAL, "bn_act: bad dtype %d", dtype);
}

static int bn_bwd_check(const char* who, const void* dy, int64_t lddy, const void* z, int64_t M, int C) {
  UNET_REQUIRE(dy && z && M > 0 && C > 0 && lddy >= C, UNET_EINVAL, "%s: bad argument", who);
  UNET_REQUIRE(C % 8 == 0 && lddy % 8 == 0 && aligned16(dy) && aligned16(z), UNET_EALIGN, "%s: needs C%%8==0, ld%%8==0", who);
  UNET_REQUIRE(256 % (C / 8) == 0, UNET_EUNSUPPORTED, "%s: C/8 must divide 256 (C=%d)", who, C);
  return UNET_OK;
}

extern "C" int unet_bn_bwd_reduce(const void* dy, int64_t lddy, const void* z,
                                  const float* scale, const float* shift, const float* save_mean, const float* save_rstd,
                                  float* dgamma, float* dbeta, int64_t M, int C, int dtype, int relu,
                                  const unet_dropout* drop, void* stream) {
  const DropArgs dp = make_drop(drop);
  if (int e = bn_bwd_check("bn_bwd_reduce", dy, lddy, z, M, C)) return e;
  UNET_REQUIRE(dbeta, UNET_EINVAL, "bn_bwd_reduce: dbeta is null");
  UNET_REQUIRE(!save_mean || (scale && shift && save_rstd && dgamma), UNET_EINVAL, "bn_bwd_reduce: incomplete BN state");
  const int rows_per_block = 256 / (C / 8);
  const unsigned grid = (unsigned)i64min(ceil_div(M, (int64_t)rows_per_block * kBnUnroll), (int64_t)sm_count() * 8);
  const size_t smem = (size_t)2 * C * sizeof(float);
  if (dtype == UNET_F32)
    launch_pdl(bn_bwd_reduce_kernel<float>, grid, 256, smem, ST, (const float*)dy, lddy, (const float*)z, scale, shift, save_mean,
                                                        save_rstd, dgamma, dbeta, M, C, relu, dp);
  else if (dtype == UNET_BF16)
    launch_pdl(bn_bwd_reduce_kernel<__nv_bfloat16>, grid, 256, smem, ST, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z,
                                                                scale, shift, save_mean, save_rstd, dgamma, dbeta, M, C, relu, dp);
  else return set_error(UNET_EINVAL, "bn_bwd_reduce: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("bn_bwd_reduce");
  return UNET_OK;
}

extern "C" int unet_bn_bwd_apply(const void* dy, int64_t lddy, const void* z,
                                 const float* scale, const float* shift, const float* save_mean, const float* save_rstd,
                                 const float* dgamma, const float* dbeta, void* dz,
                                 int64_t M, int C, int dtype, int relu, const unet_dropout* drop, void* stream) {
  if (int e = bn_bwd_check("bn_bwd_apply", dy, lddy, z, M, C)) return e;
  const DropArgs dp = make_drop(drop);
  UNET_REQUIRE(dz && aligned16(dz), UNET_EINVAL, "bn_bwd_apply: dz null or unaligned");
  UNET_REQUIRE(!save_mean || (scale && shift && save_rstd && dgamma && dbeta), UNET_EINVAL, "bn_bwd_apply: incomplete BN state");
  const float inv_m = (float)(1.0 / (double)M);
  const int rows_per_block = 256 / (C / 8);
  const unsigned grid = (unsigned)i64min(ceil_div(M, (int64_t)rows_per_block * kBnUnroll), (int64_t)sm_count() * 16);
  if (dtype == UNET_F32)
    launch_pdl(bn_bwd_apply_kernel<float>, grid, 256, 0, ST, (const float*)dy, lddy, (const float*)z, scale, shift,
                                                    save_mean, save_rstd, dgamma, dbeta, (float*)dz, M, C, relu, inv_m, dp);
  else if (dtype == UNET_BF16)
    launch_pdl(bn_bwd_apply_kernel<__nv_bfloat16>, grid, 256, 0, ST, (const __nv_bfloat16*)dy, lddy, (const __nv_bfloat16*)z,
                                                            scale, shift, save_mean, save_rstd, dgamma, dbeta,
                                                            (__nv_bfloat16*)dz, M, C, relu, inv_m, dp);
  else return set_error(UNET_EINVAL, "bn_bwd_apply: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("bn_bwd_apply");
  return UNET_OK;
}

extern "C" int unet_maxpool2x2_fwd(const void* x, int64_t ldx, void* y, int N, int H, int W, int C, int dtype, void* stream) {
  UNET_REQUIRE(x && y && N > 0 && H > 1 && W > 1 && C > 0 && ldx >= C, UNET_EINVAL, "maxpool2x2_fwd: bad argument");
  UNET_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && aligned16(x) && aligned16(y), UNET_EALIGN, "maxpool2x2_fwd: needs C%%8==0, ld%%8==0");
  const int64_t threads = (int64_t)N * (H / 2) * (W / 2) * (C / 8);
  if (dtype == UNET_F32)
    launch_pdl(maxpool_fwd_kernel<float>, grid_for(threads), 256, 0, ST, (const float*)x, ldx, (float*)y, N, H, W, C);
  else if (dtype == UNET_BF16)
    launch_pdl(maxpool_fwd_kernel<__nv_bfloat16>, grid_for(threads), 256, 0, ST, (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, N, H, W, C);
  else return set_error(UNET_EINVAL, "maxpool2x2_fwd: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("maxpool2x2_fwd");
  return UNET_OK;
}

extern "C" int unet_maxpool2x2_bwd(const void* z, int64_t ldz, const float* scale, const float* shift,
                                   const void* dpool, const void* dskip, int64_t lddskip, void* dy,
                                   int N, int H, int W, int C, int dtype, float* bn_sums, const unet_dropout* skip_drop, void* stream) {
  const DropArgs dp = make_drop(skip_drop);
  UNET_REQUIRE(z && dpool && dy && N > 0 && H > 1 && W > 1 && C > 0 && ldz >= C, UNET_EINVAL, "maxpool2x2_bwd: bad argument");
  UNET_REQUIRE(H % 2 == 0 && W % 2 == 0, UNET_EINVAL, "maxpool2x2_bwd: needs even H,W");
  UNET_REQUIRE(C % 8 == 0 && ldz % 8 == 0 && (!dskip || lddskip % 8 == 0), UNET_EALIGN, "maxpool2x2_bwd: needs C%%8==0, ld%%8==0");
  UNET_REQUIRE((scale == nullptr) == (shift == nullptr), UNET_EINVAL, "maxpool2x2_bwd: scale/shift must come together");
  UNET_REQUIRE(256 % (C / 8) == 0, UNET_EUNSUPPORTED, "maxpool2x2_bwd: C/8 must divide 256 (C=%d)", C);
  const int64_t items = (int64_t)N * (H / 2) * (W / 2);
  const unsigned grid = (unsigned)i64min(ceil_div(items, 256 / (C / 8)), (int64_t)sm_count() * 16);
  UNET_REQUIRE(!bn_sums || scale, UNET_EINVAL, "maxpool2x2_bwd: bn_sums needs the activation recomputed from z (scale/shift)");
  const size_t smem = bn_sums ? (size_t)2 * C * sizeof(float) : 0;
#define UNET_MPB(T, S_, D_) launch_pdl(maxpool_bwd_kernel<T, S_, D_>, grid, 256, smem, ST, (const T*)z, ldz, scale, shift, (const T*)dpool, \
                                                                          (const T*)dskip, lddskip, (T*)dy, N, H, W, C, bn_sums, dp)
#define UNET_MPB4(T) do { if (bn_sums) { if (dp.on) UNET_MPB(T, true, true); else UNET_MPB(T, true, false); } \
                          else { if (dp.on) UNET_MPB(T, false, true); else UNET_MPB(T, false, false); } } while (0)
  if (dtype == UNET_F32) UNET_MPB4(float);
  else if (dtype == UNET_BF16) UNET_MPB4(__nv_bfloat16);
  else return set_error(UNET_EINVAL, "maxpool2x2_bwd: bad dtype %d", dtype);
#undef UNET_MPB4
#undef UNET_MPB
  UNET_LAUNCH_CHECK("maxpool2x2_bwd");
  return UNET_OK;
}

extern "C" int unet_convt_bwd_gather(const void* du, int64_t lddu, void* g, float* dbias,
                                     int N, int H, int W, int Cout, int dtype, const unet_dropout* drop, void* stream) {
  const DropArgs dp = make_drop(drop);
  UNET_REQUIRE(du && g && N > 0 && H > 0 && W > 0 && Cout > 0 && lddu >= Cout, UNET_EINVAL, "convt_bwd_gather: bad argument");
  UNET_REQUIRE(Cout % 8 == 0 && lddu % 8 == 0 && aligned16(du) && aligned16(g), UNET_EALIGN, "convt_bwd_gather: needs Cout%%8==0");
  UNET_REQUIRE(256 % (Cout / 8) == 0, UNET_EUNSUPPORTED, "convt_bwd_gather: Cout/8 must divide 256");
  const unsigned grid = (unsigned)i64min((int64_t)N * H * 2, (int64_t)sm_count() * 8);
  const size_t smem = (size_t)Cout * sizeof(float);
  if (dtype == UNET_F32)
    launch_pdl(convt_bwd_gather_kernel<float>, grid, 256, smem, ST, (const float*)du, lddu, (float*)g, dbias, N, H, W, Cout, dp);
  else if (dtype == UNET_BF16)
    launch_pdl(convt_bwd_gather_kernel<__nv_bfloat16>, grid, 256, smem, ST, (const __nv_bfloat16*)du, lddu, (__nv_bfloat16*)g, dbias, N, H, W, Cout, dp);
  else return set_error(UNET_EINVAL, "convt_bwd_gather: bad dtype %d", dtype);
  UNET_LAUNCH_CHECK("convt_bwd_gather");
  return UNET_OK;
}

extern "C" int unet_adamw_step(float* w, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream) {
  UNET_REQUIRE(w && g && m && v && hyper && n > 0, UNET_EINVAL, "adamw_step: bad argument");
  UNET_REQUIRE(aligned16(w) && aligned16(g) && aligned16(m) && aligned16(v), UNET_EALIGN, "adamw_step: buffers must be 16B aligned");
  launch_pdl(adamw_kernel, grid_for(ceil_div(n, 4)), 256, 0, ST, w, g, m, v, n, hyper);
  UNET_LAUNCH_CHECK("adamw_step");
  return UNET_OK;
}

extern "C" int unet_step_advance(float* hyper, uint32_t* counter, void* stream) {
  launch_pdl(step_advance_kernel, 1, 1, 0, ST, hyper, counter);
  UNET_LAUNCH_CHECK("step_advance");
  return UNET_OK;
}

extern "C" int unet_cast_transpose_bf16(const float* src, void* dst, void* dst_t, int R, int C, const float* col_scale, void* stream) {
  UNET_REQUIRE(src && (dst || dst_t) && R > 0 && C > 0, UNET_EINVAL, "cast_transpose_bf16: bad argument");
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)ceil_div(R, 32));
  launch_pdl(cast_transpose_bf16_kernel, grid, dim3(32, 8), 0, ST, src, (__nv_bfloat16*)dst, (__nv_bfloat16*)dst_t, R, C, col_scale);
  UNET_LAUNCH_CHECK("cast_transpose_bf16");
  return UNET_OK;
}

extern "C" int unet_cast_transpose_bf16_batched(const float* base, const int64_t* table, int n, int64_t total_tiles, void* stream) {
  UNET_REQUIRE(base && table && n > 0 && total_tiles > 0 && total_tiles < ((int64_t)1 << 31), UNET_EINVAL,
               "cast_transpose_bf16_batched: bad argument");
  launch_pdl(cast_transpose_bf16_batched_kernel, (unsigned)total_tiles, dim3(32, 8), 0, ST, base, table, n);
  UNET_LAUNCH_CHECK("cast_transpose_bf16_batched");
  return UNET_OK;
}

extern "C" int unet_split_tf32(const float* src, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo, int transpose, void* stream) {
  UNET_REQUIRE(src && lo && rows > 0 && cols > 0 && ld >= cols, UNET_EINVAL, "split_tf32: bad argument");
  if (transpose) {
    UNET_REQUIRE(rows < ((int64_t)1 << 31) && cols < ((int64_t)1 << 31), UNET_EUNSUPPORTED, "split_tf32: transposed split is for weight matrices");
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    launch_pdl(split_tf32_t_kernel, grid, dim3(32, 8), 0, ST, src, ld, (int)rows, (int)cols, hi, lo);
  } else {
    UNET_REQUIRE(cols % 4 == 0 && ld % 4 == 0 && aligned16(src) && (!hi || aligned16(hi)) && aligned16(lo), UNET_EALIGN,
                 "split_tf32: needs cols%%4==0, ld%%4==0 and 16B-aligned pointers");
    const unsigned grid = (unsigned)i64min(ceil_div(rows * (cols / 4), 256), (int64_t)sm_count() * 16);
    launch_pdl(split_tf32_kernel, grid, 256, 0, ST, src, ld, rows, cols, hi, lo);
  }
  UNET_LAUNCH_CHECK("split_tf32");
  return UNET_OK;
}

extern "C" int unet_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
  UNET_REQUIRE(src && dst && n > 0, UNET_EINVAL, "cast: bad argument");
  const unsigned grid = (unsigned)i64min(ceil_div(n, 256), (int64_t)sm_count() * 16);
  if (src_dtype == UNET_F32 && dst_dtype == UNET_BF16)
    launch_pdl(cast_kernel<float, __nv_bfloat16>, grid, 256, 0, ST, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == UNET_BF16 && dst_dtype == UNET_F32)
    launch_pdl(cast_kernel<__nv_bfloat16, float>, grid, 256, 0, ST, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == UNET_F32 && dst_dtype == UNET_F32)
    launch_pdl(cast_kernel<float, float>, grid, 256, 0, ST, (const float*)src, (float*)dst, n);
  else if (src_dtype == UNET_BF16 && dst_dtype == UNET_BF16)
    launch_pdl(cast_kernel<__nv_bfloat16, __nv_bfloat16>, grid, 256, 0, ST, (const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else return set_error(UNET_EINVAL, "cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
  UNET_LAUNCH_CHECK("cast");
  return UNET_OK;
}

extern "C" int unet_confusion_matrix_update(const float* y_true, const float* y_pred, int64_t n, int num_classes,
                                            unsigned long long* counts, void* stream) {
  UNET_REQUIRE(y_true && y_pred && counts && n > 0, UNET_EINVAL, "confusion_matrix_update: bad argument");
  UNET_REQUIRE(num_classes >= 1 && num_classes <= 64, UNET_EUNSUPPORTED, "confusion_matrix_update: 1 <= num_classes <= 64");
  const unsigned grid = (unsigned)i64min(ceil_div(n, 256 * 8), (int64_t)sm_count() * 8);
  launch_pdl(confusion_kernel, grid, 256, (size_t)num_classes * num_classes * sizeof(unsigned), ST, y_true, y_pred, n, num_classes, counts, 0, 0.f);
  UNET_LAUNCH_CHECK("confusion_matrix_update");
  return UNET_OK;
}

extern "C" int unet_confusion_matrix_update_thr(const float* y_true, const float* prob, float thr, int64_t n,
                                                unsigned long long* counts, void* stream) {
  UNET_REQUIRE(y_true && prob && counts && n > 0, UNET_EINVAL, "confusion_matrix_update_thr: bad argument");
  const unsigned grid = (unsigned)i64min(ceil_div(n, 256 * 8), (int64_t)sm_count() * 8);
  launch_pdl(confusion_kernel, grid, 256, 4 * sizeof(unsigned), ST, y_true, prob, n, 2, counts, 1, thr);
  UNET_LAUNCH_CHECK("confusion_matrix_update_thr");
  return UNET_OK;
}

extern "C" int unet_sample_confusion_thr(const float* y_true, const float* prob, float thr, int64_t NB, int64_t per_sample,
                                         unsigned long long* counts, void* stream) {
  UNET_REQUIRE(y_true && prob && counts && NB > 0 && per_sample > 0, UNET_EINVAL, "sample_confusion_thr: bad argument");
  UNET_REQUIRE(NB <= 65535, UNET_EUNSUPPORTED, "sample_confusion_thr: NB <= 65535");
  const unsigned gx = (unsigned)i64max(1, i64min(ceil_div(per_sample, 256 * 8), ceil_div((int64_t)sm_count() * 8, NB)));
  launch_pdl(sample_confusion_thr_kernel, dim3(gx, (unsigned)NB), 256, 0, ST, y_true, prob, per_sample, counts, thr);
  UNET_LAUNCH_CHECK("sample_confusion_thr");
  return UNET_OK;
}

extern "C" int unet_seg_sums(const float* y_true, const float* y_pred, double* sums, int64_t NB, int64_t hw, int C, void* stream) {
  UNET_REQUIRE(y_true && y_pred && sums && NB > 0 && hw > 0 && C > 0, UNET_EINVAL, "seg_sums: bad argument");
  UNET_REQUIRE(C <= 256 && NB <= 65535, UNET_EUNSUPPORTED, "seg_sums: C <= 256, NB <= 65535");
  const int block = (256 / C) * C;
  const unsigned gx = (unsigned)i64max(1, i64min(ceil_div(hw * C, (int64_t)block * 16), ceil_div((int64_t)sm_count() * 8, NB)));
  launch_pdl(seg_sums_kernel, dim3(gx, (unsigned)NB), block, (size_t)3 * C * sizeof(double), ST, y_true, y_pred, sums, hw, C);
  UNET_LAUNCH_CHECK("seg_sums");
  return UNET_OK;
}

extern "C" int unet_seg_loss_finalize(const double* sums, int NC_pairs, float smooth, int kind, float grad_scale,
                                      float* out3, float* coef, void* stream) {
  UNET_REQUIRE(sums && out3 && NC_pairs > 0 && (kind == 0 || kind == 1), UNET_EINVAL, "seg_loss_finalize: bad argument");
  launch_pdl(seg_loss_finalize_kernel, 1, 256, 0, ST, sums, NC_pairs, smooth, kind, grad_scale, out3, coef);
  UNET_LAUNCH_CHECK("seg_loss_finalize");
  return UNET_OK;
}

extern "C" int unet_bn_bwd_coef(const float* sums, const float* gamma, const float* beta, const float* save_mean,
                                const float* save_rstd, int64_t count, float* dgamma, float* dbeta, float* coef,
                                const float* w, int Cin, int C, void* wab, float* bias, int* ill_conditioned, void* stream) {
  UNET_REQUIRE(sums && gamma && beta && save_mean && save_rstd && C > 0 && count > 0, UNET_EINVAL, "bn_bwd_coef: bad argument");
  UNET_REQUIRE(!w || (wab && bias && Cin > 0), UNET_EINVAL, "bn_bwd_coef: w needs wab, bias and Cin");
  launch_pdl(bn_bwd_coef_kernel, w ? Cin : 1, 256, 0, ST, sums, gamma, beta, save_mean, save_rstd, 1.f / (float)count, dgamma, dbeta, coef,
                                                 w, Cin, C, (__nv_bfloat16*)wab, bias, ill_conditioned);
  UNET_LAUNCH_CHECK("bn_bwd_coef");
  return UNET_OK;
}

extern "C" int unet_bn_bwd_wgrad_combine(const float* G, const float* coef, const float* sd, float* dw, int Cin, int C, void* stream) {
  UNET_REQUIRE(G && coef && sd && dw && Cin > 0 && C > 0, UNET_EINVAL, "bn_bwd_wgrad_combine: bad argument");
  launch_pdl(bn_bwd_wgrad_combine_kernel, grid_for((int64_t)Cin * C), 256, 0, ST, G, coef, sd, dw, Cin, C);
  UNET_LAUNCH_CHECK("bn_bwd_wgrad_combine");
  return UNET_OK;
}
