// Output head: Conv2D(num_classes, 1, activation=sigmoid|softmax) (reference model/u_net.py:105-112) fused with the
// per-(image, class) Dice/IoU sums of utils/metrics.py:29-31 (forward) and with the gradient of
// utils/loss.py:9-45 through the activation and the 1x1 convolution (backward).
//
// The GEMM is degenerate (N = 1 or 8 output channels, ~1 flop/B) so it runs on CUDA cores: 8 lanes share one
// pixel, each lane owns K/8 input channels with 16-byte loads, partial dot products are combined with 3 shuffles.
// Probabilities are always produced in fp32 from the fp32 accumulator (MeanIoU in train.py:231 truncates
// probabilities to int, so they must not be rounded through bf16).
#include "common.cuh"
#include "ptx.cuh"

namespace unet {

constexpr int kHeadMaxK = 512;

template <typename T, int MAXC>
__global__ void __launch_bounds__(256, MAXC == 1 ? 4 : 1)
head_fwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ probs, const float* __restrict__ y_true, double* __restrict__ sums,
                int64_t hw, int K, int C, int pix_per_block) {
  pdl_enter();
  __shared__ float s_w[kHeadMaxK * MAXC];
  __shared__ float s_b[MAXC];
  __shared__ double s_sum[MAXC * 3];
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) s_w[(i / C) * MAXC + (i % C)] = w[i];
  if (threadIdx.x < C) s_b[threadIdx.x] = b ? b[threadIdx.x] : 0.f;
  if (threadIdx.x < MAXC * 3) s_sum[threadIdx.x] = 0.0;
  __syncthreads();

  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7;          // which eighth of the channels
  const int slot = threadIdx.x >> 3;        // pixel slot inside the block (32 pixels per pass)
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);

  float si[MAXC], st[MAXC], sp[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) { si[c] = 0.f; st[c] = 0.f; sp[c] = 0.f; }

  constexpr int U = 4;                                  // pixel slots per trip: U independent 16-byte loads in flight
  for (int64_t p0 = p_begin; p0 < p_end; p0 += 32 * U) {
    float acc[U][MAXC];
    int64_t mrow[U];
    bool live[U];
    float tq[U];                                        // binary head: y_true of the trip's pixels, requested with the x loads
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * 32 + slot;
      live[u] = p < p_end;
      mrow[u] = n * hw + (live[u] ? p : p_begin);
      tq[u] = (MAXC == 1 && y_true && sub == 0) ? __ldg(y_true + mrow[u]) : 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c) acc[u][c] = 0.f;
    }
    for (int k0 = sub * 8; k0 < K; k0 += 64) {
      float v[U][8], wk[8][MAXC];
#pragma unroll
      for (int u = 0; u < U; ++u) load8(x + mrow[u] * ldx + k0, v[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < MAXC; ++c) wk[j][c] = s_w[(k0 + j) * MAXC + c];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int c = 0; c < MAXC; ++c) acc[u][c] = fmaf(v[u][j], wk[j][c], acc[u][c]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        acc[u][c] += __shfl_xor_sync(0xffffffffu, acc[u][c], 1);
        acc[u][c] += __shfl_xor_sync(0xffffffffu, acc[u][c], 2);
        acc[u][c] += __shfl_xor_sync(0xffffffffu, acc[u][c], 4);
      }
      if (live[u] && sub == 0) {
        const int64_t m = mrow[u];
        float pr[MAXC];
        if (C == 1) {
          pr[0] = 1.f / (1.f + expf(-(acc[u][0] + s_b[0])));
        } else {
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) if (c < C) { acc[u][c] += s_b[c]; mx = fmaxf(mx, acc[u][c]); }
          float den = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) if (c < C) { pr[c] = expf(acc[u][c] - mx); den += pr[c]; }
          const float inv = 1.f / den;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) if (c < C) pr[c] *= inv;
        }
#pragma unroll
        for (int c = 0; c < MAXC; ++c) if (c < C) {
          probs[m * C + c] = pr[c];
          if (y_true) {
            const float t = MAXC == 1 ? tq[u] : y_true[m * C + c];
            si[c] = fmaf(t, pr[c], si[c]); st[c] += t; sp[c] += pr[c];
          }
        }
      }
    }
  }
  if (y_true && sums) {
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
      if (c < C) {   // lanes with sub != 0 contribute zeros
        const float a = warp_sum(si[c]), bq = warp_sum(st[c]), cq = warp_sum(sp[c]);
        if ((threadIdx.x & 31) == 0) {
          atomicAdd(&s_sum[c * 3 + 0], (double)a);
          atomicAdd(&s_sum[c * 3 + 1], (double)bq);
          atomicAdd(&s_sum[c * 3 + 2], (double)cq);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < C * 3) atomicAdd(&sums[n * C * 3 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

template <typename T, int MAXC, bool SUMS>
__global__ void __launch_bounds__(256, MAXC == 1 ? 3 : 1)
head_bwd_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ probs,
                const float* __restrict__ y_true, const float* __restrict__ coef, T* __restrict__ dx, int64_t lddx,
                float* __restrict__ dw, float* __restrict__ db, int64_t hw, int K, int C, int pix_per_block,
                float* __restrict__ bn_sums) {
  pdl_enter();
  __shared__ float s_w[kHeadMaxK * MAXC];
  __shared__ float s_dw[kHeadMaxK * MAXC];
  __shared__ float s_db[MAXC];
  __shared__ float s_bn[SUMS ? 2 * kHeadMaxK : 1];
  if (SUMS) for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) s_bn[i] = 0.f;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) s_w[(i / C) * MAXC + (i % C)] = w[i];
  for (int i = threadIdx.x; i < K * MAXC; i += blockDim.x) s_dw[i] = 0.f;
  if (threadIdx.x < MAXC) s_db[threadIdx.x] = 0.f;
  __syncthreads();

  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7;
  const int slot = threadIdx.x >> 3;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);
  float ca[MAXC], cb[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    ca[c] = c < C ? coef[(n * C + c) * 2] : 0.f;
    cb[c] = c < C ? coef[(n * C + c) * 2 + 1] : 0.f;
  }
  // K == 64 in the reference model: one 8-channel group per lane.  Larger K loops.
  float dbs[MAXC];
#pragma unroll
  for (int c = 0; c < MAXC; ++c) dbs[c] = 0.f;

  for (int k0 = sub * 8; k0 < K; k0 += 64) {
    float dwacc[8][MAXC];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) dwacc[j][c] = 0.f;
    constexpr int U = 4;                                // pixel slots per trip, loads issued before use
    float wk[8][MAXC];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) wk[j][c] = s_w[(k0 + j) * MAXC + c];
    // SUMS: sum(g) needs its own accumulators; sum(g*x) = sum_c w[k,c] * dw[k,c] because x >= 0 is its own ReLU mask
    float s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = 0.f;
    for (int64_t p0 = p_begin; p0 < p_end; p0 += 32 * U) {
      float v[U][8], dz[U][MAXC];
      int64_t mrow[U];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t p = p0 + u * 32 + slot;
        live[u] = p < p_end;
        mrow[u] = n * hw + (live[u] ? p : p_begin);
        load8(x + mrow[u] * ldx + k0, v[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t m = mrow[u];
        if (C == 1) {
          const float pr = __ldg(probs + m), t = __ldg(y_true + m);
          dz[u][0] = live[u] ? fmaf(ca[0], t, cb[0]) * pr * (1.f - pr) : 0.f;
        } else {
          float pr[MAXC], gl[MAXC], dot = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) if (c < C) {
            pr[c] = __ldg(probs + m * C + c);
            gl[c] = fmaf(ca[c], __ldg(y_true + m * C + c), cb[c]);
            dot = fmaf(gl[c], pr[c], dot);
          }
#pragma unroll
          for (int c = 0; c < MAXC; ++c) dz[u][c] = (c < C && live[u]) ? pr[c] * (gl[c] - dot) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float sacc = 0.f;
#pragma unroll
          for (int c = 0; c < MAXC; ++c) {
            sacc = fmaf(dz[u][c], wk[j][c], sacc);
            dwacc[j][c] = fmaf(v[u][j], dz[u][c], dwacc[j][c]);
          }
          o[j] = sacc;
        }
        if (SUMS && live[u]) {       // x is the post-ReLU activation: mask, and BatchNormalization's two backward reductions
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float g = v[u][j] > 0.f ? o[j] : 0.f;
            o[j] = g; s1[j] += g;
          }
        }
        if (dx && live[u]) store8(dx + mrow[u] * lddx + k0, o);
        if (k0 == sub * 8 && sub == 0) {
#pragma unroll
          for (int c = 0; c < MAXC; ++c) dbs[c] += dz[u][c];
        }
      }
    }
    // lanes 8 apart share `sub`: fold them, then one shared atomic per (channel, class) per warp
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < MAXC; ++c) {
        float s = dwacc[j][c];
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if ((threadIdx.x & 31) < 8 && c < C) atomicAdd(&s_dw[(k0 + j) * MAXC + c], s);
      }
    if (SUMS) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = s1[j], b = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) b = fmaf(wk[j][c], dwacc[j][c], b);
        a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
        b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16);
        if ((threadIdx.x & 31) < 8) { atomicAdd(&s_bn[k0 + j], a); atomicAdd(&s_bn[K + k0 + j], b); }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < MAXC; ++c) {
    const float s = warp_sum(dbs[c]);
    if ((threadIdx.x & 31) == 0 && c < C) atomicAdd(&s_db[c], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) atomicAdd(&dw[i], s_dw[(i / C) * MAXC + (i % C)]);
  if (threadIdx.x < C) atomicAdd(&db[threadIdx.x], s_db[threadIdx.x]);
  if (SUMS) for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) atomicAdd(&bn_sums[i], s_bn[i]);
}

// ------------------------------------------------------------------------------------------------ binary head forward, streamed
// probs[m] = sigmoid(x[m,:] . w + b) and the per-image Dice sums (I, T, P), same cp.async ring as the backward kernel below.
constexpr int kH1FD = 8;
__global__ void __launch_bounds__(256, 4)
head1_fwd_stream_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                        float* __restrict__ probs, const float* __restrict__ y_true, double* __restrict__ sums,
                        int64_t hw, int pix_per_block, const float* __restrict__ x_scale, const float* __restrict__ x_shift) {
  pdl_enter();
  __shared__ uint4 ring_x[kH1FD][256];
  __shared__ float ring_t[kH1FD][256];
  __shared__ double s_sum[3];
  if (threadIdx.x < 3) s_sum[threadIdx.x] = 0.0;
  __syncthreads();
  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);
  float wk[8], xs[8], xt[8];
  load8(w + sub * 8, wk);
  const bool aff = x_scale != nullptr;             // x is dec1_block2's pre-BN tensor: BN + ReLU on load
  if (aff) { load8(x_scale + sub * 8, xs); load8(x_shift + sub * 8, xt); }
  const float bias = b ? __ldg(b) : 0.f;
  float si = 0.f, st_ = 0.f, sp_ = 0.f;
  const uint32_t sx = smem_u32(&ring_x[0][threadIdx.x]), stq = smem_u32(&ring_t[0][threadIdx.x]);
  const int64_t first = p_begin + slot;
  const int count = first < p_end ? (int)((p_end - first + 31) / 32) : 0;
  auto issue = [&](int i) {
    if (i < count) {
      const int64_t m = n * hw + first + (int64_t)i * 32;
      const int sl = i % kH1FD;
      cp_async16(sx + sl * (256 * 16), x + m * 64 + sub * 8);
      if (y_true && sub == 0) cp_async4(stq + sl * (256 * 4), y_true + m);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < kH1FD; ++i) issue(i);
  for (int i = 0; i < count; ++i) {
    cp_async_wait<kH1FD - 1>();
    const int sl = i % kH1FD;
    const uint4 raw = lds128u(sx + sl * (256 * 16));
    const float t = (y_true && sub == 0) ? lds32f(stq + sl * (256 * 4)) : 0.f;
    const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
    // packed FFMA2 (two channels per instruction); the even / odd channel partial sums are added before the lane reduction
    float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 v = make_float2(bf16lo_to_f32(u[q]), __uint_as_float(u[q] & 0xffff0000u));
      if (aff) {
        v = fma2(v, make_float2(xs[2 * q], xs[2 * q + 1]), make_float2(xt[2 * q], xt[2 * q + 1]));
        v = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
      }
      acc2 = fma2(v, make_float2(wk[2 * q], wk[2 * q + 1]), acc2);
    }
    float acc = acc2.x + acc2.y;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2); acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (sub == 0) {
      const float pr = 1.f / (1.f + expf(-(acc + bias)));
      probs[n * hw + first + (int64_t)i * 32] = pr;
      si = fmaf(t, pr, si); st_ += t; sp_ += pr;
    }
    issue(i + kH1FD);
  }
  cp_async_wait<0>();
  if (y_true && sums) {
    const float a = warp_sum(si), bq = warp_sum(st_), cq = warp_sum(sp_);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_sum[0], (double)a); atomicAdd(&s_sum[1], (double)bq); atomicAdd(&s_sum[2], (double)cq); }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&sums[n * 3 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

// ------------------------------------------------------------------------------------------------ binary head backward, streamed
// The reference configuration (num_classes = 1, 64 channels, bf16): dx[m,k] = dz[m]*w[k], dw[k] += sum_m x[m,k]*dz[m],
// db += sum dz, with dz = (ca*t + cb)*p*(1-p).  A thread owns 16 bytes of channels of one pixel slot and walks its pixels
// through a ring of D per-thread cp.async slots (x, p and t): D*24 bytes per thread are in flight with no registers held,
// which is what the register-staged kernel above lacks (it runs at 0.45 of the HBM peak).  SUMS: dx is ReLU-masked by
// x > 0 and sum(dx), sum(dx*x) = w[k]*dw[k] (x >= 0 is its own mask) are accumulated for the folded BatchNormalization backward.
constexpr int kH1D = 6;
template <bool SUMS>
__global__ void __launch_bounds__(256, 4)
head1_bwd_stream_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ probs,
                        const float* __restrict__ y_true, const float* __restrict__ coef, __nv_bfloat16* __restrict__ dx,
                        float* __restrict__ dw, float* __restrict__ db, int64_t hw, int pix_per_block, float* __restrict__ bn_sums,
                        const float* __restrict__ x_scale, const float* __restrict__ x_shift) {
  pdl_enter();
  __shared__ uint4 ring_x[kH1D][256];
  __shared__ float ring_p[kH1D][256], ring_t[kH1D][256];
  __shared__ float s_red[3 * 64 + 1];              // dw | sum g | (unused) | db
  for (int i = threadIdx.x; i < 3 * 64 + 1; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);
  const float ca = coef[n * 2], cb = coef[n * 2 + 1];
  float wk[8], dwacc[8], s1[8], xs[8], xt[8];
  load8(w + sub * 8, wk);
  const bool aff = x_scale != nullptr;             // x is dec1_block2's pre-BN tensor: BN + ReLU on load
  if (aff) { load8(x_scale + sub * 8, xs); load8(x_shift + sub * 8, xt); }
#pragma unroll
  for (int j = 0; j < 8; ++j) { dwacc[j] = 0.f; s1[j] = 0.f; }
  float dbs = 0.f;
  const uint32_t sx = smem_u32(&ring_x[0][threadIdx.x]), sp = smem_u32(&ring_p[0][threadIdx.x]), st = smem_u32(&ring_t[0][threadIdx.x]);
  const int64_t first = p_begin + slot;
  const int count = first < p_end ? (int)((p_end - first + 31) / 32) : 0;      // pixels of this thread
  auto issue = [&](int i) {
    if (i < count) {
      const int64_t m = n * hw + first + (int64_t)i * 32;
      const int sl = i % kH1D;
      cp_async16(sx + sl * (256 * 16), x + m * 64 + sub * 8);
      cp_async4(sp + sl * (256 * 4), probs + m);
      cp_async4(st + sl * (256 * 4), y_true + m);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int i = 0; i < kH1D; ++i) issue(i);
  for (int i = 0; i < count; ++i) {
    cp_async_wait<kH1D - 1>();
    const int sl = i % kH1D;
    const uint4 raw = lds128u(sx + sl * (256 * 16));
    const float pr = lds32f(sp + sl * (256 * 4)), t = lds32f(st + sl * (256 * 4));
    const int64_t m = n * hw + first + (int64_t)i * 32;
    const float dz = fmaf(ca, t, cb) * pr * (1.f - pr);
    const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
    float o[8];
    const float2 dz2 = make_float2(dz, dz);
#pragma unroll
    for (int q = 0; q < 4; ++q) {                   // packed FFMA2 / FMUL2 / FADD2: two channels per instruction
      float2 v = make_float2(bf16lo_to_f32(u[q]), __uint_as_float(u[q] & 0xffff0000u));
      if (aff) {
        v = fma2(v, make_float2(xs[2 * q], xs[2 * q + 1]), make_float2(xt[2 * q], xt[2 * q + 1]));
        v = make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
      }
      const float2 da = fma2(v, dz2, make_float2(dwacc[2 * q], dwacc[2 * q + 1]));
      dwacc[2 * q] = da.x; dwacc[2 * q + 1] = da.y;
      float2 g = mul2(dz2, make_float2(wk[2 * q], wk[2 * q + 1]));
      if (SUMS) {
        g.x = v.x > 0.f ? g.x : 0.f; g.y = v.y > 0.f ? g.y : 0.f;
        const float2 sa = add2(make_float2(s1[2 * q], s1[2 * q + 1]), g);
        s1[2 * q] = sa.x; s1[2 * q + 1] = sa.y;
      }
      o[2 * q] = g.x; o[2 * q + 1] = g.y;
    }
    store8(dx + m * 64 + sub * 8, o);
    if (sub == 0) dbs += dz;
    issue(i + kH1D);
  }
  cp_async_wait<0>();
  // lanes 8 apart share `sub`
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float a = dwacc[j], b = s1[j];
    a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
    if (SUMS) { b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16); }
    if ((threadIdx.x & 31) < 8) { atomicAdd(&s_red[sub * 8 + j], a); if (SUMS) atomicAdd(&s_red[64 + sub * 8 + j], b); }
  }
  dbs = warp_sum(dbs);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[192], dbs);
  __syncthreads();
  if (threadIdx.x < 64) {
    const float dwk = s_red[threadIdx.x];
    atomicAdd(&dw[threadIdx.x], dwk);
    if (SUMS) { atomicAdd(&bn_sums[threadIdx.x], s_red[64 + threadIdx.x]); atomicAdd(&bn_sums[64 + threadIdx.x], w[threadIdx.x] * dwk); }
  }
  if (threadIdx.x == 0) atomicAdd(&db[0], s_red[192]);
}

// ------------------------------------------------------------------------------------------------ multi-class head (2 <= C <= 8)
// Same 8-lanes-per-pixel layout, but after the contraction the 8 lanes of a pixel each OWN one class: partial logits are
// combined with a 7-shuffle reduce-scatter, softmax runs across the 8 lanes (max / sum by butterfly), and probabilities,
// y_true and the Dice sums are touched as one coalesced 32-byte segment per pixel.
// weight row r (8 class slots, 32 B) lives at float offset r*8 + (r/8)*4: the 8 lanes of a pixel read rows 8 apart, and the
// 16-byte pad per 8 rows spreads them over all banks (conflict-free 16-byte loads)
__device__ __forceinline__ int mc_row(int r) { return r * 8 + ((r >> 3) << 2); }
constexpr int kMcWeightFloats = kHeadMaxK * 8 + (kHeadMaxK / 8) * 4;

template <typename T>
__global__ void __launch_bounds__(256, 2)
head_fwd_mc_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ probs, const float* __restrict__ y_true, double* __restrict__ sums,
                   int64_t hw, int K, int C, int pix_per_block) {
  pdl_enter();
  __shared__ __align__(16) float s_w[kMcWeightFloats];
  __shared__ double s_sum[8 * 3];
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) s_w[mc_row(i >> 3) + (i & 7)] = (i & 7) < C ? w[(i >> 3) * C + (i & 7)] : 0.f;
  if (threadIdx.x < 24) s_sum[threadIdx.x] = 0.0;
  __syncthreads();
  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const float bias = (sub < C && b) ? b[sub] : 0.f;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);
  float si = 0.f, st = 0.f, sp = 0.f;
  constexpr int U = 4;                    // pixels per thread in flight: 4 x 16-byte loads before the first FMA
  for (int64_t p0 = p_begin; p0 < p_end; p0 += 32 * U) {
    float2 acc2[U][4];                     // class pairs: every FMA below is a packed FFMA2
    int64_t mrow[U]; bool live[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t p = p0 + u * 32 + slot;
      live[u] = p < p_end;
      mrow[u] = n * hw + (live[u] ? p : p_begin);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc2[u][c] = make_float2(0.f, 0.f);
    }
    for (int k0 = sub * 8; k0 < K; k0 += 64) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) load8(x + mrow[u] * ldx + k0, v[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[mc_row(k0 + j)]);
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[mc_row(k0 + j) + 4]);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const float2 vv = make_float2(v[u][j], v[u][j]);
          acc2[u][0] = fma2(vv, make_float2(w0.x, w0.y), acc2[u][0]); acc2[u][1] = fma2(vv, make_float2(w0.z, w0.w), acc2[u][1]);
          acc2[u][2] = fma2(vv, make_float2(w1.x, w1.y), acc2[u][2]); acc2[u][3] = fma2(vv, make_float2(w1.z, w1.w), acc2[u][3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float acc[1][8];
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[0][2 * c] = acc2[u][c].x; acc[0][2 * c + 1] = acc2[u][c].y; }
      // reduce-scatter over the 8 lanes of the pixel: lane `sub` ends with the full logit of class `sub`
      float r4[4], r2[2], logit;
      const bool hi4 = sub & 4, hi2 = sub & 2, hi1 = sub & 1;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = hi4 ? acc[0][i] : acc[0][i + 4];
        const float keep = hi4 ? acc[0][i + 4] : acc[0][i];
        r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = hi2 ? r4[i] : r4[i + 2];
        const float keep = hi2 ? r4[i + 2] : r4[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      {
        const float send = hi1 ? r2[0] : r2[1];
        const float keep = hi1 ? r2[1] : r2[0];
        logit = keep + __shfl_xor_sync(0xffffffffu, send, 1);
      }
      logit = sub < C ? logit + bias : -INFINITY;
      float mx = logit;
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
      const float e = sub < C ? expf(logit - mx) : 0.f;
      float den = e;
      den += __shfl_xor_sync(0xffffffffu, den, 1); den += __shfl_xor_sync(0xffffffffu, den, 2); den += __shfl_xor_sync(0xffffffffu, den, 4);
      const float pr = e / den;
      if (live[u] && sub < C) {
        probs[mrow[u] * C + sub] = pr;
        if (y_true) {
          const float t = __ldg(y_true + mrow[u] * C + sub);
          si = fmaf(t, pr, si); st += t; sp += pr;
        }
      }
    }
  }
  if (y_true && sums) {
    // lanes with equal `sub` own the same class
    si += __shfl_xor_sync(0xffffffffu, si, 8); si += __shfl_xor_sync(0xffffffffu, si, 16);
    st += __shfl_xor_sync(0xffffffffu, st, 8); st += __shfl_xor_sync(0xffffffffu, st, 16);
    sp += __shfl_xor_sync(0xffffffffu, sp, 8); sp += __shfl_xor_sync(0xffffffffu, sp, 16);
    if ((threadIdx.x & 31) < 8 && sub < C) {
      atomicAdd(&s_sum[sub * 3 + 0], (double)si); atomicAdd(&s_sum[sub * 3 + 1], (double)st); atomicAdd(&s_sum[sub * 3 + 2], (double)sp);
    }
    __syncthreads();
    if (threadIdx.x < C * 3) atomicAdd(&sums[n * C * 3 + threadIdx.x], s_sum[threadIdx.x]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 2)
head_bwd_mc_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ probs,
                   const float* __restrict__ y_true, const float* __restrict__ coef, T* __restrict__ dx, int64_t lddx,
                   float* __restrict__ dw, float* __restrict__ db, int64_t hw, int K, int C, int pix_per_block,
                   float* __restrict__ bn_sums) {
  pdl_enter();
  __shared__ __align__(16) float s_w[kMcWeightFloats];
  __shared__ float s_dw[kHeadMaxK * 8];
  __shared__ float s_db[8];
  __shared__ float s_bn[2 * kHeadMaxK];
  if (bn_sums) for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) s_bn[i] = 0.f;
  for (int i = threadIdx.x; i < K * 8; i += blockDim.x) { s_w[mc_row(i >> 3) + (i & 7)] = (i & 7) < C ? w[(i >> 3) * C + (i & 7)] : 0.f; s_dw[i] = 0.f; }
  if (threadIdx.x < 8) s_db[threadIdx.x] = 0.f;
  __syncthreads();
  const int64_t n = blockIdx.y;
  const int sub = threadIdx.x & 7, slot = threadIdx.x >> 3, lane = threadIdx.x & 31;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = i64min(hw, p_begin + pix_per_block);
  const float ca = sub < C ? coef[(n * C + sub) * 2] : 0.f, cb = sub < C ? coef[(n * C + sub) * 2 + 1] : 0.f;
  float dbs = 0.f;
  for (int k0 = sub * 8; k0 < K; k0 += 64) {
    float2 dwacc[8][4];                    // [channel][class pair]: packed FFMA2
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) dwacc[j][c] = make_float2(0.f, 0.f);
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
    // software-pipelined by one pixel: the loads of the next pixel are issued before the 128 FMAs of the current one
    float vn[8], prn = 0.f, ytn = 0.f;
    {
      const int64_t p = p_begin + slot;
      const int64_t m = n * hw + (p < p_end ? p : p_begin);
      load8(x + m * ldx + k0, vn);
      if (sub < C) { prn = __ldg(probs + m * C + sub); ytn = __ldg(y_true + m * C + sub); }
    }
    for (int64_t p0 = p_begin; p0 < p_end; p0 += 32) {
      const int64_t p = p0 + slot;
      const bool live = p < p_end;
      const int64_t m = n * hw + (live ? p : p_begin);
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = vn[j];
      // lane `sub` owns class `sub`: dz_c = p_c (g_c - sum_c' g_c' p_c')
      const float pr = prn, g = sub < C ? fmaf(ca, ytn, cb) : 0.f;
      if (p0 + 32 < p_end) {
        const int64_t pn = p0 + 32 + slot;
        const int64_t mn = n * hw + (pn < p_end ? pn : p_begin);
        load8(x + mn * ldx + k0, vn);
        if (sub < C) { prn = __ldg(probs + mn * C + sub); ytn = __ldg(y_true + mn * C + sub); }
      }
      float dot = g * pr;
      dot += __shfl_xor_sync(0xffffffffu, dot, 1); dot += __shfl_xor_sync(0xffffffffu, dot, 2); dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      const float dz_own = live ? pr * (g - dot) : 0.f;
      float2 dz2[4];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dz2[c] = make_float2(__shfl_sync(0xffffffffu, dz_own, (lane & ~7) + 2 * c), __shfl_sync(0xffffffffu, dz_own, (lane & ~7) + 2 * c + 1));
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w0 = *reinterpret_cast<const float4*>(&s_w[mc_row(k0 + j)]);      // 16-byte shared loads, 8 distinct rows per warp
        const float4 w1 = *reinterpret_cast<const float4*>(&s_w[mc_row(k0 + j) + 4]);
        float2 sacc = mul2(dz2[0], make_float2(w0.x, w0.y));
        sacc = fma2(dz2[1], make_float2(w0.z, w0.w), sacc);
        sacc = fma2(dz2[2], make_float2(w1.x, w1.y), sacc);
        sacc = fma2(dz2[3], make_float2(w1.z, w1.w), sacc);
        const float2 vv = make_float2(v[j], v[j]);
#pragma unroll
        for (int c = 0; c < 4; ++c) dwacc[j][c] = fma2(vv, dz2[c], dwacc[j][c]);
        o[j] = sacc.x + sacc.y;
      }
      if (bn_sums && live) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float gm = v[j] > 0.f ? round_to<T>(o[j]) : 0.f;
          o[j] = gm; s1[j] += gm; s2[j] = fmaf(gm, v[j], s2[j]);
        }
      }
      if (dx && live) store8(dx + m * lddx + k0, o);
      if (k0 == sub * 8) dbs += dz_own;
    }
    if (bn_sums) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a = s1[j], b = s2[j];
        a += __shfl_xor_sync(0xffffffffu, a, 8); a += __shfl_xor_sync(0xffffffffu, a, 16);
        b += __shfl_xor_sync(0xffffffffu, b, 8); b += __shfl_xor_sync(0xffffffffu, b, 16);
        if (lane < 8) { atomicAdd(&s_bn[k0 + j], a); atomicAdd(&s_bn[K + k0 + j], b); }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float sacc = (c & 1) ? dwacc[j][c >> 1].y : dwacc[j][c >> 1].x;
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 8);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 16);
        if (lane < 8) atomicAdd(&s_dw[(k0 + j) * 8 + c], sacc);
      }
  }
  dbs += __shfl_xor_sync(0xffffffffu, dbs, 8); dbs += __shfl_xor_sync(0xffffffffu, dbs, 16);
  if (lane < 8 && sub < C) atomicAdd(&s_db[sub], dbs);
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) atomicAdd(&dw[i], s_dw[(i / C) * 8 + (i % C)]);
  if (threadIdx.x < C) atomicAdd(&db[threadIdx.x], s_db[threadIdx.x]);
  if (bn_sums) for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) atomicAdd(&bn_sums[i], s_bn[i]);
}

static void head_grid(int64_t NB, int64_t hw, dim3* grid, int* pix_per_block, int waves = 16) {
  // enough blocks to fill the machine ~4x, each block a multiple of 32 pixels inside one image
  int64_t blocks_per_img = i64max(1, ceil_div((int64_t)sm_count() * waves, NB));
  int64_t ppb = ceil_div(ceil_div(hw, blocks_per_img), 32) * 32;
  if (ppb < 512) ppb = 512;
  blocks_per_img = ceil_div(hw, ppb);
  *grid = dim3((unsigned)blocks_per_img, (unsigned)NB);
  *pix_per_block = (int)ppb;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_head_fwd(const void* x, int64_t ldx, const float* w, const float* b, float* probs,
                             const float* y_true, double* sums, int64_t M, int64_t hw, int K, int C, int dtype,
                             const float* x_scale, const float* x_shift, void* stream) {
  UNET_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), UNET_EINVAL, "head_fwd: x_scale/x_shift must come together");
  UNET_REQUIRE(!x_scale || (dtype == UNET_BF16 && C == 1 && K == 64 && ldx == 64 && aligned16(x_scale) && aligned16(x_shift)), UNET_EUNSUPPORTED,
               "head_fwd: BN+ReLU on load exists for the streamed binary head only (bf16, 64 channels, contiguous x)");
  UNET_REQUIRE(x && w && probs && M > 0 && hw > 0 && K > 0 && C > 0 && ldx >= K, UNET_EINVAL, "head_fwd: bad argument");
  UNET_REQUIRE(M % hw == 0, UNET_EINVAL, "head_fwd: M must be a multiple of hw");
  UNET_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && aligned16(x), UNET_EALIGN, "head_fwd: needs K%%8==0, ld%%8==0, 16B pointer");
  UNET_REQUIRE(K <= kHeadMaxK && C <= 8, UNET_EUNSUPPORTED, "head_fwd: K <= %d and num_classes <= 8 (got %d, %d)", kHeadMaxK, K, C);
  UNET_REQUIRE(M / hw <= 65535, UNET_EUNSUPPORTED, "head_fwd: batch <= 65535");
  UNET_REQUIRE(!y_true || sums, UNET_EINVAL, "head_fwd: y_true given without sums");
  dim3 grid; int ppb;
  head_grid(M / hw, hw, &grid, &ppb);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_BF16 && C == 1 && K == 64 && ldx == 64) {     // the reference head: streamed kernel
    launch_pdl(head1_fwd_stream_kernel, grid, 256, 0, st, (const __nv_bfloat16*)x, w, b, probs, y_true, sums, hw, ppb, x_scale, x_shift);
    UNET_LAUNCH_CHECK("head_fwd(stream)");
    return UNET_OK;
  }
#define LAUNCH(T, MC) launch_pdl(head_fwd_kernel<T, MC>, grid, 256, 0, st, (const T*)x, ldx, w, b, probs, y_true, sums, hw, K, C, ppb)
#define LAUNCH_MC(T) launch_pdl(head_fwd_mc_kernel<T>, grid, 256, 0, st, (const T*)x, ldx, w, b, probs, y_true, sums, hw, K, C, ppb)
  if (dtype == UNET_F32)       { if (C == 1) LAUNCH(float, 1); else LAUNCH_MC(float); }
  else if (dtype == UNET_BF16) { if (C == 1) LAUNCH(__nv_bfloat16, 1); else LAUNCH_MC(__nv_bfloat16); }
#undef LAUNCH_MC
  else return set_error(UNET_EINVAL, "head_fwd: bad dtype %d", dtype);
#undef LAUNCH
  UNET_LAUNCH_CHECK("head_fwd");
  return UNET_OK;
}

extern "C" int unet_head_bwd(const void* x, int64_t ldx, const float* w, const float* probs, const float* y_true,
                             const float* coef, void* dx, int64_t lddx, float* dw, float* db,
                             int64_t M, int64_t hw, int K, int C, int dtype, float* bn_sums,
                             const float* x_scale, const float* x_shift, void* stream) {
  UNET_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), UNET_EINVAL, "head_bwd: x_scale/x_shift must come together");
  UNET_REQUIRE(!x_scale || (dtype == UNET_BF16 && C == 1 && K == 64 && ldx == 64 && dx && lddx == 64 && aligned16(x_scale) && aligned16(x_shift)),
               UNET_EUNSUPPORTED, "head_bwd: BN+ReLU on load exists for the streamed binary head only (bf16, 64 channels, contiguous x and dx)");
  UNET_REQUIRE(x && w && probs && y_true && coef && dw && db, UNET_EINVAL, "head_bwd: null pointer");
  UNET_REQUIRE(M > 0 && hw > 0 && K > 0 && C > 0 && ldx >= K && M % hw == 0, UNET_EINVAL, "head_bwd: bad dims");
  UNET_REQUIRE(K % 8 == 0 && ldx % 8 == 0 && aligned16(x) && (!dx || (lddx % 8 == 0 && aligned16(dx))), UNET_EALIGN,
               "head_bwd: needs K%%8==0, ld%%8==0, 16B pointers");
  UNET_REQUIRE(K <= kHeadMaxK && C <= 8, UNET_EUNSUPPORTED, "head_bwd: K <= %d and num_classes <= 8", kHeadMaxK);
  UNET_REQUIRE(M / hw <= 65535, UNET_EUNSUPPORTED, "head_bwd: batch <= 65535");
  UNET_REQUIRE(!bn_sums || dx, UNET_EINVAL, "head_bwd: bn_sums needs dx");
  dim3 grid; int ppb;
  head_grid(M / hw, hw, &grid, &ppb, 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_BF16 && C == 1 && K == 64 && ldx == 64 && dx && lddx == 64) {     // the reference head: streamed kernel
    head_grid(M / hw, hw, &grid, &ppb, 16);
    if (bn_sums) launch_pdl(head1_bwd_stream_kernel<true>, grid, 256, 0, st, (const __nv_bfloat16*)x, w, probs, y_true, coef, (__nv_bfloat16*)dx, dw, db, hw, ppb, bn_sums, x_scale, x_shift);
    else launch_pdl(head1_bwd_stream_kernel<false>, grid, 256, 0, st, (const __nv_bfloat16*)x, w, probs, y_true, coef, (__nv_bfloat16*)dx, dw, db, hw, ppb, bn_sums, x_scale, x_shift);
    UNET_LAUNCH_CHECK("head_bwd(stream)");
    return UNET_OK;
  }
#define LAUNCH(T, MC) do { if (bn_sums) launch_pdl(head_bwd_kernel<T, MC, true>, grid, 256, 0, st, (const T*)x, ldx, w, probs, y_true, coef, (T*)dx, lddx, dw, db, hw, K, C, ppb, bn_sums); \
                           else launch_pdl(head_bwd_kernel<T, MC, false>, grid, 256, 0, st, (const T*)x, ldx, w, probs, y_true, coef, (T*)dx, lddx, dw, db, hw, K, C, ppb, bn_sums); } while (0)
#define LAUNCH_MC(T) launch_pdl(head_bwd_mc_kernel<T>, grid, 256, 0, st, (const T*)x, ldx, w, probs, y_true, coef, (T*)dx, lddx, dw, db, hw, K, C, ppb, bn_sums)
  if (dtype == UNET_F32)       { if (C == 1) LAUNCH(float, 1); else LAUNCH_MC(float); }
  else if (dtype == UNET_BF16) { if (C == 1) LAUNCH(__nv_bfloat16, 1); else LAUNCH_MC(__nv_bfloat16); }
#undef LAUNCH_MC
  else return set_error(UNET_EINVAL, "head_bwd: bad dtype %d", dtype);
#undef LAUNCH
  UNET_LAUNCH_CHECK("head_bwd");
  return UNET_OK;
}
