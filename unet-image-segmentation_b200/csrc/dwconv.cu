// Depthwise 3x3 convolution (the depthwise half of SeparableConv2D, reference model/u_net.py:14-20).
// NHWC, zero 'same' padding, stride 1, depth multiplier 1, cross-correlation (no kernel flip).
//
// HBM-bound (18 flop per output element, ~4.5 flop/B in bf16): the kernels are organised so that every
// input element is requested from DRAM once.  Each thread owns 8 (forward) or 4 (weight gradient) consecutive
// channels of one image column and slides down a segment of rows keeping the running partial sums in
// registers; the three column taps of a row come from the two neighbouring threads' lines in L1.
#include <type_traits>
#include "common.cuh"
#include "ptx.cuh"

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
  pdl_enter();
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
        dropout_apply(o, base, drop_seed(dp), dp.keep, dp.inv_keep);
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


// ------------------------------------------------------------------------------------------------ forward, TMA strips
// One CTA = one strip: TW = PXT*CW output columns x 128 B of channels x a segment of rows of one image.  Thread 0 streams
// the strip top to bottom through a ring of S shared-memory stages with 4-D TMA boxes (channels, columns + 2 halo, RH rows,
// image), S-1 stages ahead; rows -1 / H and columns -1 / W come back as zeros from the TMA out-of-bounds fill, which IS
// the 'same' padding.  The PXT*16 threads slide down the strip: a thread owns 8 B of channels of CW adjacent columns, keeps
// the two open partial sums of each in registers, and per input row does CW+2 conflict-free 8-byte shared loads and
// 9*CW packed FFMA2 per channel pair; outputs leave as 8-byte global stores (128 B contiguous per column).
// HBM sees every input byte once plus 2/TW column halo (served by L2) and 2/segment rows.
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;

// CW = output columns per thread, PXT = column threads per CTA (x 16 channel groups of 8 bytes).  CW = 2 loads 4 columns
// for 2 outputs instead of 3 for 1: a third fewer shared loads / unpacks / on-load transforms per output element.
template <typename T, int CW_ = 1, int PXT_ = 32> struct StripCfg {
  static constexpr int NV = 8 / (int)sizeof(T);          // channels per thread (8 bytes)
  static constexpr int CB = 128 / (int)sizeof(T);        // channels per CTA (128 bytes)
  static constexpr int CW = CW_, PXT = PXT_;
  static constexpr int TW = PXT * CW, RH = 4, S = (TW > 32) ? 6 : (PXT == 8 ? 4 : (PXT < 32 ? 5 : 8));
  static constexpr int kThreads = PXT * 16;
  static constexpr int kMinBlocks = PXT == 8 ? 3 : (PXT < 32 ? 2 : 1);
  static constexpr int kStageBytes = RH * (TW + 2) * 128;
  static constexpr int kSmemBytes = S * kStageBytes + 2 * S * 8 + CB * 4 + 128;
};

// 8 bytes of channels -> fp32 pairs (the operands of the packed FFMA2 path)
template <typename T> __device__ __forceinline__ void unpack8(const uint2& r, float2 (&v)[4 / sizeof(T)]);
template <> __device__ __forceinline__ void unpack8<__nv_bfloat16>(const uint2& r, float2 (&v)[2]) {
  v[0] = make_float2(bf16lo_to_f32(r.x), __uint_as_float(r.x & 0xffff0000u));
  v[1] = make_float2(bf16lo_to_f32(r.y), __uint_as_float(r.y & 0xffff0000u));
}
template <> __device__ __forceinline__ void unpack8<float>(const uint2& r, float2 (&v)[1]) {
  v[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
}
__device__ __forceinline__ uint2 pack8(const float (&v)[4], __nv_bfloat16*) {
  return make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}
__device__ __forceinline__ uint2 pack8(const float (&v)[2], float*) {
  return make_uint2(__float_as_uint(v[0]), __float_as_uint(v[1]));
}

// AFFINE: the input tensor is the producer's PRE-BatchNormalization output z; x' = max(z*in_scale + in_shift, 0) is formed in
// registers on every load (CW+2 columns x the thread's channels per row), and the 'same' padding is re-imposed in x' space:
// rows outside the image are zeroed by a CTA-uniform branch, the two halo columns are multiplied by a per-thread 0/1 factor
// (0 only for the threads at the image's left / right border).
template <typename T, int CW, int PXT, bool DROP, bool SUMS, bool AFFINE>
__global__ void __launch_bounds__((StripCfg<T, CW, PXT>::kThreads), (StripCfg<T, CW, PXT>::kMinBlocks))
dwconv3x3_strip_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w9c, T* __restrict__ y, int64_t ldy,
                       int H, int W, int C, int seg_rows, int nseg, int ntw, int ncb, int flip, DropArgs dp,
                       float* __restrict__ colsum, const float* __restrict__ in_scale, const float* __restrict__ in_shift) {
  pdl_launch_dependents();
  using Cfg = StripCfg<T, CW, PXT>;
  constexpr int NV = Cfg::NV, TW = Cfg::TW, RH = Cfg::RH, S = Cfg::S, NQ = CW + 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + S;
  float* s_sum = reinterpret_cast<float*>(empty_bar + S);      // [CB] column sums of the stored outputs (SUMS)

  int item = blockIdx.x;
  const int cb = item % ncb; item /= ncb;
  const int tw = item % ntw; item /= ntw;
  const int hs = item % nseg;
  const int n = item / nseg;
  const int c0 = cb * Cfg::CB, w0 = tw * TW;
  const int h0 = hs * seg_rows, h1 = min(H, h0 + seg_rows);
  const int nst = (h1 - h0 + 2 + RH - 1) / RH;          // input rows h0-1 .. h1
  const int lane = threadIdx.x & 31;

  if (SUMS) for (int i = threadIdx.x; i < Cfg::CB; i += blockDim.x) s_sum[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], Cfg::kThreads / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();            // above: shared memory and kernel parameters only

  // thread 0 doubles as the TMA producer: it keeps S-1 stages in flight ahead of the stage being consumed
  auto issue = [&](int k) {
    const int s = k % S;
    if (k >= S) mbar_wait(&empty_bar[s], ((k / S) - 1) & 1);
    mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
    tma_load_4d(smem + s * Cfg::kStageBytes, &tmX, &full_bar[s], c0, w0 - 1, h0 - 1 + k * RH, n, kEvictNormal);
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int k = 0; k < S - 1 && k < nst; ++k) issue(k);
  }

  const int px = threadIdx.x >> 4, cg = threadIdx.x & 15;
  const int c = c0 + cg * NV;
  const int col0 = w0 + px * CW;                   // first output column of this thread (W % CW == 0: all CW are live or none)
  const bool live = (col0 < W) && (c < C);
  constexpr int NP = NV / 2;                       // fp32 pairs per thread: every FMA below is a packed FFMA2
  float2 k9[9][NP];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int src = flip ? 8 - i : i;
#pragma unroll
    for (int j = 0; j < NP; ++j) k9[i][j] = make_float2(0.f, 0.f);
    if (c < C) {
      if (NV == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + (int64_t)src * C + c));
        k9[i][0] = make_float2(a.x, a.y); k9[i][1 % NP] = make_float2(a.z, a.w);
      } else {
        const float2 a = __ldg(reinterpret_cast<const float2*>(w9c + (int64_t)src * C + c));
        k9[i][0] = a;
      }
    }
  }
  uint32_t seed = 0u;
  if (DROP) seed = drop_seed(dp);
  float2 asc[NP], ash[NP];                          // AFFINE: scale / shift of this thread's channels
  float2 lm2 = make_float2(1.f, 1.f), rm2 = make_float2(1.f, 1.f);
  if (AFFINE) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { asc[j] = make_float2(0.f, 0.f); ash[j] = make_float2(0.f, 0.f); }
    if (c < C) {
      if (NV == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in_scale + c)), b = __ldg(reinterpret_cast<const float4*>(in_shift + c));
        asc[0] = make_float2(a.x, a.y); asc[1 % NP] = make_float2(a.z, a.w); ash[0] = make_float2(b.x, b.y); ash[1 % NP] = make_float2(b.z, b.w);
      } else {
        asc[0] = __ldg(reinterpret_cast<const float2*>(in_scale + c)); ash[0] = __ldg(reinterpret_cast<const float2*>(in_shift + c));
      }
    }
    if (col0 - 1 < 0) lm2 = make_float2(0.f, 0.f);
    if (col0 + CW >= W) rm2 = make_float2(0.f, 0.f);
  }
  float2 prev[CW][NP], cur[CW][NP], csum[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    csum[j] = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < CW; ++q) { prev[q][j] = make_float2(0.f, 0.f); cur[q][j] = make_float2(0.f, 0.f); }
  }
  T* yptr = y + (((int64_t)n * H + (h0 - 2)) * W + col0) * ldy + c;   // advanced one row per input row
  const int64_t yrow = (int64_t)W * ldy;
  const uint32_t tile_off = (uint32_t)(px * CW) * 128u + (uint32_t)cg * 8u;
  const uint32_t smem_base = smem_u32(smem);
  int r_in = h0 - 1;

  for (int k = 0; k < nst; ++k) {
    const int s = k % S;
    if (threadIdx.x == 0 && k + S - 1 < nst) issue(k + S - 1);
    mbar_wait(&full_bar[s], (k / S) & 1);
    const uint32_t st = smem_base + s * Cfg::kStageBytes + tile_off;
    uint2 raw[RH][NQ];
#pragma unroll
    for (int rr = 0; rr < RH; ++rr)          // all shared loads of the stage first: RH * (CW+2) independent requests in flight
#pragma unroll
      for (int q = 0; q < NQ; ++q) raw[rr][q] = lds64(st + rr * (TW + 2) * 128 + q * 128);
#pragma unroll
    for (int rr = 0; rr < RH; ++rr, ++r_in, yptr += yrow) {
      float2 v[NQ][NP];
#pragma unroll
      for (int q = 0; q < NQ; ++q) unpack8<T>(raw[rr][q], v[q]);
      if (AFFINE) {
        if ((unsigned)r_in < (unsigned)H) {       // uniform over the CTA
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              const float2 t = fma2(v[q][j], asc[j], ash[j]);
              v[q][j] = make_float2(fmaxf(t.x, 0.f), fmaxf(t.y, 0.f));
            }
#pragma unroll
          for (int j = 0; j < NP; ++j) { v[0][j] = mul2(v[0][j], lm2); v[NQ - 1][j] = mul2(v[NQ - 1][j], rm2); }
        } else {
#pragma unroll
          for (int q = 0; q < NQ; ++q)
#pragma unroll
            for (int j = 0; j < NP; ++j) v[q][j] = make_float2(0.f, 0.f);
        }
      }
      // the 9 FMAs per (column, channel pair) of this input row, ordered by the input column they read (the second operand of
      // each): order-pinned FFMA2 runs that take it from the operand-reuse cache (see fma2v); same summation order as
      // out = k8*v[q+2] + (k7*v[q+1] + (k6*v[q] + prev))
      float2 rv[CW][NP], pn[CW][NP], cn[CW][NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
#pragma unroll
        for (int q = 0; q < CW; ++q) { rv[q][j] = prev[q][j]; pn[q][j] = cur[q][j]; cn[q][j] = make_float2(0.f, 0.f); }
#pragma unroll
        for (int cc = 0; cc < NQ; ++cc)
#pragma unroll
          for (int q = 0; q < CW; ++q) {
            const int sft = cc - q;                 // tap column
            if (sft < 0 || sft > 2) continue;
            rv[q][j] = fma2v(k9[6 + sft][j], v[cc][j], rv[q][j]);
            pn[q][j] = fma2v(k9[3 + sft][j], v[cc][j], pn[q][j]);
            cn[q][j] = sft == 0 ? mul2v(k9[0][j], v[cc][j]) : fma2v(k9[sft][j], v[cc][j], cn[q][j]);
          }
      }
      if (r_in > h0 && r_in <= h1 && live) {   // output row r_in-1 is complete once kernel row 2 has seen input row r_in
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          float o[NV];
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float2 r = rv[q][j];
            o[2 * j] = r.x; o[2 * j + 1] = r.y;
          }
          if (DROP) {
            const uint64_t base = (uint64_t)(((int64_t)n * H + (r_in - 1)) * W + (col0 + q)) * dp.ctot + dp.c0 + c;
            dropout_apply(o, base, seed, dp.keep, dp.inv_keep);
          }
          const uint2 packed = pack8(o, (T*)nullptr);
          *reinterpret_cast<uint2*>(yptr + q * ldy) = packed;
          if (SUMS) {                               // column sums of the values as stored (what the pointwise GEMM reads)
            float2 st2[NP];
            unpack8<T>(packed, st2);
#pragma unroll
            for (int j = 0; j < NP; ++j) csum[j] = add2(csum[j], st2[j]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < CW; ++q)
#pragma unroll
        for (int j = 0; j < NP; ++j) { prev[q][j] = pn[q][j]; cur[q][j] = cn[q][j]; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  if (SUMS) {                                     // lanes l and l^16 hold the same channels of neighbouring columns
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float mine = (j & 1) ? csum[j / 2].y : csum[j / 2].x;
      const float v = mine + __shfl_xor_sync(0xffffffffu, mine, 16);
      if (lane < 16) red_shared_add(smem_u32(&s_sum[cg * NV + j]), v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Cfg::CB; i += blockDim.x)
      if (c0 + i < C) atomicAdd(&colsum[c0 + i], s_sum[i]);
  }
}

// NHWC view (ptr, ld) as a 4-D tensor map {C, W, H, N}; box {128 B of channels, box_w, box_h, 1}; zero OOB fill
template <typename T>
static int make_nhwc_tmap(CUtensorMap* map, const void* base, int64_t ld, int N, int H, int W, int C, int box_w, int box_h,
                          const char* who) {
  PFN_encodeTiled fn = get_encode_fn();
  UNET_REQUIRE(fn, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  constexpr int es = (int)sizeof(T);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * es, (cuuint64_t)W * ld * es, (cuuint64_t)H * W * ld * es};
  cuuint32_t box[4] = {(cuuint32_t)(128 / es), (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = fn(map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                        const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return UNET_OK;
}

static int pick_seg_rows(int N, int H, int ntw, int ncb, int min_rows, int64_t want_items) {
  int seg = H;
  while (seg > min_rows && (int64_t)N * ceil_div(H, seg) * ntw * ncb < want_items) seg = (seg + 1) / 2;
  return seg;
}

template <typename T, int CW, int PXT>
static int dw_fwd_strip_launch_cfg(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy, int N, int H, int W, int C,
                                   int flip, DropArgs dp, float* colsum, const float* in_scale, const float* in_shift, cudaStream_t st) {
  using Cfg = StripCfg<T, CW, PXT>;
  CUtensorMap tm;
  if (int e = make_nhwc_tmap<T>(&tm, x, ldx, N, H, W, C, Cfg::TW + 2, Cfg::RH, "dwconv3x3_fwd")) return e;
  static SmemAttrOnce once_plain, once_drop, once_sums, once_aff, once_aff_sums;
  cudaError_t ea = ensure_dynamic_smem(once_plain, dwconv3x3_strip_kernel<T, CW, PXT, false, false, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(once_drop, dwconv3x3_strip_kernel<T, CW, PXT, true, false, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(once_sums, dwconv3x3_strip_kernel<T, CW, PXT, false, true, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(once_aff, dwconv3x3_strip_kernel<T, CW, PXT, false, false, true>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(once_aff_sums, dwconv3x3_strip_kernel<T, CW, PXT, false, true, true>, Cfg::kSmemBytes);
  if (ea != cudaSuccess) return set_cuda_error(ea, "dwconv3x3_fwd: cudaFuncSetAttribute");
  const int ntw = (int)ceil_div(W, Cfg::TW), ncb = (int)ceil_div(C, Cfg::CB);
  const int seg = pick_seg_rows(N, H, ntw, ncb, 32, (int64_t)sm_count() * 6 * Cfg::kMinBlocks);
  const int nseg = (int)ceil_div(H, seg);
  const int64_t items = (int64_t)N * nseg * ntw * ncb;
  UNET_REQUIRE(items < ((int64_t)1 << 31), UNET_EUNSUPPORTED, "dwconv3x3_fwd: too many strips");
#define UNET_DWF(D_, S_, A_) launch_pdl(dwconv3x3_strip_kernel<T, CW, PXT, D_, S_, A_>, (unsigned)items, Cfg::kThreads, Cfg::kSmemBytes, st, \
      tm, w9c, (T*)y, ldy, H, W, C, seg, nseg, ntw, ncb, flip, dp, colsum, in_scale, in_shift)
  if (dp.on) UNET_DWF(true, false, false);
  else if (in_scale) { if (colsum) UNET_DWF(false, true, true); else UNET_DWF(false, false, true); }
  else if (colsum) UNET_DWF(false, true, false);
  else UNET_DWF(false, false, false);
#undef UNET_DWF
  UNET_LAUNCH_CHECK("dwconv3x3_fwd(strip)");
  return UNET_OK;
}

template <typename T>
static int dw_fwd_strip_launch(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy, int N, int H, int W, int C,
                               int flip, DropArgs dp, float* colsum, const float* in_scale, const float* in_shift, cudaStream_t st) {
  UNET_REQUIRE(!(colsum && dp.on), UNET_EUNSUPPORTED, "dwconv3x3_fwd: colsum and dropout cannot be combined");
  UNET_REQUIRE(!(in_scale && dp.on), UNET_EUNSUPPORTED, "dwconv3x3_fwd: input affine and dropout cannot be combined");
  // 4 output columns per thread (6 loaded: half the shared loads / unpacks / on-load transforms of the 1-column form) when
  // the row splits into whole groups of 4; 2 per thread for other even widths; 1 per thread otherwise and for narrow rows
  if (W % 4 == 0 && W >= 32)
    return dw_fwd_strip_launch_cfg<T, 4, 8>(x, ldx, w9c, y, ldy, N, H, W, C, flip, dp, colsum, in_scale, in_shift, st);
  if (W % 2 == 0 && W >= 32)
    return dw_fwd_strip_launch_cfg<T, 2, 16>(x, ldx, w9c, y, ldy, N, H, W, C, flip, dp, colsum, in_scale, in_shift, st);
  return dw_fwd_strip_launch_cfg<T, 1, 32>(x, ldx, w9c, y, ldy, N, H, W, C, flip, dp, colsum, in_scale, in_shift, st);
}

// any channel count (used for the 3-channel input image)
template <typename T, bool AFFINE>
__global__ void dwconv3x3_scalar_kernel(const T* __restrict__ x, int64_t ldx, const float* __restrict__ w9c,
                                        T* __restrict__ y, int64_t ldy, int N, int H, int W, int C, int flip,
                                        const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                                        DropArgs dp) {
  pdl_enter();
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
                         int flip, const float* in_scale, const float* in_shift, DropArgs dp, float* colsum, cudaStream_t st) {
  const bool vec = (C % 8 == 0) && (ldx % 8 == 0) && (ldy % 8 == 0) && aligned16(x) && aligned16(y) &&
                   aligned16(w9c) && (!in_scale || (aligned16(in_scale) && aligned16(in_shift)));
  if (vec && !(in_scale && dp.on))   // TMA strips; a fused input affine re-imposes the zero padding after the transform
    return dw_fwd_strip_launch<T>(x, ldx, w9c, y, ldy, N, H, W, C, flip, dp, colsum, in_scale, in_shift, st);
  UNET_REQUIRE(!colsum, UNET_EUNSUPPORTED, "dwconv3x3_fwd: colsum needs the strip path (C%%8==0, 16B-aligned views, no input affine)");
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
      launch_pdl(dwconv3x3_vec8_kernel<T, true>, grid, 256, 0, st, (const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, R, nseg, flip,
                                                          in_scale, in_shift, dp);
    else
      launch_pdl(dwconv3x3_vec8_kernel<T, false>, grid, 256, 0, st, (const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, R, nseg, flip,
                                                           nullptr, nullptr, dp);
  } else {
    const int64_t total = (int64_t)N * H * W * C;
    const unsigned grid = (unsigned)i64min(ceil_div(total, 256), (int64_t)sm_count() * 32);
    if (in_scale)
      launch_pdl(dwconv3x3_scalar_kernel<T, true>, grid, 256, 0, st, (const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, flip,
                                                            in_scale, in_shift, dp);
    else
      launch_pdl(dwconv3x3_scalar_kernel<T, false>, grid, 256, 0, st, (const T*)x, ldx, w9c, (T*)y, ldy, N, H, W, C, flip,
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
  pdl_enter();
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
  pdl_enter();
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


// ------------------------------------------------------------------------------------------------ weight gradient, TMA strips
// Same strip decomposition as the forward kernel, two TMA streams per stage: x with its column halo (rows h0-1 ..)
// and dy one row ahead (rows h0 ..).  512 consumer threads: a thread owns 8 B of channels of one column, keeps the
// 9 tap accumulators of those channels in registers for the whole strip and the three live dy rows in a register
// window; per input row it does 4 conflict-free 8-byte shared loads and 9 FMAs per channel.  The strip's partial
// sums are folded with one shuffle, shared-memory atomics across the 16 warps, then 9 x 128 B of global atomics.
template <typename T> struct WgCfg {
  static constexpr int NV = 8 / (int)sizeof(T);          // channels per thread
  static constexpr int CB = 128 / (int)sizeof(T);        // channels per CTA
  static constexpr int TW = 32, RH = 4, S = 5;
  static constexpr int kXBytes = RH * (TW + 2) * 128;
  static constexpr int kDBytes = RH * TW * 128;
  static constexpr int kStageBytes = kXBytes + kDBytes;
  static constexpr int kSmemBytes = S * kStageBytes + 9 * CB * 4 + 2 * S * 8 + 128;
};


template <typename T>
__global__ void __launch_bounds__(512, 1)
dwconv3x3_wgrad_strip_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmD,
                             float* __restrict__ dw9c, int H, int W, int C, int seg_rows, int nseg, int ntw, int ncb) {
  pdl_launch_dependents();
  using Cfg = WgCfg<T>;
  constexpr int NV = Cfg::NV, TW = Cfg::TW, RH = Cfg::RH, S = Cfg::S;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* s_acc = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);       // [9][CB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_acc + 9 * Cfg::CB);
  uint64_t* empty_bar = full_bar + S;

  int item = blockIdx.x;
  const int cb = item % ncb; item /= ncb;
  const int tw = item % ntw; item /= ntw;
  const int hs = item % nseg;
  const int n = item / nseg;
  const int c0 = cb * Cfg::CB, w0 = tw * TW;
  const int h0 = hs * seg_rows, h1 = min(H, h0 + seg_rows);
  const int nst = (h1 - h0 + 2 + RH - 1) / RH;          // x rows h0-1 .. h1
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 9 * Cfg::CB; i += blockDim.x) s_acc[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 16); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();            // above: shared memory and kernel parameters only

  auto issue = [&](int k) {
    const int s = k % S;
    if (k >= S) mbar_wait(&empty_bar[s], ((k / S) - 1) & 1);
    mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
    uint8_t* st = smem + s * Cfg::kStageBytes;
    tma_load_4d(st, &tmX, &full_bar[s], c0, w0 - 1, h0 - 1 + k * RH, n, kEvictNormal);
    tma_load_4d(st + Cfg::kXBytes, &tmD, &full_bar[s], c0, w0, h0 + k * RH, n, kEvictNormal);
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmD);
    for (int k = 0; k < S - 1 && k < nst; ++k) issue(k);
  }

  const int px = threadIdx.x >> 4, cg = threadIdx.x & 15;
  constexpr int NP = NV / 2;                       // fp32 pairs per thread (packed FFMA2)
  float2 acc[9][NP], dm[NP], d0[NP];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[i][j] = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < NP; ++j) { dm[j] = make_float2(0.f, 0.f); d0[j] = make_float2(0.f, 0.f); }
  const uint32_t xoff = (uint32_t)px * 128u + (uint32_t)cg * 8u;
  const uint32_t smem_base = smem_u32(smem);

  for (int k = 0; k < nst; ++k) {
    const int s = k % S;
    if (threadIdx.x == 0 && k + S - 1 < nst) issue(k + S - 1);
    mbar_wait(&full_bar[s], (k / S) & 1);
    const uint32_t sx = smem_base + s * Cfg::kStageBytes + xoff;
    const uint32_t sd = sx + Cfg::kXBytes;
    uint2 ra[RH], rb[RH], rc[RH], rd[RH];
#pragma unroll
    for (int rr = 0; rr < RH; ++rr) {        // all shared loads of the stage first (16 independent requests in flight)
      ra[rr] = lds64(sx + rr * (TW + 2) * 128);
      rb[rr] = lds64(sx + rr * (TW + 2) * 128 + 128);
      rc[rr] = lds64(sx + rr * (TW + 2) * 128 + 256);
      rd[rr] = lds64(sd + rr * TW * 128);
    }
#pragma unroll
    for (int rr = 0; rr < RH; ++rr) {
      const int q = h0 - 1 + k * RH + rr;           // x row; dy row q+1 sits at the same stage row
      float2 a[NP], b[NP], c[NP], dp[NP];
      unpack8<T>(ra[rr], a); unpack8<T>(rb[rr], b); unpack8<T>(rc[rr], c); unpack8<T>(rd[rr], dp);
      const bool dp_live = q + 1 < h1;              // dy rows of other segments (or past the image) do not belong to this strip
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float2 dpj = dp_live ? dp[j] : make_float2(0.f, 0.f);
        const float2 aj = a[j], bj = b[j], cj = c[j];   // rows past h1 meet only zeroed dy rows
        // kernel row r multiplies input row q into output row q - r + 1
        acc[0][j] = fma2(aj, dpj, acc[0][j]);   acc[1][j] = fma2(bj, dpj, acc[1][j]);   acc[2][j] = fma2(cj, dpj, acc[2][j]);
        acc[3][j] = fma2(aj, d0[j], acc[3][j]); acc[4][j] = fma2(bj, d0[j], acc[4][j]); acc[5][j] = fma2(cj, d0[j], acc[5][j]);
        acc[6][j] = fma2(aj, dm[j], acc[6][j]); acc[7][j] = fma2(bj, dm[j], acc[7][j]); acc[8][j] = fma2(cj, dm[j], acc[8][j]);
        dm[j] = d0[j]; d0[j] = dpj;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // lanes l and l^16 hold the same channels of neighbouring columns
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float mine = (j & 1) ? acc[i][j / 2].y : acc[i][j / 2].x;
      const float v = mine + __shfl_xor_sync(0xffffffffu, mine, 16);
      if (lane < 16) red_shared_add(smem_u32(&s_acc[i * Cfg::CB + cg * NV + j]), v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * Cfg::CB; i += blockDim.x) {
    const int tap = i / Cfg::CB, ch = c0 + i % Cfg::CB;
    if (ch < C) atomicAdd(&dw9c[(int64_t)tap * C + ch], s_acc[i]);
  }
}

template <typename T>
static int dw_wgrad_strip_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c, int N, int H, int W, int C,
                                 cudaStream_t st) {
  using Cfg = WgCfg<T>;
  CUtensorMap tmX, tmD;
  if (int e = make_nhwc_tmap<T>(&tmX, x, ldx, N, H, W, C, Cfg::TW + 2, Cfg::RH, "dwconv3x3_bwd_weight(x)")) return e;
  if (int e = make_nhwc_tmap<T>(&tmD, dy, lddy, N, H, W, C, Cfg::TW, Cfg::RH, "dwconv3x3_bwd_weight(dy)")) return e;
  static SmemAttrOnce once;
  if (cudaError_t e = ensure_dynamic_smem(once, dwconv3x3_wgrad_strip_kernel<T>, Cfg::kSmemBytes))
    return set_cuda_error(e, "dwconv3x3_bwd_weight: cudaFuncSetAttribute");
  const int ntw = (int)ceil_div(W, Cfg::TW), ncb = (int)ceil_div(C, Cfg::CB);
  const int seg = pick_seg_rows(N, H, ntw, ncb, 32, (int64_t)sm_count() * 4);
  const int nseg = (int)ceil_div(H, seg);
  const int64_t items = (int64_t)N * nseg * ntw * ncb;
  UNET_REQUIRE(items < ((int64_t)1 << 31), UNET_EUNSUPPORTED, "dwconv3x3_bwd_weight: too many strips");
  launch_pdl(dwconv3x3_wgrad_strip_kernel<T>, (unsigned)items, 512, Cfg::kSmemBytes, st, tmX, tmD, dw9c, H, W, C, seg, nseg, ntw, ncb);
  UNET_LAUNCH_CHECK("dwconv3x3_bwd_weight(strip)");
  return UNET_OK;
}

// ------------------------------------------------------------------------------------------------ fused backward, TMA strips
// Both gradients of the depthwise convolution from ONE pass over dy (SeparableConv2D backward, u_net.py:14-20):
//   dx = dy (*) rot180(w)                      (the forward strip kernel with the flipped taps)
//   dw[r][s] += sum x[i][j] * dy[i-r+1][j-s+1]  (x centre column only; the column halo is on dy, which dx needs anyway)
// so dy is read once instead of twice and x needs no halo.  Two TMA streams per stage: dy with its column halo (rows
// h0-1 ..) and x one row ahead (rows h0 ..).  RELU_MASK: x is the post-ReLU output y of the producing conv_block, so
// (x > 0) is that block's ReLU mask: the kernel stores dx * (x > 0) — the gradient w.r.t. the BatchNormalization output —
// and accumulates the two reductions BatchNormalization backward needs, sum(g) and sum(g * y), from the stored values.
// CW = output columns per thread, PXT = column threads per CTA.  CW = 2: the 9 taps and the 9 tap accumulators are shared by
// both columns (the weight gradient sums over pixels anyway), 4 dy columns are loaded for 2 outputs instead of 3 for 1, and
// every thread carries two independent sliding-sum chains — more instruction-level parallelism at the same register budget
// per SM (the kernel is latency-bound at 12 warps/SM, not pipe-bound).
template <typename T, int CW_ = 1, int PXT_ = 24> struct BwCfg {
  static constexpr int NV = 8 / (int)sizeof(T);         // channels per thread (8 bytes)
  static constexpr int CB = 128 / (int)sizeof(T);       // channels per CTA (128 bytes)
  static constexpr int CW = CW_, PXT = PXT_;
  // <= 384 threads per SM: up to 170 registers per thread hold the 9 taps, the 9 tap accumulators, the sliding sums and a
  // 4-row window of shared loads without spilling (512 threads would cap at 128)
  static constexpr int TW = PXT * CW, RH = 4;
  static constexpr int kThreads = PXT * 16;
  // (4 CTAs / 128 registers per thread was measured in round 2: the ~40 spilled registers double the kernel's time)
  static constexpr int kMinBlocks = 384 / kThreads;
  static constexpr int kDBytes = RH * (TW + 2) * 128;
  static constexpr int kXBytes = RH * TW * 128;
  static constexpr int kStageBytes = kDBytes + kXBytes;
  static constexpr int kFixedBytes = 11 * CB * 4 + 2 * 8 * 8 + 128 + 1024;
  static constexpr int S_fit = (227 * 1024 / kMinBlocks - kFixedBytes) / kStageBytes;
  static constexpr int S = S_fit > 6 ? 6 : S_fit;
  static constexpr int kSmemBytes = S * kStageBytes + 11 * CB * 4 + 2 * S * 8 + 128;
};

// Conv2DTranspose(k=2,s=2) consumer of the first `c_end` channels of dx (the upsampled half of a skip-concat gradient,
// u_net.py:88-96): those channel blocks are stored un-pixel-shuffled, row (n, i/2, j/2), column block (i%2, j%2), as the
// [N*H/2*W/2, 4*c_end] operand of the transposed convolution's weight / data gradient GEMMs (what unet_convt_bwd_gather
// would produce from dx), and their per-channel sums (the bias gradient) are accumulated into `colsum`.
struct UpArgs { void* out; int c_end; float* colsum; };

template <typename T, int CW, int PXT, bool DROP, bool RELU_MASK, bool AFFINE, bool UP = false>
__global__ void __launch_bounds__((BwCfg<T, CW, PXT>::kThreads), (BwCfg<T, CW, PXT>::kMinBlocks))
dwconv3x3_bwd_strip_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmX,
                           const float* __restrict__ w9c, T* __restrict__ dx, int64_t lddx, float* __restrict__ dw9c,
                           float* __restrict__ bn_sums, int H, int W, int C, int seg_rows, int nseg, int ntw, int ncb, DropArgs dp,
                           int drop_c_from, const float* __restrict__ x_scale, const float* __restrict__ x_shift, UpArgs up) {
  pdl_launch_dependents();
  using Cfg = BwCfg<T, CW, PXT>;
  constexpr int NV = Cfg::NV, NP = NV / 2, TW = Cfg::TW, RH = Cfg::RH, S = Cfg::S, NQ = CW + 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* s_acc = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);       // [9 taps + 2 sums][CB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_acc + 11 * Cfg::CB);
  uint64_t* empty_bar = full_bar + S;

  int item = blockIdx.x;
  const int cb = item % ncb; item /= ncb;
  const int tw = item % ntw; item /= ntw;
  const int hs = item % nseg;
  const int n = item / nseg;
  const int c0 = cb * Cfg::CB, w0 = tw * TW;
  const int h0 = hs * seg_rows, h1 = min(H, h0 + seg_rows);
  const int nst = (h1 - h0 + 2 + RH - 1) / RH;          // dy rows h0-1 .. h1
  const int lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < 11 * Cfg::CB; i += blockDim.x) s_acc[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], Cfg::kThreads / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();            // above: shared memory and kernel parameters only

  auto issue = [&](int k) {
    const int s = k % S;
    if (k >= S) mbar_wait(&empty_bar[s], ((k / S) - 1) & 1);
    mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
    uint8_t* st = smem + s * Cfg::kStageBytes;
    tma_load_4d(st, &tmD, &full_bar[s], c0, w0 - 1, h0 - 1 + k * RH, n, kEvictNormal);
    tma_load_4d(st + Cfg::kDBytes, &tmX, &full_bar[s], c0, w0, h0 + k * RH, n, kEvictNormal);
  };
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmX);
    for (int k = 0; k < S - 1 && k < nst; ++k) issue(k);
  }

  const int px = threadIdx.x >> 4, cg = threadIdx.x & 15;
  const int c = c0 + cg * NV;
  const int col0 = w0 + px * CW;                   // first column of this thread (W % CW == 0: all CW are live or none)
  const bool live = (col0 < W) && (c < C);
  float2 kf[9][NP];                                // flipped taps: kf[i] = w[8 - i]
#pragma unroll
  for (int i = 0; i < 9; ++i) {
#pragma unroll
    for (int j = 0; j < NP; ++j) kf[i][j] = make_float2(0.f, 0.f);
    if (c < C) {
      if (NV == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w9c + (int64_t)(8 - i) * C + c));
        kf[i][0] = make_float2(a.x, a.y); kf[i][1 % NP] = make_float2(a.z, a.w);
      } else {
        kf[i][0] = __ldg(reinterpret_cast<const float2*>(w9c + (int64_t)(8 - i) * C + c));
      }
    }
  }
  uint32_t seed = 0u;
  if (DROP) seed = drop_seed(dp);
  const float2 zero2 = make_float2(0.f, 0.f);
  // AFFINE: the x stream is the producer's pre-BN output z; the activation y = max(z*scale + shift, 0) is formed on load.
  // Columns past W and channels past C get scale = shift = 0, i.e. y = 0 there (what the TMA zero fill gives without AFFINE).
  float2 asc[NP], ash[NP];
  if (AFFINE) {
#pragma unroll
    for (int j = 0; j < NP; ++j) { asc[j] = zero2; ash[j] = zero2; }
    if (live) {
      if (NV == 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(x_scale + c)), b = __ldg(reinterpret_cast<const float4*>(x_shift + c));
        asc[0] = make_float2(a.x, a.y); asc[1 % NP] = make_float2(a.z, a.w); ash[0] = make_float2(b.x, b.y); ash[1 % NP] = make_float2(b.z, b.w);
      } else {
        asc[0] = __ldg(reinterpret_cast<const float2*>(x_scale + c)); ash[0] = __ldg(reinterpret_cast<const float2*>(x_shift + c));
      }
    }
  }
  float2 acc[9][NP], prev[CW][NP], cur[CW][NP], xm[CW][NP], x0[CW][NP], s1[NP], s2[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
#pragma unroll
    for (int i = 0; i < 9; ++i) acc[i][j] = zero2;
#pragma unroll
    for (int q = 0; q < CW; ++q) prev[q][j] = cur[q][j] = xm[q][j] = x0[q][j] = zero2;
    s1[j] = s2[j] = zero2;
  }
  T* optr = dx + (((int64_t)n * H + (h0 - 2)) * W + col0) * lddx + c;   // advanced one row per dy row
  const int64_t orow = (int64_t)W * lddx;
  const bool redirect = UP && c0 < up.c_end;      // CTA-uniform
  // un-pixel-shuffled destination of pixel (n, hh, ww), channel c: ((n*H/2 + hh/2) * W/2 + ww/2) * 4F + ((hh%2)*2 + ww%2) * F + c
  T* const up_base = reinterpret_cast<T*>(up.out);
  const int up_nh = n * (H >> 1), up_rowstride = (W >> 1) * 4 * up.c_end;        // 2*W*F < 2^31 (checked by the launcher)
  int up_col[CW];
#pragma unroll
  for (int q = 0; q < CW; ++q) up_col[q] = ((col0 + q) >> 1) * 4 * up.c_end + ((col0 + q) & 1) * up.c_end + c;
  const uint32_t off = (uint32_t)(px * CW) * 128u + (uint32_t)cg * 8u;
  const uint32_t smem_base = smem_u32(smem);
  int t = h0 - 1;                                   // dy row being consumed; the x row of the same stage slot is t + 1

  // EDGE stages (the first and the last two of a strip) carry the segment predicates; the stages in between never need them
  auto stage_rows = [&](uint32_t sd, auto edge_tag) {
    constexpr bool EDGE = decltype(edge_tag)::value;
    const uint32_t sx = sd + Cfg::kDBytes;
    constexpr int HR = RH / CW;                     // rows whose shared loads are issued together (register budget)
#pragma unroll
    for (int rb = 0; rb < RH; rb += HR) {
    uint2 rd[HR][NQ], rx[HR][CW];
#pragma unroll
    for (int rr = 0; rr < HR; ++rr) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) rd[rr][q] = lds64(sd + (rb + rr) * (TW + 2) * 128 + q * 128);
#pragma unroll
      for (int q = 0; q < CW; ++q) rx[rr][q] = lds64(sx + (rb + rr) * TW * 128 + q * 128);
    }
#pragma unroll
    for (int rr = 0; rr < HR; ++rr, ++t, optr += orow) {
      float2 d[NQ][NP], xp[CW][NP];
#pragma unroll
      for (int q = 0; q < NQ; ++q) unpack8<T>(rd[rr][q], d[q]);
#pragma unroll
      for (int q = 0; q < CW; ++q) unpack8<T>(rx[rr][q], xp[q]);
      if (AFFINE) {
#pragma unroll
        for (int q = 0; q < CW; ++q)
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const float2 ty = fma2(xp[q][j], asc[j], ash[j]);
            xp[q][j] = make_float2(fmaxf(ty.x, 0.f), fmaxf(ty.y, 0.f));
          }
      }
      const bool xp_dead = EDGE && !((t + 1 >= h0) && (t + 1 < h1));   // x rows of other segments belong to other strips
#pragma unroll
      for (int q = 0; q < CW; ++q)
#pragma unroll
        for (int j = 0; j < NP; ++j)
          if (xp_dead) xp[q][j] = zero2;
      // All 18 FMAs per (column, channel pair) of this dy row, ordered by the dy column they read (d[c] is the second operand
      // of every one of them): 6 / 12 / 12 / 6 order-pinned FFMA2s per dy column with CW = 2, so 32 of 36 take that operand
      // from the operand-reuse cache.  Same summation order per accumulator as the straightforward nesting.
      //   weight gradient: dw[r][s] += x[t+r-1][w] * dy[t][w-s+1]   (d[q], d[q+1], d[q+2] = dy at columns w-1, w, w+1)
      //   data gradient:   rows t-1 (vv, complete after this row), t (prev), t+1 (cur) of dx gain flipped-kernel rows 2, 1, 0
      float2 vv[CW][NP], pn[CW][NP], cn[CW][NP];
#pragma unroll
      for (int j = 0; j < NP; ++j) {
#pragma unroll
        for (int q = 0; q < CW; ++q) { vv[q][j] = prev[q][j]; pn[q][j] = cur[q][j]; cn[q][j] = zero2; }
#pragma unroll
        for (int c = 0; c < NQ; ++c)
#pragma unroll
          for (int q = 0; q < CW; ++q) {
            const int sft = c - q;                  // tap column
            if (sft < 0 || sft > 2) continue;
            vv[q][j] = fma2v(kf[6 + sft][j], d[c][j], vv[q][j]);
            pn[q][j] = fma2v(kf[3 + sft][j], d[c][j], pn[q][j]);
            cn[q][j] = sft == 0 ? mul2v(kf[0][j], d[c][j]) : fma2v(kf[sft][j], d[c][j], cn[q][j]);
            acc[2 - sft][j] = fma2v(xm[q][j], d[c][j], acc[2 - sft][j]);
            acc[5 - sft][j] = fma2v(x0[q][j], d[c][j], acc[5 - sft][j]);
            acc[8 - sft][j] = fma2v(xp[q][j], d[c][j], acc[8 - sft][j]);
          }
      }
      if ((!EDGE || (t > h0 && t <= h1)) && live) {   // dx row t-1 is complete once flipped-kernel row 2 has seen dy row t
        T* up_row = nullptr;
        if (UP && redirect)     // row (n, (t-1)/2) of the un-pixel-shuffled operand, column block ((t-1)%2, .): one wide multiply per row
          up_row = up_base + (int64_t)(up_nh + ((t - 1) >> 1)) * up_rowstride + ((t - 1) & 1) * 2 * up.c_end;
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          float o[NV];
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            float2 v = vv[q][j];
            if (RELU_MASK) { v.x = xm[q][j].x > 0.f ? v.x : 0.f; v.y = xm[q][j].y > 0.f ? v.y : 0.f; }   // xm = x[t-1] = y of the producer
            o[2 * j] = v.x; o[2 * j + 1] = v.y;
          }
          if (DROP && c0 >= drop_c_from) {          // CTA-uniform: a 128-byte channel block lies on one side of drop_c_from
            const uint64_t base = (uint64_t)(((int64_t)n * H + (t - 1)) * W + (col0 + q)) * dp.ctot + dp.c0 + c;
            dropout_apply(o, base, seed, dp.keep, dp.inv_keep);
          }
          const uint2 packed = pack8(o, (T*)nullptr);
          if (RELU_MASK) {                          // the reductions see the values as stored
            float2 g[NP];
            unpack8<T>(packed, g);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              s1[j] = add2(s1[j], g[j]);
              s2[j] = fma2(g[j], xm[q][j], s2[j]);
            }
          }
          if (UP && redirect) {                     // un-pixel-shuffled store + bias-gradient sums (see UpArgs)
            *reinterpret_cast<uint2*>(up_row + up_col[q]) = packed;
#pragma unroll
            for (int j = 0; j < NP; ++j) s1[j] = add2(s1[j], make_float2(o[2 * j], o[2 * j + 1]));   // fp32 values before rounding
          } else {
            *reinterpret_cast<uint2*>(optr + q * lddx) = packed;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < CW; ++q)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          prev[q][j] = pn[q][j]; cur[q][j] = cn[q][j];
          xm[q][j] = x0[q][j]; x0[q][j] = xp[q][j];
        }
    }
    }
  };

  for (int k = 0; k < nst; ++k) {
    const int s = k % S;
    if (threadIdx.x == 0 && k + S - 1 < nst) issue(k + S - 1);
    mbar_wait(&full_bar[s], (k / S) & 1);
    const uint32_t sd = smem_base + s * Cfg::kStageBytes + off;
    if (k == 0 || k >= nst - 2) stage_rows(sd, std::true_type{});
    else stage_rows(sd, std::false_type{});
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }

  // lanes l and l^16 hold the same channels of neighbouring columns
  auto fold = [&](float mine, int slot, int j) {
    const float v = mine + __shfl_xor_sync(0xffffffffu, mine, 16);
    if (lane < 16) red_shared_add(smem_u32(&s_acc[slot * Cfg::CB + cg * NV + j]), v);
  };
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int j = 0; j < NV; ++j) fold((j & 1) ? acc[i][j / 2].y : acc[i][j / 2].x, i, j);
  if (RELU_MASK) {
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      fold((j & 1) ? s1[j / 2].y : s1[j / 2].x, 9, j);
      fold((j & 1) ? s2[j / 2].y : s2[j / 2].x, 10, j);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * Cfg::CB; i += blockDim.x) {
    const int tap = i / Cfg::CB, ch = c0 + i % Cfg::CB;
    if (ch < C) atomicAdd(&dw9c[(int64_t)tap * C + ch], s_acc[i]);
  }
  if (UP && redirect && up.colsum) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) fold((j & 1) ? s1[j / 2].y : s1[j / 2].x, 9, j);
    __syncthreads();
    for (int i = threadIdx.x; i < Cfg::CB; i += blockDim.x)
      if (c0 + i < up.c_end) atomicAdd(&up.colsum[c0 + i], s_acc[9 * Cfg::CB + i]);
  }
  if (RELU_MASK && bn_sums) {
    for (int i = threadIdx.x; i < 2 * Cfg::CB; i += blockDim.x) {
      const int which = i / Cfg::CB, ch = c0 + i % Cfg::CB;
      if (ch < C) atomicAdd(&bn_sums[(int64_t)which * C + ch], s_acc[9 * Cfg::CB + i]);
    }
  }
}

template <typename T> static bool dw_strip_ok(const void* a, int64_t lda, const void* b, int64_t ldb, int C) {
  constexpr int kNV = 8 / (int)sizeof(T);
  return (C % kNV == 0) && C >= 8 && ((lda * sizeof(T)) % 16 == 0) && ((ldb * sizeof(T)) % 16 == 0) && aligned16(a) && aligned16(b);
}

template <typename T, int CW, int PXT>
static int dw_bwd_strip_launch_cfg(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* w9c, void* dx, int64_t lddx,
                                   float* dw9c, float* bn_sums, int relu_mask, int N, int H, int W, int C, DropArgs dp, int drop_c_from,
                                   const float* x_scale, const float* x_shift, UpArgs up, cudaStream_t st) {
  using Cfg = BwCfg<T, CW, PXT>;
  static_assert(Cfg::S >= 3, "dwconv3x3_bwd: the stage ring needs at least 3 stages");
  CUtensorMap tmD, tmX;
  if (int e = make_nhwc_tmap<T>(&tmD, dy, lddy, N, H, W, C, Cfg::TW + 2, Cfg::RH, "dwconv3x3_bwd(dy)")) return e;
  if (int e = make_nhwc_tmap<T>(&tmX, x, ldx, N, H, W, C, Cfg::TW, Cfg::RH, "dwconv3x3_bwd(x)")) return e;
  static SmemAttrOnce o000, o010, o100, o110, o001, o011;
  cudaError_t ea = ensure_dynamic_smem(o000, dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, false, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(o010, dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, true, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(o100, dwconv3x3_bwd_strip_kernel<T, CW, PXT, true, false, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(o110, dwconv3x3_bwd_strip_kernel<T, CW, PXT, true, true, false>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(o001, dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, false, true>, Cfg::kSmemBytes);
  if (ea == cudaSuccess) ea = ensure_dynamic_smem(o011, dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, true, true>, Cfg::kSmemBytes);
  if (ea != cudaSuccess) return set_cuda_error(ea, "dwconv3x3_bwd: cudaFuncSetAttribute");
  const int ntw = (int)ceil_div(W, Cfg::TW), ncb = (int)ceil_div(C, Cfg::CB);
  const int seg = pick_seg_rows(N, H, ntw, ncb, 32, (int64_t)sm_count() * 4 * Cfg::kMinBlocks);
  const int nseg = (int)ceil_div(H, seg);
  const int64_t items = (int64_t)N * nseg * ntw * ncb;
  UNET_REQUIRE(items < ((int64_t)1 << 31), UNET_EUNSUPPORTED, "dwconv3x3_bwd: too many strips");
#define UNET_BW_LAUNCH(D, M, A) launch_pdl(dwconv3x3_bwd_strip_kernel<T, CW, PXT, D, M, A>, (unsigned)items, Cfg::kThreads, Cfg::kSmemBytes, st, \
      tmD, tmX, w9c, (T*)dx, lddx, dw9c, bn_sums, H, W, C, seg, nseg, ntw, ncb, dp, drop_c_from, x_scale, x_shift, up)
  if (up.out) {       // validated by the caller: no ReLU mask, no x affine
    static SmemAttrOnce u0, u1;
    ea = ensure_dynamic_smem(u0, dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, false, false, true>, Cfg::kSmemBytes);
    if (ea == cudaSuccess) ea = ensure_dynamic_smem(u1, dwconv3x3_bwd_strip_kernel<T, CW, PXT, true, false, false, true>, Cfg::kSmemBytes);
    if (ea != cudaSuccess) return set_cuda_error(ea, "dwconv3x3_bwd: cudaFuncSetAttribute");
    if (dp.on) launch_pdl(dwconv3x3_bwd_strip_kernel<T, CW, PXT, true, false, false, true>, (unsigned)items, Cfg::kThreads, Cfg::kSmemBytes, st, tmD, tmX, w9c, (T*)dx, lddx, dw9c, bn_sums, H, W, C, seg, nseg, ntw, ncb, dp, drop_c_from, x_scale, x_shift, up);
    else launch_pdl(dwconv3x3_bwd_strip_kernel<T, CW, PXT, false, false, false, true>, (unsigned)items, Cfg::kThreads, Cfg::kSmemBytes, st, tmD, tmX, w9c, (T*)dx, lddx, dw9c, bn_sums, H, W, C, seg, nseg, ntw, ncb, dp, drop_c_from, x_scale, x_shift, up);
  }
  else if (x_scale) { if (relu_mask) UNET_BW_LAUNCH(false, true, true); else UNET_BW_LAUNCH(false, false, true); }
  else if (dp.on) { if (relu_mask) UNET_BW_LAUNCH(true, true, false); else UNET_BW_LAUNCH(true, false, false); }
  else            { if (relu_mask) UNET_BW_LAUNCH(false, true, false); else UNET_BW_LAUNCH(false, false, false); }
#undef UNET_BW_LAUNCH
  UNET_LAUNCH_CHECK("dwconv3x3_bwd(strip)");
  return UNET_OK;
}

template <typename T>
static int dw_bwd_strip_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* w9c, void* dx, int64_t lddx,
                               float* dw9c, float* bn_sums, int relu_mask, int N, int H, int W, int C, DropArgs dp, int drop_c_from,
                               const float* x_scale, const float* x_shift, UpArgs up, cudaStream_t st) {
  UNET_REQUIRE(!(x_scale && dp.on), UNET_EUNSUPPORTED, "dwconv3x3_bwd: x affine and dropout cannot be combined");
#define UNET_BW_CFG(CW_, PXT_) dw_bwd_strip_launch_cfg<T, CW_, PXT_>(x, ldx, dy, lddy, w9c, dx, lddx, dw9c, bn_sums, relu_mask, N, H, W, C, \
                                                                   dp, drop_c_from, x_scale, x_shift, up, st)
  // even widths: 2 columns per thread, strips of 16 columns, 3 CTAs of 128 threads per SM (measured best of {1,2} columns x
  // {128,192,384} threads: 5.1 -> 6.5 TB/s plain, 4.4 -> 5.4 masked at 64 x 512 x 512 x 64); otherwise 1 column, 384 threads
  if (W % 2 == 0 && W >= 16) return UNET_BW_CFG(2, 8);
  return UNET_BW_CFG(1, 24);
#undef UNET_BW_CFG
}

// C <= 4 (the RGB input image): one thread per (image, row segment, column), all channels, 9*C register accumulators,
// warp shuffle -> shared -> one global atomic per (tap, channel) per block.
template <typename T, int CC>
__global__ void __launch_bounds__(256)
dwconv3x3_bwd_weight_smallc_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                                   float* __restrict__ dw9c, int N, int H, int W, int R, int nseg) {
  pdl_enter();
  __shared__ float s_acc[9 * CC];
  if (threadIdx.x < 9 * CC) s_acc[threadIdx.x] = 0.f;
  __syncthreads();
  float acc[9][CC];
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int c = 0; c < CC; ++c) acc[i][c] = 0.f;
  const int64_t n_items = (int64_t)N * nseg * W;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * blockDim.x) {
    const int wq = (int)(item % W); const int64_t t = item / W;
    const int hs = (int)(t % nseg); const int64_t n = t / nseg;
    const int h0 = hs * R, h1 = min(H, h0 + R);
    const bool has_l = wq > 0, has_r = wq + 1 < W;
    const T* xcol = x + ((n * H) * (int64_t)W + wq) * ldx;
    const T* dcol = dy + ((n * H) * (int64_t)W + wq) * lddy;
    const int64_t xrow = (int64_t)W * ldx, drow = (int64_t)W * lddy;
    float dm[CC], d0[CC], dp[CC];
#pragma unroll
    for (int c = 0; c < CC; ++c) { dm[c] = 0.f; d0[c] = 0.f; dp[c] = h0 < h1 ? to_f32(dcol[(int64_t)h0 * drow + c]) : 0.f; }
    for (int q = h0 - 1; q <= h1; ++q) {
      if (q >= 0 && q < H) {
        const T* p = xcol + q * xrow;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
          const float b = to_f32(p[c]);
          const float a = has_l ? to_f32(p[c - ldx]) : 0.f;
          const float cc = has_r ? to_f32(p[c + ldx]) : 0.f;
          acc[0][c] = fmaf(a, dp[c], acc[0][c]); acc[1][c] = fmaf(b, dp[c], acc[1][c]); acc[2][c] = fmaf(cc, dp[c], acc[2][c]);
          acc[3][c] = fmaf(a, d0[c], acc[3][c]); acc[4][c] = fmaf(b, d0[c], acc[4][c]); acc[5][c] = fmaf(cc, d0[c], acc[5][c]);
          acc[6][c] = fmaf(a, dm[c], acc[6][c]); acc[7][c] = fmaf(b, dm[c], acc[7][c]); acc[8][c] = fmaf(cc, dm[c], acc[8][c]);
        }
      }
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        dm[c] = d0[c]; d0[c] = dp[c];
        dp[c] = (q + 2 < h1) ? to_f32(dcol[(int64_t)(q + 2) * drow + c]) : 0.f;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9; ++i)
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      const float v = warp_sum(acc[i][c]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&s_acc[i * CC + c], v);
    }
  __syncthreads();
  if (threadIdx.x < 9 * CC) atomicAdd(&dw9c[threadIdx.x], s_acc[threadIdx.x]);
}

template <typename T, int CC>
static void dw_bwd_weight_smallc_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c,
                                        int N, int H, int W, cudaStream_t st) {
  const int R = 16;
  const int nseg = (int)ceil_div(H, R);
  const int64_t items = (int64_t)N * nseg * W;
  const unsigned grid = (unsigned)i64min(ceil_div(items, 256), (int64_t)sm_count() * 8);
  launch_pdl(dwconv3x3_bwd_weight_smallc_kernel<T, CC>, grid, 256, 0, st, (const T*)x, ldx, (const T*)dy, lddy, dw9c, N, H, W, R, nseg);
}

template <typename T>
static int dw_bwd_weight_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c,
                                int N, int H, int W, int C, cudaStream_t st) {
  const int cv = C / 4;
  const bool vec = (C % 4 == 0) && (ldx % 8 == 0) && (lddy % 8 == 0) && aligned16(x) && aligned16(dy) &&
                   (256 % cv == 0 || cv % 256 == 0) && (9 * C * 4 <= 160 * 1024);
  constexpr int kNV = 8 / (int)sizeof(T);
  if ((C % kNV == 0) && ((ldx * sizeof(T)) % 16 == 0) && ((lddy * sizeof(T)) % 16 == 0) && aligned16(x) && aligned16(dy) && C >= 8)
    return dw_wgrad_strip_launch<T>(x, ldx, dy, lddy, dw9c, N, H, W, C, st);
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
    launch_pdl(dwconv3x3_bwd_weight_vec4_kernel<T>, (unsigned)grid, 256, smem, st, (const T*)x, ldx, (const T*)dy, lddy, dw9c,
                                                                           N, H, W, C, R, nseg);
  } else if (C <= 4) {
    switch (C) {
      case 1: dw_bwd_weight_smallc_launch<T, 1>(x, ldx, dy, lddy, dw9c, N, H, W, st); break;
      case 2: dw_bwd_weight_smallc_launch<T, 2>(x, ldx, dy, lddy, dw9c, N, H, W, st); break;
      case 3: dw_bwd_weight_smallc_launch<T, 3>(x, ldx, dy, lddy, dw9c, N, H, W, st); break;
      default: dw_bwd_weight_smallc_launch<T, 4>(x, ldx, dy, lddy, dw9c, N, H, W, st); break;
    }
  } else {
    UNET_REQUIRE(C <= 64, UNET_EUNSUPPORTED, "dwconv3x3_bwd_weight: C=%d needs C%%4==0 and 16B-aligned views", C);
    launch_pdl(dwconv3x3_bwd_weight_scalar_kernel<T>, 9 * C, 256, 0, st, (const T*)x, ldx, (const T*)dy, lddy, dw9c, N, H, W, C);
  }
  UNET_LAUNCH_CHECK("dwconv3x3_bwd_weight");
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy,
                                  int N, int H, int W, int C, int dtype, int flip,
                                  const float* in_scale, const float* in_shift,
                                  const unet_dropout* drop, float* colsum, void* stream) {
  UNET_REQUIRE(x && w9c && y, UNET_EINVAL, "dwconv3x3_fwd: null pointer");
  UNET_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, UNET_EINVAL, "dwconv3x3_fwd: bad dims %d %d %d %d", N, H, W, C);
  UNET_REQUIRE(ldx >= C && ldy >= C, UNET_EINVAL, "dwconv3x3_fwd: ld < C");
  UNET_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), UNET_EINVAL, "dwconv3x3_fwd: scale/shift must come together");
  const DropArgs dp = make_drop(drop);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32)  return dw_fwd_launch<float>(x, ldx, w9c, y, ldy, N, H, W, C, flip, in_scale, in_shift, dp, colsum, st);
  if (dtype == UNET_BF16) return dw_fwd_launch<__nv_bfloat16>(x, ldx, w9c, y, ldy, N, H, W, C, flip, in_scale, in_shift, dp, colsum, st);
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

extern "C" int unet_dwconv3x3_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* w9c,
                                  void* dx, int64_t lddx, float* dw9c, int N, int H, int W, int C, int dtype,
                                  int relu_mask, float* bn_sums, const unet_dropout* drop, int drop_c_from,
                                  const float* x_scale, const float* x_shift,
                                  void* up_out, int up_c, float* up_colsum, void* stream) {
  UNET_REQUIRE(x && dy && w9c && dx && dw9c, UNET_EINVAL, "dwconv3x3_bwd: null pointer");
  const int cblk = 128 / (dtype == UNET_F32 ? 4 : 2);
  UNET_REQUIRE(!up_out || (!relu_mask && !x_scale && up_c > 0 && up_c <= C && up_c % cblk == 0 && H % 2 == 0 && W % 2 == 0 && aligned16(up_out)), UNET_EINVAL,
               "dwconv3x3_bwd: up_out needs relu_mask=0, no x affine, even H and W, and up_c a multiple of the 128-byte channel block inside C");
  UNET_REQUIRE(up_out || !up_colsum, UNET_EINVAL, "dwconv3x3_bwd: up_colsum needs up_out");
  UNET_REQUIRE(!up_out || ((int64_t)2 * W * up_c < ((int64_t)1 << 31) && (int64_t)N * (H / 2) < ((int64_t)1 << 31)), UNET_EUNSUPPORTED,
               "dwconv3x3_bwd: up_out row stride overflows 32 bits");
  const UpArgs up{up_out, up_out ? up_c : 0, up_colsum};
  UNET_REQUIRE((x_scale == nullptr) == (x_shift == nullptr), UNET_EINVAL, "dwconv3x3_bwd: x_scale/x_shift must come together");
  UNET_REQUIRE(!x_scale || (aligned16(x_scale) && aligned16(x_shift)), UNET_EALIGN, "dwconv3x3_bwd: x_scale/x_shift must be 16B aligned");
  UNET_REQUIRE(drop_c_from >= 0 && drop_c_from % (128 / (dtype == UNET_F32 ? 4 : 2)) == 0, UNET_EINVAL,
               "dwconv3x3_bwd: drop_c_from must be a multiple of the 128-byte channel block");
  UNET_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, UNET_EINVAL, "dwconv3x3_bwd: bad dims %d %d %d %d", N, H, W, C);
  UNET_REQUIRE(ldx >= C && lddy >= C && lddx >= C, UNET_EINVAL, "dwconv3x3_bwd: ld < C");
  UNET_REQUIRE(!bn_sums || relu_mask, UNET_EINVAL, "dwconv3x3_bwd: bn_sums needs relu_mask");
  const DropArgs dp = make_drop(drop);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UNET_F32) {
    UNET_REQUIRE(dw_strip_ok<float>(x, ldx, dy, lddy, C) && dw_strip_ok<float>(dx, lddx, dy, lddy, C), UNET_EUNSUPPORTED,
                 "dwconv3x3_bwd: needs C%%2==0, C>=8 and 16B-aligned views (use dwconv3x3_fwd(flip) + dwconv3x3_bwd_weight)");
    return dw_bwd_strip_launch<float>(x, ldx, dy, lddy, w9c, dx, lddx, dw9c, bn_sums, relu_mask, N, H, W, C, dp, drop_c_from, x_scale, x_shift, up, st);
  }
  if (dtype == UNET_BF16) {
    UNET_REQUIRE(dw_strip_ok<__nv_bfloat16>(x, ldx, dy, lddy, C) && dw_strip_ok<__nv_bfloat16>(dx, lddx, dy, lddy, C), UNET_EUNSUPPORTED,
                 "dwconv3x3_bwd: needs C%%4==0, C>=8 and 16B-aligned views (use dwconv3x3_fwd(flip) + dwconv3x3_bwd_weight)");
    return dw_bwd_strip_launch<__nv_bfloat16>(x, ldx, dy, lddy, w9c, dx, lddx, dw9c, bn_sums, relu_mask, N, H, W, C, dp, drop_c_from, x_scale, x_shift, up, st);
  }
  return set_error(UNET_EINVAL, "dwconv3x3_bwd: bad dtype %d", dtype);
}
