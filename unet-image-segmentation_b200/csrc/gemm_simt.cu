// fp32-exact dense contraction on CUDA cores.  This is the arithmetic of the fp32 parity mode (north star: fp32
// probabilities within 1e-4 of the reference, which tf32/bf16 tensor-core inputs cannot give) and the path for
// shapes the tensor-core kernel does not take (K = 3 first pointwise layer).  It serves the pointwise half of
// SeparableConv2D (reference model/u_net.py:14-20), Conv2DTranspose (:88-94) and their data / weight gradients.
//
// 64x64x16 tiles, 256 threads, 4x4 register micro-tiles, split-K with atomic accumulation for weight gradients.
#include "common.cuh"

namespace unet {

constexpr int BM = 64, BN = 64, BK = 16;

struct SimtParams {
  int64_t M, N, K;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* C; int64_t ldc;
  int a_trans, b_trans, accumulate, epilogue;
  const float* scale; const float* shift;
  double* colsum; double* colsq;
  int convt_H, convt_W; int64_t convt_cout;
  float keep, inv_keep; uint32_t seed; int drop_on; int64_t ctot, c0; const uint32_t* seed_dev;
  int64_t k_per_split;
};

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const SimtParams p) {
  pdl_enter();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float s_cs[BN], s_cq[BN];

  const TIn* __restrict__ A = (const TIn*)p.A;
  const TIn* __restrict__ B = (const TIn*)p.B;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t k_end = i64min(p.K, k_begin + p.k_per_split);

  if (p.epilogue == UNET_EPI_STATS && tid < BN) { s_cs[tid] = 0.f; s_cq[tid] = 0.f; }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int mm, kk;
      if (p.a_trans) { kk = idx >> 6; mm = idx & 63; } else { mm = idx >> 4; kk = idx & 15; }
      const int64_t gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < p.M && gk < k_end) v = to_f32(p.a_trans ? A[gk * p.lda + gm] : A[gm * p.lda + gk]);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int idx = tid + e * 256;
      int nn, kk;
      if (p.b_trans) { nn = idx >> 4; kk = idx & 15; } else { kk = idx >> 6; nn = idx & 63; }
      const int64_t gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < p.N && gk < k_end) v = to_f32(p.b_trans ? B[gn * p.ldb + gk] : B[gk * p.ldb + gn]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  TOut* __restrict__ C = (TOut*)p.C;
  float cs[4] = {0, 0, 0, 0}, cq[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    int64_t convt_base = 0, convt_pix = 0;
    if (p.epilogue == UNET_EPI_CONVT) {
      const int64_t j = m % p.convt_W, q = m / p.convt_W;
      const int64_t ii = q % p.convt_H, img = q / p.convt_H;
      convt_pix = (img * 2 * p.convt_H + 2 * ii) * (2 * p.convt_W) + 2 * j;   // pixel (2i, 2j) of the upsampled image
      convt_base = convt_pix;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      switch (p.epilogue) {
        case UNET_EPI_AFFINE:
        case UNET_EPI_AFFINE_RELU:
          v = fmaf(v, p.scale ? p.scale[n] : 1.f, p.shift ? p.shift[n] : 0.f);
          if (p.epilogue == UNET_EPI_AFFINE_RELU) v = fmaxf(v, 0.f);
          break;
        default: break;
      }
      if (p.epilogue == UNET_EPI_CONVT) {
        const int64_t ab = n / p.convt_cout, co = n % p.convt_cout;
        const int64_t pix = convt_base + (ab >> 1) * (2 * p.convt_W) + (ab & 1);
        v += p.shift ? p.shift[co] : 0.f;
        if (p.drop_on) v *= dropout_mult((uint64_t)pix * p.ctot + p.c0 + co, p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0u), p.keep, p.inv_keep);
        C[pix * p.ldc + co] = from_f32<TOut>(v);
      } else if (p.accumulate) {
        atomicAdd(reinterpret_cast<float*>(p.C) + m * p.ldc + n, v);
      } else {
        C[m * p.ldc + n] = from_f32<TOut>(v);
        if (p.epilogue == UNET_EPI_STATS) { const float r = round_to<TOut>(v); cs[j] += r; cq[j] = fmaf(r, r, cq[j]); }
      }
    }
  }
  if (p.epilogue == UNET_EPI_STATS) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { atomicAdd(&s_cs[tx * 4 + j], cs[j]); atomicAdd(&s_cq[tx * 4 + j], cq[j]); }
    __syncthreads();
    if (tid < BN && n0 + tid < p.N) {
      atomicAdd(&p.colsum[n0 + tid], (double)s_cs[tid]);
      atomicAdd(&p.colsq[n0 + tid], (double)s_cq[tid]);
    }
  }
}

int gemm_validate(const unet_gemm_args* a, const char* who) {
  UNET_REQUIRE(a, UNET_EINVAL, "%s: null args", who);
  UNET_REQUIRE(a->A && a->B && (a->C || a->epilogue == UNET_EPI_HEAD), UNET_EINVAL, "%s: null operand", who);
  UNET_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, UNET_EINVAL, "%s: bad shape %lld %lld %lld", who,
               (long long)a->M, (long long)a->N, (long long)a->K);
  UNET_REQUIRE(a->lda >= (a->a_trans ? a->M : (a->A2 ? a->k_split : a->K)), UNET_EINVAL, "%s: lda too small", who);
  UNET_REQUIRE(a->ldb >= (a->b_trans ? a->K : (a->B2 ? a->n_split : a->N)), UNET_EINVAL, "%s: ldb too small", who);
  UNET_REQUIRE(a->in_dtype == UNET_F32 || a->in_dtype == UNET_BF16, UNET_EINVAL, "%s: bad in_dtype", who);
  UNET_REQUIRE(a->out_dtype == UNET_F32 || a->out_dtype == UNET_BF16, UNET_EINVAL, "%s: bad out_dtype", who);
  UNET_REQUIRE(a->epilogue >= UNET_EPI_NONE && a->epilogue <= UNET_EPI_HEAD, UNET_EINVAL, "%s: bad epilogue", who);
  if (a->epilogue == UNET_EPI_HEAD)
    UNET_REQUIRE(a->head_w && a->head_out && a->head_classes >= 1 && a->head_classes <= 8 && a->N <= 64, UNET_EINVAL,
                 "%s: HEAD needs head_w, head_out, 1 <= classes <= 8 and N <= 64", who);
  UNET_REQUIRE(!a->accumulate || (a->out_dtype == UNET_F32 && a->epilogue == UNET_EPI_NONE), UNET_EINVAL,
               "%s: accumulate needs fp32 output and no epilogue", who);
  UNET_REQUIRE(a->epilogue != UNET_EPI_STATS || (a->colsum && a->colsq), UNET_EINVAL, "%s: STATS needs colsum/colsq", who);
  if (a->epilogue == UNET_EPI_CONVT) {
    UNET_REQUIRE(a->convt_H > 0 && a->convt_W > 0 && a->N % 4 == 0, UNET_EINVAL, "%s: CONVT needs H,W and N%%4==0", who);
    UNET_REQUIRE(a->M % ((int64_t)a->convt_H * a->convt_W) == 0, UNET_EINVAL, "%s: CONVT M must be images*H*W", who);
    UNET_REQUIRE(a->ldc >= a->N / 4, UNET_EINVAL, "%s: CONVT ldc < Cout", who);
  } else {
    UNET_REQUIRE(a->C == nullptr || a->ldc >= a->N, UNET_EINVAL, "%s: ldc too small", who);
  }
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_gemm_simt(const unet_gemm_args* a, void* stream) {
  if (int e = gemm_validate(a, "gemm_simt")) return e;
  UNET_REQUIRE(a->epilogue != UNET_EPI_HEAD, UNET_EUNSUPPORTED, "gemm_simt: the fused output head exists on the tensor-core path only");
  UNET_REQUIRE(!a->A2 && !a->B2, UNET_EUNSUPPORTED, "gemm_simt: operand concatenation exists on the tensor-core path only");
  SimtParams p{};
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.A = a->A; p.lda = a->lda; p.B = a->B; p.ldb = a->ldb; p.C = a->C; p.ldc = a->ldc;
  p.a_trans = a->a_trans; p.b_trans = a->b_trans; p.accumulate = a->accumulate; p.epilogue = a->epilogue;
  p.scale = a->scale; p.shift = a->shift; p.colsum = a->colsum; p.colsq = a->colsq;
  p.convt_H = a->convt_H; p.convt_W = a->convt_W; p.convt_cout = a->N / 4;
  p.drop_on = 0; p.keep = 1.f; p.inv_keep = 1.f;
  if (a->epilogue == UNET_EPI_CONVT && a->drop.rate > 0.f) {
    p.drop_on = 1; p.keep = 1.f - a->drop.rate; p.inv_keep = 1.f / (1.f - a->drop.rate);
    p.seed = a->drop.seed; p.ctot = a->drop.ctot; p.c0 = a->drop.c0; p.seed_dev = a->drop.seed_dev;
  }
  const int64_t tiles = ceil_div(a->M, BM) * ceil_div(a->N, BN);
  int64_t splits = 1;
  if (a->accumulate) {
    splits = i64max(1, ((int64_t)sm_count() * 4) / tiles);
    splits = i64min(splits, i64max(1, a->K / 512));
    splits = i64min(splits, 65535);
  }
  p.k_per_split = ceil_div(ceil_div(a->K, splits), BK) * BK;
  splits = ceil_div(a->K, p.k_per_split);
  UNET_REQUIRE(ceil_div(a->N, BN) <= 65535, UNET_EUNSUPPORTED, "gemm_simt: N too large for grid.y");
  dim3 grid((unsigned)ceil_div(a->M, BM), (unsigned)ceil_div(a->N, BN), (unsigned)splits);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->in_dtype == UNET_F32 && a->out_dtype == UNET_F32)        launch_pdl(gemm_simt_kernel<float, float>, grid, 256, 0, st, p);
  else if (a->in_dtype == UNET_BF16 && a->out_dtype == UNET_BF16) launch_pdl(gemm_simt_kernel<__nv_bfloat16, __nv_bfloat16>, grid, 256, 0, st, p);
  else if (a->in_dtype == UNET_BF16 && a->out_dtype == UNET_F32)  launch_pdl(gemm_simt_kernel<__nv_bfloat16, float>, grid, 256, 0, st, p);
  else                                                            launch_pdl(gemm_simt_kernel<float, __nv_bfloat16>, grid, 256, 0, st, p);
  UNET_LAUNCH_CHECK("gemm_simt");
  return UNET_OK;
}
