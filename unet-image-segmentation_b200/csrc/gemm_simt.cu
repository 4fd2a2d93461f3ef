// fp32-exact dense contraction on CUDA cores.  This is the arithmetic of the fp32 parity mode (north star: fp32
// probabilities within 1e-4 of the reference, which tf32/bf16 tensor-core inputs cannot give) and the path for
// shapes the tensor-core kernel does not take (K = 3 first pointwise layer).  It serves the pointwise half of
// SeparableConv2D (reference model/u_net.py:14-20), Conv2DTranspose (:88-94) and their data / weight gradients.
//
// 128 x BN x 16 tiles (BN = 128, or 64 for the 64-channel layers), 256 threads, 8 x (BN/16) register micro-tiles split in
// 4-wide halves (conflict-free 16-byte shared loads), 16-byte global loads along whichever dimension is contiguous for the
// operand's layout (transposing stores into shared memory otherwise), global loads of tile k+1 in flight while tile k is
// multiplied (two shared-memory buffers, one barrier per k-step), split-K with atomic accumulation for weight gradients.
#include "common.cuh"

namespace unet {

constexpr int BM = 128, BK = 16, LDS_PAD = 4;

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
  int a_vec, b_vec;       // operand base and leading dimension allow 16-byte loads (fp32 only)
};

// One operand tile (ROWS x BK, ROWS = BM or BN) of op(X): element (r, k) = trans ? X[k*ld + r] : X[r*ld + k].  `trans` == the
// ROWS dimension is the contiguous one.  Each thread moves NV4 = ROWS*BK/1024 groups of 4 elements: global -> registers
// (issued one k-step ahead) -> shared [BK][ROWS + pad].
template <typename TIn, int ROWS>
struct TileMover {
  static constexpr int NV4 = ROWS * BK / 1024;
  float4 v[NV4];
  __device__ __forceinline__ void load(const TIn* __restrict__ X, int64_t ld, bool trans, bool vec, int64_t r0, int64_t rows_total,
                                       int64_t k0, int64_t k_end, int tid) {
#pragma unroll
    for (int e = 0; e < NV4; ++e) {
      const int idx = tid + e * 256;
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      if (trans) {                      // 4 consecutive rows at one k
        const int kk = idx / (ROWS / 4), rr = (idx % (ROWS / 4)) * 4;
        const int64_t gk = k0 + kk, gr = r0 + rr;
        if (gk < k_end) {
          const TIn* src = X + gk * ld + gr;
          if (vec && sizeof(TIn) == 4 && gr + 3 < rows_total) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (gr + i < rows_total) t[i] = to_f32(src[i]);
          }
        }
      } else {                          // 4 consecutive k of one row
        const int rr = idx / (BK / 4), kk = (idx % (BK / 4)) * 4;
        const int64_t gk = k0 + kk, gr = r0 + rr;
        if (gr < rows_total) {
          const TIn* src = X + gr * ld + gk;
          if (vec && sizeof(TIn) == 4 && gk + 3 < k_end) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(src));
            t[0] = q.x; t[1] = q.y; t[2] = q.z; t[3] = q.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (gk + i < k_end) t[i] = to_f32(src[i]);
          }
        }
      }
      v[e] = make_float4(t[0], t[1], t[2], t[3]);
    }
  }
  __device__ __forceinline__ void store(float (*S)[ROWS + LDS_PAD], bool trans, int tid) const {
#pragma unroll
    for (int e = 0; e < NV4; ++e) {
      const int idx = tid + e * 256;
      if (trans) {
        const int kk = idx / (ROWS / 4), rr = (idx % (ROWS / 4)) * 4;
        *reinterpret_cast<float4*>(&S[kk][rr]) = v[e];
      } else {
        const int rr = idx / (BK / 4), kk = (idx % (BK / 4)) * 4;
        S[kk][rr] = v[e].x; S[kk + 1][rr] = v[e].y; S[kk + 2][rr] = v[e].z; S[kk + 3][rr] = v[e].w;
      }
    }
  }
};

template <typename TIn, typename TOut, int BN>
__global__ void __launch_bounds__(256, 2)
gemm_simt_kernel(const SimtParams p) {
  pdl_enter();
  constexpr int TN = BN / 16;            // columns per thread: 8 (two 4-wide halves) or 4
  constexpr int NH = TN / 4;             // column halves
  __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
  __shared__ float s_cs[BN], s_cq[BN];

  const TIn* __restrict__ A = (const TIn*)p.A;
  const TIn* __restrict__ B = (const TIn*)p.B;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t k_begin = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t k_end = i64min(p.K, k_begin + p.k_per_split);

  if (p.epilogue == UNET_EPI_STATS && tid < BN) { s_cs[tid] = 0.f; s_cq[tid] = 0.f; }

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  TileMover<TIn, BM> ma;
  TileMover<TIn, BN> mb;
  const bool a_rows_contig = p.a_trans != 0, b_rows_contig = p.b_trans == 0;     // B tile rows are the N dimension
  if (k_begin < k_end) {
    ma.load(A, p.lda, a_rows_contig, p.a_vec != 0, m0, p.M, k_begin, k_end, tid);
    mb.load(B, p.ldb, b_rows_contig, p.b_vec != 0, n0, p.N, k_begin, k_end, tid);
    ma.store(As[0], a_rows_contig, tid);
    mb.store(Bs[0], b_rows_contig, tid);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t k0 = k_begin; k0 < k_end; k0 += BK, buf ^= 1) {
    const bool more = k0 + BK < k_end;
    if (more) {                                   // next tile's global loads fly while this tile is multiplied
      ma.load(A, p.lda, a_rows_contig, p.a_vec != 0, m0, p.M, k0 + BK, k_end, tid);
      mb.load(B, p.ldb, b_rows_contig, p.b_vec != 0, n0, p.N, k0 + BK, k_end, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float4 bq = *reinterpret_cast<const float4*>(&Bs[buf][kk][h * (BN / 2) + tx * 4]);
        b[4 * h] = bq.x; b[4 * h + 1] = bq.y; b[4 * h + 2] = bq.z; b[4 * h + 3] = bq.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      ma.store(As[buf ^ 1], a_rows_contig, tid);
      mb.store(Bs[buf ^ 1], b_rows_contig, tid);
    }
    __syncthreads();
  }

  TOut* __restrict__ C = (TOut*)p.C;
  float cs[TN], cq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { cs[j] = 0.f; cq[j] = 0.f; }
  const uint32_t seed = p.drop_on ? p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0u) : 0u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (m >= p.M) continue;
    int64_t convt_base = 0;
    if (p.epilogue == UNET_EPI_CONVT) {
      const int64_t j = m % p.convt_W, q = m / p.convt_W;
      const int64_t ii = q % p.convt_H, img = q / p.convt_H;
      convt_base = (img * 2 * p.convt_H + 2 * ii) * (2 * p.convt_W) + 2 * j;   // pixel (2i, 2j) of the upsampled image
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + (j >> 2) * (BN / 2) + tx * 4 + (j & 3);
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.epilogue == UNET_EPI_AFFINE || p.epilogue == UNET_EPI_AFFINE_RELU) {
        v = fmaf(v, p.scale ? p.scale[n] : 1.f, p.shift ? p.shift[n] : 0.f);
        if (p.epilogue == UNET_EPI_AFFINE_RELU) v = fmaxf(v, 0.f);
      }
      if (p.epilogue == UNET_EPI_CONVT) {
        const int64_t ab = n / p.convt_cout, co = n % p.convt_cout;
        const int64_t pix = convt_base + (ab >> 1) * (2 * p.convt_W) + (ab & 1);
        v += p.shift ? p.shift[co] : 0.f;
        if (p.drop_on) v *= dropout_mult((uint64_t)pix * p.ctot + p.c0 + co, seed, p.keep, p.inv_keep);
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
    for (int j = 0; j < TN; ++j) {
      const int col = (j >> 2) * (BN / 2) + tx * 4 + (j & 3);
      atomicAdd(&s_cs[col], cs[j]); atomicAdd(&s_cq[col], cq[j]);
    }
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
  const int bn = a->N > 64 ? 128 : 64;
  const int64_t tiles = ceil_div(a->M, BM) * ceil_div(a->N, bn);
  int64_t splits = 1;
  if (a->accumulate) {
    splits = i64max(1, ((int64_t)sm_count() * 4) / tiles);
    splits = i64min(splits, i64max(1, a->K / 512));
    splits = i64min(splits, 65535);
  }
  p.k_per_split = ceil_div(ceil_div(a->K, splits), BK) * BK;
  splits = ceil_div(a->K, p.k_per_split);
  UNET_REQUIRE(ceil_div(a->N, bn) <= 65535, UNET_EUNSUPPORTED, "gemm_simt: N too large for grid.y");
  p.a_vec = a->in_dtype == UNET_F32 && aligned16(a->A) && a->lda % 4 == 0;
  p.b_vec = a->in_dtype == UNET_F32 && aligned16(a->B) && a->ldb % 4 == 0;
  dim3 grid((unsigned)ceil_div(a->M, BM), (unsigned)ceil_div(a->N, bn), (unsigned)splits);
  cudaStream_t st = (cudaStream_t)stream;
#define UNET_SIMT(TI, TO) do { if (bn == 128) launch_pdl(gemm_simt_kernel<TI, TO, 128>, grid, 256, 0, st, p); \
                               else launch_pdl(gemm_simt_kernel<TI, TO, 64>, grid, 256, 0, st, p); } while (0)
  if (a->in_dtype == UNET_F32 && a->out_dtype == UNET_F32)        UNET_SIMT(float, float);
  else if (a->in_dtype == UNET_BF16 && a->out_dtype == UNET_BF16) UNET_SIMT(__nv_bfloat16, __nv_bfloat16);
  else if (a->in_dtype == UNET_BF16 && a->out_dtype == UNET_F32)  UNET_SIMT(__nv_bfloat16, float);
  else                                                            UNET_SIMT(float, __nv_bfloat16);
#undef UNET_SIMT
  UNET_LAUNCH_CHECK("gemm_simt");
  return UNET_OK;
}
