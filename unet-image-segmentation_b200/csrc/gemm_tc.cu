// Dense contractions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
// Serves the pointwise half of SeparableConv2D (reference model/u_net.py:14-20), Conv2DTranspose(k=2,s=2) (:88-94) and
// their data / weight gradients in bf16 mode, with BatchNormalization scale/shift (:23), ReLU (:25), training batch
// statistics, bias, the transposed-convolution pixel shuffle into the skip-concat buffer (:96) and Dropout (:98) fused
// into the epilogue.
//
// Two kernels:
//   gemm_tc_nt_kernel    C[M,N] = A[M,K] * Bt[N,K]^T      both operands K-major (forward, data gradient)
//   gemm_tc_wgrad_kernel C[Mo,No] += A[P,Mo]^T * B[P,No]   both operands MN-major (weight gradient, split over P)
//
// Warp roles (256 threads, 1 CTA/SM, persistent over output tiles):
//   warp 0   TMA producer   (one thread)      smem ring of kStages {A 128x64, B BLOCK_Nx64} bf16 tiles, SWIZZLE_128B
//   warp 1   MMA issuer     (one thread)      tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16, fp32 accumulate
//   warp 2   TMEM allocator                   2 accumulator stages x BLOCK_N columns
//   warps 4-7 epilogue                        tcgen05.ld -> fp32 staging slab in smem (per warp, padded) -> each lane owns
//                                             4 fixed columns: scale/shift/ReLU/bias/stats in registers -> coalesced stores
#include "common.cuh"
#include "ptx.cuh"

namespace unet {

int gemm_validate(const unet_gemm_args* a, const char* who);

// ================================================================================================ descriptors
// Shared-memory matrix descriptor (SM100 format, version 1), SWIZZLE_128B.
//   K-major  tile (rows x 64 bf16, 128 B per row): SBO = 1024 B (8 rows), LBO unused.
//   MN-major tile (64 k-rows x 64 bf16 per 8 KB chunk): SBO = 1024 B (8 k-rows), LBO = chunk stride.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;          // SWIZZLE_128B
  return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ================================================================================================ kernel parameters
struct TcParams {
  int64_t M, N, K;
  void* C; int64_t ldc;
  int epilogue, out_bf16, accumulate;
  const float* scale; const float* shift;
  double* colsum; double* colsq;
  int convt_H, convt_W; int64_t convt_cout;
  float keep, inv_keep; uint32_t seed; int drop_on; int64_t ctot, c0; const uint32_t* seed_dev;
  int64_t kb_per_split;     // wgrad: k-blocks per CTA along P
  int num_m_tiles, num_n_tiles;
};

constexpr int kBlockM = 128, kBlockK = 64;
constexpr int kSlabPitch = 64 * 4 + 16;        // bytes per staged row: 64 fp32 + 16 B pad (17 x 16 B: conflict-free)
constexpr int kSlabBytes = 32 * kSlabPitch;    // per epilogue warp
constexpr int kNumEpiWarps = 4;

template <int BLOCK_N> struct TcCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 5 : 7);
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kNumEpiWarps * kSlabBytes + 256;
};

// ------------------------------------------------------------------------------------------------ shared epilogue
// One 64-column chunk of one warp's 32 accumulator rows: TMEM -> slab -> (lane owns 4 fixed columns) -> global.
template <bool OUT_BF16>
__device__ __forceinline__ void epilogue_chunk(const TcParams& p, uint32_t tmem_addr, uint8_t* slab, int lane,
                                               int64_t row0 /*global row of slab row 0*/, int64_t n_base /*global col of chunk*/,
                                               float (&st_sum)[4], float (&st_sq)[4]) {
  // 1) accumulator rows -> staging slab (lane == row)
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_addr + half * 32, r);
    tmem_ld_wait();
    uint4* dst = reinterpret_cast<uint4*>(slab + lane * kSlabPitch + half * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
  }
  __syncwarp();

  // 2) lane owns columns [n_base + 4*seg, +4) of rows (it*2 + lane/16)
  const int seg = lane & 15, rsub = lane >> 4;
  const int64_t n = n_base + seg * 4;
  const bool col_ok = n < p.N;     // N is a multiple of 8: a 4-column group is all-or-nothing
  float sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
  uint32_t cv_ab = 0, cv_co = 0;
  if (col_ok) {
    if (p.epilogue == UNET_EPI_AFFINE || p.epilogue == UNET_EPI_AFFINE_RELU) {
      if (p.scale) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.scale + n)); sc[0] = t.x; sc[1] = t.y; sc[2] = t.z; sc[3] = t.w; }
      if (p.shift) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.shift + n)); sh[0] = t.x; sh[1] = t.y; sh[2] = t.z; sh[3] = t.w; }
    } else if (p.epilogue == UNET_EPI_CONVT) {
      cv_ab = (uint32_t)n / (uint32_t)p.convt_cout; cv_co = (uint32_t)n % (uint32_t)p.convt_cout;
      if (p.shift) { const float4 t = __ldg(reinterpret_cast<const float4*>(p.shift + cv_co)); sh[0] = t.x; sh[1] = t.y; sh[2] = t.z; sh[3] = t.w; }
    }
  }
  const bool relu = p.epilogue == UNET_EPI_AFFINE_RELU;
  const bool stats = p.epilogue == UNET_EPI_STATS;
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int r = it * 2 + rsub;
    const int64_t m = row0 + r;
    const float4 a = *reinterpret_cast<const float4*>(slab + r * kSlabPitch + seg * 16);
    if (!col_ok || m >= p.M) continue;
    float v[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = fmaf(v[j], sc[j], sh[j]);
      if (relu) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.accumulate) {
      float* dst = reinterpret_cast<float*>(p.C) + m * p.ldc + n;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(dst + j, v[j]);
      continue;
    }
    int64_t off;
    if (p.epilogue == UNET_EPI_CONVT) {
      // 32-bit index arithmetic (M < 2^31 is checked on the host): 64-bit div/mod here cost more than the tile's MMA
      const uint32_t m32 = (uint32_t)m, cw = (uint32_t)p.convt_W, ch = (uint32_t)p.convt_H;
      const uint32_t q = m32 / cw, jj = m32 - q * cw;
      const uint32_t img = q / ch, ii = q - img * ch;
      const int64_t pix = ((int64_t)img * 2 * ch + 2 * ii + (cv_ab >> 1)) * (2 * cw) + 2 * jj + (cv_ab & 1);
      if (p.drop_on) {
        const uint64_t base = (uint64_t)pix * p.ctot + p.c0 + cv_co;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] *= dropout_mult(base + j, p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0u), p.keep, p.inv_keep);
      }
      off = pix * p.ldc + cv_co;
    } else {
      off = m * p.ldc + n;
    }
    if (OUT_BF16) {
      uint2 o;
      o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.C) + off) = o;
      if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float rr = round_to<__nv_bfloat16>(v[j]); st_sum[j] += rr; st_sq[j] = fmaf(rr, rr, st_sq[j]); }
      }
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off) = make_float4(v[0], v[1], v[2], v[3]);
      if (stats) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { st_sum[j] += v[j]; st_sq[j] = fmaf(v[j], v[j], st_sq[j]); }
      }
    }
  }
  __syncwarp();   // slab is overwritten by the next chunk
}

// fold the two half-warps (same columns) and publish
__device__ __forceinline__ void stats_flush(const TcParams& p, int lane, int64_t n_base, float (&st_sum)[4], float (&st_sq)[4]) {
  const int seg = lane & 15;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    st_sum[j] += __shfl_xor_sync(0xffffffffu, st_sum[j], 16);
    st_sq[j] += __shfl_xor_sync(0xffffffffu, st_sq[j], 16);
  }
  if (lane < 16) {
    const int64_t n = n_base + seg * 4;
    if (n < p.N) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(p.colsum + n + j, (double)st_sum[j]);
        atomicAdd(p.colsq + n + j, (double)st_sq[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { st_sum[j] = 0.f; st_sq[j] = 0.f; }
}

// ------------------------------------------------------------------------------------------------ NT kernel
template <int BLOCK_N, bool OUT_BF16>
__global__ void __launch_bounds__(256, 1)
gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = BLOCK_N / 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* slabs = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + kNumEpiWarps * kSlabBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (int)((p.K + kBlockK - 1) / kBlockK);
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kNumEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kBlockK, m_blk * kBlockM, kEvictFirst);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N, kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N, false, false);
      int stage = 0; uint32_t phase = 0; int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes), 0, 1024);
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes), 0, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)   // +32 B per UMMA_K inside the 128 B swizzle atom
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;                       // == warp % 4: the TMEM lane quarter this warp may read
    uint8_t* slab = slabs + q * kSlabBytes;
    float st_sum[kChunks][4], st_sq[kChunks][4];
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
      for (int j = 0; j < 4; ++j) { st_sum[c][j] = 0.f; st_sq[c][j] = 0.f; }
    int it = 0, last_n_blk = -1;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
      if (p.epilogue == UNET_EPI_STATS && last_n_blk >= 0 && last_n_blk != n_blk) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) stats_flush(p, lane, (int64_t)last_n_blk * BLOCK_N + c * 64, st_sum[c], st_sq[c]);
      }
      last_n_blk = n_blk;
      const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N;
      const int64_t row0 = (int64_t)m_blk * kBlockM + q * 32;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int64_t n_base = (int64_t)n_blk * BLOCK_N + c * 64;
        epilogue_chunk<OUT_BF16>(p, t_addr + c * 64, slab, lane, row0, n_base, st_sum[c], st_sq[c]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (p.epilogue == UNET_EPI_STATS && last_n_blk >= 0) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c) stats_flush(p, lane, (int64_t)last_n_blk * BLOCK_N + c * 64, st_sum[c], st_sq[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------ weight-gradient kernel
// One CTA = one (128 x BLOCK_N) tile of C and one slice of the reduction dimension P; partial products are added to C
// with fp32 atomics.  Operands are MN-major: each 64-row (P) x 64-column box lands as an 8 KB SWIZZLE_128B chunk.
template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = BLOCK_N / 64;
  constexpr int kBoxBytes = 64 * 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* slabs = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + kNumEpiWarps * kSlabBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
  const int64_t total_kb = (p.K + kBlockK - 1) / kBlockK;
  const int64_t kb_begin = (int64_t)blockIdx.y * p.kb_per_split;
  const int64_t kb_end = min(total_kb, kb_begin + p.kb_per_split);
  const int num_k = (int)i64max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (num_k > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int prow = (int)((kb_begin + kb) * kBlockK);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem_a + stage * Cfg::kABytes + c * kBoxBytes, &tmA, &full_bar[stage], m_blk * kBlockM + c * 64, prow, kEvictFirst);
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            tma_load_2d(smem_b + stage * Cfg::kBBytes + c * kBoxBytes, &tmB, &full_bar[stage], n_blk * BLOCK_N + c * 64, prow, kEvictFirst);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(BLOCK_N, true, true);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes), kBoxBytes, 1024);
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes), kBoxBytes, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)   // 16 k-rows x 128 B = 2048 B per UMMA_K
            umma_bf16(tmem_base, adesc + (uint64_t)(k * 128), bdesc + (uint64_t)(k * 128), idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tmem_full[0]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp >= 4) {
      const int q = warp - 4;
      uint8_t* slab = slabs + q * kSlabBytes;
      float dummy_a[4] = {0, 0, 0, 0}, dummy_b[4] = {0, 0, 0, 0};
      mbar_wait(&tmem_full[0], 0);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int64_t row0 = (int64_t)m_blk * kBlockM + q * 32;
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        epilogue_chunk<false>(p, t_addr + c * 64, slab, lane, row0, (int64_t)n_blk * BLOCK_N + c * 64, dummy_a, dummy_b);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ================================================================================================ host side
// bf16 2-D tensor [outer, inner] with row pitch `ld` elements; box = {64, box_rows}; SWIZZLE_128B; OOB reads give zero
static int make_tmap(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_rows, const char* who) {
  PFN_encodeTiled fn = get_encode_fn();
  UNET_REQUIRE(fn, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  UNET_REQUIRE(aligned16(base) && (ld % 8 == 0), UNET_EALIGN, "%s: TMA operand needs a 16B-aligned base and ld%%8==0", who);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return UNET_OK;
}

static void fill_params(TcParams& p, const unet_gemm_args* a) {
  p.M = a->M; p.N = a->N; p.K = a->K; p.C = a->C; p.ldc = a->ldc;
  p.epilogue = a->epilogue; p.out_bf16 = a->out_dtype == UNET_BF16; p.accumulate = a->accumulate;
  p.scale = a->scale; p.shift = a->shift; p.colsum = a->colsum; p.colsq = a->colsq;
  p.convt_H = a->convt_H; p.convt_W = a->convt_W; p.convt_cout = a->N / 4;
  p.drop_on = 0; p.keep = 1.f; p.inv_keep = 1.f; p.seed = 0; p.ctot = 0; p.c0 = 0; p.seed_dev = nullptr;
  if (a->epilogue == UNET_EPI_CONVT && a->drop.rate > 0.f) {
    p.drop_on = 1; p.keep = 1.f - a->drop.rate; p.inv_keep = 1.f / (1.f - a->drop.rate);
    p.seed = a->drop.seed; p.ctot = a->drop.ctot; p.c0 = a->drop.c0; p.seed_dev = a->drop.seed_dev;
  }
}

template <int BLOCK_N, bool OUT_BF16>
static int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_nt_kernel<BLOCK_N, OUT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error(e, "gemm_tc: cudaFuncSetAttribute");
    attr_done = true;
  }
  p.num_m_tiles = (int)ceil_div(p.M, kBlockM);
  p.num_n_tiles = (int)ceil_div(p.N, BLOCK_N);
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const unsigned grid = (unsigned)i64min(tiles, sm_count());
  gemm_tc_nt_kernel<BLOCK_N, OUT_BF16><<<grid, 256, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  UNET_LAUNCH_CHECK("gemm_tc_nt");
  return UNET_OK;
}

template <int BLOCK_N>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_wgrad_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_cuda_error(e, "gemm_tc: cudaFuncSetAttribute");
    attr_done = true;
  }
  p.num_m_tiles = (int)ceil_div(p.M, kBlockM);
  p.num_n_tiles = (int)ceil_div(p.N, BLOCK_N);
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const int64_t total_kb = ceil_div(p.K, kBlockK);
  int64_t splits = i64max(1, ((int64_t)sm_count() * 2) / tiles);
  splits = i64min(splits, i64max(1, total_kb / 8));
  splits = i64min(splits, 65535);
  p.kb_per_split = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, p.kb_per_split);
  dim3 grid((unsigned)tiles, (unsigned)splits);
  gemm_tc_wgrad_kernel<BLOCK_N><<<grid, 256, Cfg::kSmemBytes, st>>>(tmA, tmB, p);
  UNET_LAUNCH_CHECK("gemm_tc_wgrad");
  return UNET_OK;
}

static int pick_block_n(int64_t N) { return N >= 256 ? 256 : (N > 64 ? 128 : 64); }

}  // namespace unet

using namespace unet;

extern "C" int unet_gemm_tc(const unet_gemm_args* a, void* stream) {
  if (int e = gemm_validate(a, "gemm_tc")) return e;
  UNET_REQUIRE(a->in_dtype == UNET_BF16, UNET_EUNSUPPORTED, "gemm_tc: operands must be bf16 (fp32 mode uses gemm_simt)");
  UNET_REQUIRE(a->N % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: N must be a multiple of 8 (got %lld)", (long long)a->N);
  UNET_REQUIRE(a->ldc % 4 == 0 && aligned16(a->C), UNET_EALIGN, "gemm_tc: C needs a 16B-aligned base and ldc%%4==0");
  UNET_REQUIRE(!a->scale || aligned16(a->scale), UNET_EALIGN, "gemm_tc: scale must be 16B aligned");
  UNET_REQUIRE(!a->shift || aligned16(a->shift), UNET_EALIGN, "gemm_tc: shift must be 16B aligned");
  if (a->epilogue == UNET_EPI_CONVT)
    UNET_REQUIRE(a->M < (int64_t)1 << 31, UNET_EUNSUPPORTED, "gemm_tc: CONVT needs M < 2^31");
  if (a->epilogue == UNET_EPI_CONVT)
    UNET_REQUIRE((a->N / 4) % 64 == 0, UNET_EUNSUPPORTED, "gemm_tc: CONVT needs Cout%%64==0 (got %lld)", (long long)(a->N / 4));
  cudaStream_t st = (cudaStream_t)stream;
  TcParams p{};
  fill_params(p, a);
  CUtensorMap tmA, tmB;
  const int bn = pick_block_n(a->N);

  if (!a->a_trans) {
    UNET_REQUIRE(a->b_trans == 1, UNET_EUNSUPPORTED, "gemm_tc: forward/dgrad needs B given as [N,K] (b_trans=1)");
    UNET_REQUIRE(!a->accumulate, UNET_EUNSUPPORTED, "gemm_tc: accumulate is only implemented for a_trans=1");
    UNET_REQUIRE(a->K % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: K must be a multiple of 8 (got %lld)", (long long)a->K);
    if (int e = make_tmap(&tmA, a->A, a->K, a->M, a->lda, kBlockM, "gemm_tc(A)")) return e;
    if (int e = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, bn, "gemm_tc(B)")) return e;
    const bool ob = a->out_dtype == UNET_BF16;
    if (bn == 256) return ob ? launch_nt<256, true>(tmA, tmB, p, st) : launch_nt<256, false>(tmA, tmB, p, st);
    if (bn == 128) return ob ? launch_nt<128, true>(tmA, tmB, p, st) : launch_nt<128, false>(tmA, tmB, p, st);
    return ob ? launch_nt<64, true>(tmA, tmB, p, st) : launch_nt<64, false>(tmA, tmB, p, st);
  }
  // weight gradient: C[M,N] += A[K,M]^T * B[K,N]
  UNET_REQUIRE(a->b_trans == 0 && a->accumulate == 1, UNET_EUNSUPPORTED, "gemm_tc: a_trans=1 needs b_trans=0 and accumulate=1");
  UNET_REQUIRE(a->M % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: wgrad M must be a multiple of 8");
  if (int e = make_tmap(&tmA, a->A, a->M, a->K, a->lda, 64, "gemm_tc(wgrad A)")) return e;
  if (int e = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, 64, "gemm_tc(wgrad B)")) return e;
  if (bn == 256) return launch_wgrad<256>(tmA, tmB, p, st);
  if (bn == 128) return launch_wgrad<128>(tmA, tmB, p, st);
  return launch_wgrad<64>(tmA, tmB, p, st);
}
