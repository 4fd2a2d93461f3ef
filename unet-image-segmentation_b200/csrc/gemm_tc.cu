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
// Warp roles are described at each kernel.  The weight-gradient kernel keeps a slab epilogue (TMEM -> padded fp32 slab ->
// lane owns 4 fixed columns -> fp32 atomics); the NT kernel stages bf16 rows for TMA stores.
#include "common.cuh"
#include "ptx.cuh"

namespace unet {

int gemm_validate(const unet_gemm_args* a, const char* who);

// ================================================================================================ descriptors
// Shared-memory matrix descriptor (SM100 format, version 1), SWIZZLE_128B.
//   K-major  tile (rows x 64 bf16, 128 B per row): SBO = 1024 B (8 rows), LBO unused.
//   MN-major tile (64 k-rows x 64 bf16 per 8 KB chunk): SBO = 1024 B (8 k-rows), LBO = chunk stride.
//   MN-major tf32 tile (fp32 mode): the only layout the hardware takes is SWIZZLE_128B_BASE32B (layout type 1; TMA:
//   CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 128 B (32 fp32 along MN) x 4 k-rows, SBO = 512 B, LBO = chunk stride.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;          // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// kind::tf32: D fp32, A/B fp32 words read as tf32 (8 k per instruction = the same 32 bytes per row as 16 bf16), M = 128, K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
  const float* head_w; const float* head_b; float* head_out; int head_classes;   // UNET_EPI_HEAD
  int split_kb;             // nt: k-blocks [split_kb, ...) come from the second A map;  wgrad: column chunks (64) >= split_kb from the second B map
};

constexpr int kBlockM = 128, kBlockK = 64;
constexpr int kSlabPitch = 64 * 4 + 16;        // bytes per staged row: 64 fp32 + 16 B pad (17 x 16 B: conflict-free)
constexpr int kSlabBytes = 32 * kSlabPitch;    // per epilogue warp
constexpr int kNumEpiWarps = 4;

// X3 (fp32 mode): fp32 operands with their tf32 `lo` parts (see NtCfg); 32 k-rows per stage, 32-column (128-byte) chunks.
template <int BLOCK_N, bool X3 = false> struct TcCfg {
  static constexpr int kKRows = X3 ? 32 : 64;                  // k-rows (pixels) per stage
  static constexpr int kChunkCols = X3 ? 32 : 64;              // elements per 128-byte chunk row
  static constexpr int kBox = kKRows * 128;                    // one chunk: kKRows rows x 128 B
  static constexpr int kAChunks = kBlockM / kChunkCols, kBChunks = BLOCK_N / kChunkCols;
  static constexpr int kABytes = kAChunks * kBox * (X3 ? 2 : 1);   // X3: [hi chunks | lo chunks]
  static constexpr int kBBytes = kBChunks * kBox * (X3 ? 2 : 1);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = X3 ? (BLOCK_N == 64 ? 3 : 2) : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 5 : 7));
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kNumEpiWarps * kSlabBytes + 256;
};

// ------------------------------------------------------------------------------------------------ weight-gradient epilogue
// One 64-column chunk of one warp's 32 accumulator rows: TMEM -> slab -> (lane owns 4 fixed columns) -> fp32 atomics into C.
// (The weight-gradient kernel is the only user: accumulate == 1, no fused epilogue — validated on the host.)
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
    const uint32_t dst = smem_u32(slab) + (uint32_t)(lane * kSlabPitch + half * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) sts128(dst + i * 16, make_uint4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]));
  }
  __syncwarp();

  // 2) lane owns columns [n_base + 4*seg, +4) of rows (it*2 + lane/16)
  const int seg = lane & 15, rsub = lane >> 4;
  const int64_t n = n_base + seg * 4;
  const bool col_ok = n < p.N;     // N is a multiple of 8: a 4-column group is all-or-nothing
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int r = it * 2 + rsub;
    const int64_t m = row0 + r;
    const float4 a = lds128f(smem_u32(slab) + (uint32_t)(r * kSlabPitch + seg * 16));
    if (!col_ok || m >= p.M) continue;
    float* dst = reinterpret_cast<float*>(p.C) + m * p.ldc + n;
    atomicAdd(dst, a.x); atomicAdd(dst + 1, a.y); atomicAdd(dst + 2, a.z); atomicAdd(dst + 3, a.w);
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
// Warp roles (384 threads, 1 CTA/SM, persistent over output tiles):
//   warp 0      TMA producer (one thread)   smem ring of NtCfg::kStages {A 128x64, B BLOCK_Nx64} bf16 tiles, SWIZZLE_128B
//   warp 1      MMA issuer   (one thread)   tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16, fp32 accumulators in TMEM
//   warp 2      TMEM allocator              2 accumulator stages x BLOCK_N columns
//   warps 4-7   epilogue group 0            tiles 0, 2, 4, ... of this CTA (accumulator stage 0)
//   warps 8-11  epilogue group 1            tiles 1, 3, 5, ... (accumulator stage 1)
// Epilogue (per group, 128 threads, thread = accumulator row): tcgen05.ld 32 columns -> scale/shift/ReLU/bias/dropout in
// registers -> bf16 (or fp32) rows into a 128B-swizzled staging tile (conflict-free 16 B shared stores) -> ONE TMA
// store per 128-byte-wide column chunk (2-D box for row-major C, 5-D box for the Conv2DTranspose pixel shuffle, so the
// scatter into the concat buffer costs no address arithmetic) -> BN batch statistics re-read column-wise from the
// staged tile (one shared load per row per lane, fp32 sums kept in registers across tiles).  Staging is double
// buffered per group, so the store of chunk c overlaps the math of chunk c+1 and the other group's tile.
constexpr int kNumEpiGroups = 2;
constexpr int kStageTileBytes = 128 * 128;       // 128 rows x 128 B

// X3 (fp32 mode): operands are fp32, each given as a (hi, lo) pair of tensors: hi = the tensor itself (kind::tf32 ignores the low
// 13 mantissa bits, i.e. multiplies trunc(x)), lo = tf32(x - trunc(x)) from unet_split_tf32;
// a stage holds [A_hi | A_lo] and [B_hi | B_lo] tiles of 32 k (128 B rows) and every k-step issues three kind::tf32 MMAs
// (hi*hi + lo*hi + hi*lo): fp32-grade products (error ~2^-21 of |a||b|) at a third of the tf32 rate.
template <int BLOCK_N, bool X3 = false> struct NtCfg {
  static constexpr int kTileA = kBlockM * 128;          // 128 rows x 128 B: 64 bf16 or 32 fp32 per row
  static constexpr int kTileB = BLOCK_N * 128;
  static constexpr int kABytes = kTileA * (X3 ? 2 : 1);
  static constexpr int kBBytes = kTileB * (X3 ? 2 : 1);
  static constexpr int kKElems = X3 ? 32 : 64;          // k per stage
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = X3 ? (BLOCK_N == 64 ? 3 : 2) : (BLOCK_N == 256 ? 3 : (BLOCK_N == 128 ? 4 : 5));
  static constexpr int kTmemCols = 2 * BLOCK_N;
  static constexpr int kHeadFloats = BLOCK_N == 64 ? 8 * 64 + 8 : 0;        // fused output head: w[class][64] + bias[8]
  static constexpr int kParBytes = kNumEpiGroups * (2 * BLOCK_N + kHeadFloats) * 4;   // per group: scale, shift[, head]
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kNumEpiGroups * 2 * kStageTileBytes + kParBytes + 256;
};

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int BLOCK_N, bool OUT_BF16, bool X3 = false>
__global__ void __launch_bounds__(384, 1)
gemm_tc_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2,
                  const __grid_constant__ CUtensorMap tmC, const TcParams p) {
  static_assert(!X3 || (!OUT_BF16 && BLOCK_N <= 128), "the tf32x3 variant serves fp32 mode: fp32 output, tiles up to 128 x 128");
  pdl_launch_dependents();
  using Cfg = NtCfg<BLOCK_N, X3>;
  constexpr int kStages = Cfg::kStages;
  constexpr int CW = OUT_BF16 ? 64 : 32;          // output columns per 128-byte staging row
  constexpr int kChunks = BLOCK_N / CW;
  constexpr int CPL = CW / 32;                    // statistic columns owned by a lane per chunk
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* stage_tiles = smem + kStages * Cfg::kStageBytes;                                // [group][2][16 KB]
  float* s_par = reinterpret_cast<float*>(stage_tiles + kNumEpiGroups * 2 * kStageTileBytes);  // [group][2][BLOCK_N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_par) + Cfg::kParBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = (int)((p.K + Cfg::kKElems - 1) / Cfg::kKElems);
  const int total_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmB2); tma_prefetch_desc(&tmC); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          if (X3) {          // (hi, lo) pairs of both operands: tmA2 / tmB2 are the lo tensors
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * Cfg::kKElems, m_blk * kBlockM, kEvictFirst);
            tma_load_2d(smem_a + stage * Cfg::kABytes + Cfg::kTileA, &tmA2, &full_bar[stage], kb * Cfg::kKElems, m_blk * kBlockM, kEvictFirst);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * Cfg::kKElems, n_blk * BLOCK_N, kEvictLast);
            tma_load_2d(smem_b + stage * Cfg::kBBytes + Cfg::kTileB, &tmB2, &full_bar[stage], kb * Cfg::kKElems, n_blk * BLOCK_N, kEvictLast);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
          if (kb < p.split_kb) tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA, &full_bar[stage], kb * kBlockK, m_blk * kBlockM, kEvictFirst);
          else tma_load_2d(smem_a + stage * Cfg::kABytes, &tmA2, &full_bar[stage], (kb - p.split_kb) * kBlockK, m_blk * kBlockM, kEvictFirst);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmB, &full_bar[stage], kb * kBlockK, n_blk * BLOCK_N, kEvictLast);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = X3 ? make_idesc_tf32(BLOCK_N) : make_idesc(BLOCK_N, false, false);
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
          if (X3) {
            const uint64_t alo = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes + Cfg::kTileA), 0, 1024);
            const uint64_t blo = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes + Cfg::kTileB), 0, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {          // 8 k (32 B) per kind::tf32 instruction; small terms first
              umma_tf32(d_tmem, alo + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
              umma_tf32(d_tmem, adesc + 2 * k, blo + 2 * k, idesc, 1);
              umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)   // +32 B per UMMA_K inside the 128 B swizzle atom
              umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    const int g = (warp - 4) >> 2;                // epilogue group = accumulator stage
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                // accumulator row of this thread inside the tile
    const int gtid = (warp - 4 - 4 * g) * 32 + lane;
    const bool issuer = gtid == 0;
    uint8_t* tiles = stage_tiles + g * 2 * kStageTileBytes;
    float* par_scale = s_par + g * (2 * BLOCK_N + Cfg::kHeadFloats);
    float* par_shift = par_scale + BLOCK_N;
    float* par_head = par_shift + BLOCK_N;         // [8][64] class-major weights, then 8 biases (BLOCK_N == 64 only)
    const int epi = p.epilogue;
    constexpr bool kHeadCapable = BLOCK_N == 64 && OUT_BF16;
    const bool head = kHeadCapable && epi == UNET_EPI_HEAD;
    const bool affine = epi == UNET_EPI_AFFINE || epi == UNET_EPI_AFFINE_RELU || epi == UNET_EPI_CONVT || epi == UNET_EPI_HEAD;
    const bool relu = epi == UNET_EPI_AFFINE_RELU || epi == UNET_EPI_HEAD;
    const bool store_c = p.C != nullptr;
    if (kHeadCapable && head) {                    // visible to the group after the first named barrier of the chunk loop
      for (int i = gtid; i < 8 * 64; i += 128) {
        const int cls = i >> 6, k = i & 63;
        par_head[i] = (cls < p.head_classes && k < p.N) ? __ldg(p.head_w + (int64_t)k * p.head_classes + cls) : 0.f;
      }
      if (gtid < 8) par_head[8 * 64 + gtid] = (gtid < p.head_classes && p.head_b) ? __ldg(p.head_b + gtid) : 0.f;
    }
    const bool stats = epi == UNET_EPI_STATS;
    const uint32_t seed = p.drop_on ? p.seed + (p.seed_dev ? __ldg(p.seed_dev) : 0u) : 0u;
    const float drop_ik = p.drop_on ? p.inv_keep : 1.0f;
    // Conv2DTranspose: tile rows are input pixels (q = image*H + i, j); the box covers min(W,128) columns x 128/min(W,128) rows
    const int cw = p.convt_W;
    float st_sum[kChunks][CPL], st_sq[kChunks][CPL];
#pragma unroll
    for (int c = 0; c < kChunks; ++c)
#pragma unroll
      for (int j = 0; j < CPL; ++j) { st_sum[c][j] = 0.f; st_sq[c][j] = 0.f; }
    const uint32_t sw = (uint32_t)(row & 7);
    int last_n_blk = -1, chunk_ctr = 0;
    uint32_t stat_off[8];                          // statistics: byte offset of this lane's word inside a staged row r (r & 7 = i)
#pragma unroll
    for (int i = 0; i < 8; ++i) stat_off[i] = ((((uint32_t)lane >> 2) ^ (uint32_t)i) << 4) + (((uint32_t)lane & 3u) << 2);

    auto flush_stats = [&](int n_blk_) {
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int64_t n = (int64_t)n_blk_ * BLOCK_N + c * CW + lane * CPL + j;
          if (n < p.N) { atomicAdd(p.colsum + n, (double)st_sum[c][j]); atomicAdd(p.colsq + n, (double)st_sq[c][j]); }
          st_sum[c][j] = 0.f; st_sq[c][j] = 0.f;
        }
    };

    int it = g;
    for (int tile = blockIdx.x + g * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, it += 2) {
      const int m_blk = tile / p.num_n_tiles, n_blk = tile % p.num_n_tiles;
      if (n_blk != last_n_blk) {
        if (stats && last_n_blk >= 0) flush_stats(last_n_blk);
        if (affine) {
          named_bar_sync(1 + g, 128);             // nobody still reads the previous tile's parameters
          for (int i = gtid; i < BLOCK_N; i += 128) {
            const int64_t n = (int64_t)n_blk * BLOCK_N + i;
            float sc = 1.f, sh = 0.f;
            if (n < p.N) {
              if (epi == UNET_EPI_CONVT) { if (p.shift) sh = __ldg(p.shift + (n % p.convt_cout)); if (p.drop_on) sh *= p.inv_keep; }
              else { if (p.scale) sc = __ldg(p.scale + n); if (p.shift) sh = __ldg(p.shift + n); }
            }
            par_scale[i] = sc; par_shift[i] = sh;
          }
          // visibility is provided by the first named barrier of the chunk loop below
        }
        last_n_blk = n_blk;
      }
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[g], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N;
      const int m0 = m_blk * kBlockM;
      float hacc[kHeadCapable ? 8 : 1];
#pragma unroll
      for (int i = 0; i < (kHeadCapable ? 8 : 1); ++i) hacc[i] = 0.f;
#pragma unroll
      for (int c = 0; c < kChunks; ++c, ++chunk_ctr) {
        uint8_t* buf = tiles + (chunk_ctr & 1) * kStageTileBytes;
        const int n_base = n_blk * BLOCK_N + c * CW;
        if (n_base >= p.N) break;                 // uniform: ragged last N tile
        if (issuer) bulk_wait_read<1>();          // the store that last used `buf` (two chunks ago) has read it
        named_bar_sync(1 + g, 128);
        const uint32_t my_row_s = smem_u32(buf) + (uint32_t)row * 128u;
        // Conv2DTranspose destination of this thread's row (dropout indexes the destination element)
        uint64_t drop_base = 0;
        int cv_a = 0, cv_b = 0, co0 = 0;
        if (epi == UNET_EPI_CONVT) {
          const uint32_t cc = (uint32_t)p.convt_cout;
          const uint32_t ab = (uint32_t)n_base / cc;
          co0 = (int)((uint32_t)n_base - ab * cc); cv_a = (int)(ab >> 1); cv_b = (int)(ab & 1);
          if (p.drop_on) {
            const uint32_t m32 = (uint32_t)(m0 + row), qq = m32 / (uint32_t)cw, jj = m32 - qq * (uint32_t)cw;
            const uint64_t pix = ((uint64_t)(2 * qq + cv_a)) * (uint64_t)(2 * cw) + 2 * jj + cv_b;
            drop_base = pix * (uint64_t)p.ctot + (uint64_t)p.c0 + (uint64_t)co0;
          }
        }
#pragma unroll
        for (int half = 0; half < CW / 32; ++half) {
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c * CW + half * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (affine) {
            const uint32_t ps_s = smem_u32(par_scale + c * CW + half * 32);
            const uint32_t ph_s = smem_u32(par_shift + c * CW + half * 32);
            if (p.scale && epi != UNET_EPI_CONVT) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {   // warp-uniform addresses: broadcast shared loads
                const float4 s4 = lds128f(ps_s + i * 4);
                const float4 h4 = lds128f(ph_s + i * 4);
                v[i] = fmaf(v[i], s4.x, h4.x); v[i + 1] = fmaf(v[i + 1], s4.y, h4.y);
                v[i + 2] = fmaf(v[i + 2], s4.z, h4.z); v[i + 3] = fmaf(v[i + 3], s4.w, h4.w);
              }
            } else {                              // bias only (Conv2DTranspose, folded BN scale, folded BN backward); with Dropout
              const float ik = drop_ik;           // the kept values' 1/keep rides in the same FMA (the staged bias is pre-divided)
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                const float4 h4 = lds128f(ph_s + i * 4);
                v[i] = fmaf(v[i], ik, h4.x); v[i + 1] = fmaf(v[i + 1], ik, h4.y);
                v[i + 2] = fmaf(v[i + 2], ik, h4.z); v[i + 3] = fmaf(v[i + 3], ik, h4.w);
              }
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (p.drop_on) dropout_apply<32, true, true>(v, drop_base + half * 32, seed, p.keep, p.inv_keep);
          }
          if (kHeadCapable && head) {             // 1x1 output convolution from the activations as they will be stored (bf16)
            const int ncls = p.head_classes;
#pragma unroll
            for (int cls = 0; cls < 8; ++cls) {
              if (cls < ncls) {
                const uint32_t hw_s = smem_u32(par_head + cls * 64 + half * 32);
                float2 a2 = make_float2(hacc[cls], 0.f);     // fp32 activations (not re-rounded to bf16), packed FMAs
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 w4 = lds128f(hw_s + i * 4);
                  a2 = fma2(make_float2(v[i], v[i + 1]), make_float2(w4.x, w4.y), a2);
                  a2 = fma2(make_float2(v[i + 2], v[i + 3]), make_float2(w4.z, w4.w), a2);
                }
                hacc[cls] = a2.x + a2.y;
              }
            }
          }
          if (!store_c) continue;
          if (OUT_BF16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {         // 4 x 16 B = 32 bf16 columns; chunk index = half*4 + j
              const uint4 o = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                         pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
              sts128(my_row_s + ((((uint32_t)(half * 4 + j)) ^ sw) << 4), o);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {         // 8 x 16 B = 32 fp32 columns
              const uint4 o = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
              sts128(my_row_s + ((((uint32_t)j) ^ sw) << 4), o);
            }
          }
        }
        if (c == kChunks - 1 || n_base + CW >= p.N) {   // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[g]);
        }
        fence_proxy_async();
        named_bar_sync(1 + g, 128);
        if (issuer && store_c) {
          if (epi == UNET_EPI_CONVT) {
            const int q0 = m0 / cw, j0 = m0 - q0 * cw;
            tma_store_5d(&tmC, buf, co0, j0, q0, cv_b, cv_a);
          } else {
            tma_store_2d(&tmC, buf, n_base, m0);
          }
          bulk_commit();
        }
        if (stats) {
          // column sums over this warp's 32 rows, from the values as stored; lane owns CPL adjacent columns.  Rows past M
          // are exact zeros (TMA zero-fills the A operand), so no bounds check; shared addresses carry the 128B-swizzle XOR
          // of the row in 8 precomputed lane offsets, everything else is an immediate.
          const uint32_t wbase = smem_u32(buf) + (uint32_t)(q * 32) * 128u;
          float2 s01 = make_float2(0.f, 0.f), q01 = make_float2(0.f, 0.f);
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t w = lds32(wbase + (uint32_t)rr * 128u + stat_off[rr & 7]);
            if (OUT_BF16) {
              const float2 ab = make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
              s01.x += ab.x; s01.y += ab.y;
              q01 = fma2(ab, ab, q01);
            } else {
              const float a = __uint_as_float(w);
              s01.x += a; q01.x = fmaf(a, a, q01.x);
            }
          }
          st_sum[c][0] += s01.x; st_sq[c][0] += q01.x;
          if (CPL == 2) { st_sum[c][CPL - 1] += s01.y; st_sq[c][CPL - 1] += q01.y; }
        }
      }
      if (kHeadCapable && head && (int64_t)m0 + row < p.M) {
        const int ncls = p.head_classes;
        float* dst = p.head_out + ((int64_t)m0 + row) * ncls;
        if (ncls == 1) {
          dst[0] = 1.f / (1.f + expf(-(hacc[0] + par_head[8 * 64])));
        } else {
          float mx = -INFINITY, e[8], den = 0.f;
#pragma unroll
          for (int cls = 0; cls < 8; ++cls) if (cls < ncls) { hacc[cls] += par_head[8 * 64 + cls]; mx = fmaxf(mx, hacc[cls]); }
#pragma unroll
          for (int cls = 0; cls < 8; ++cls) if (cls < ncls) { e[cls] = expf(hacc[cls] - mx); den += e[cls]; }
          const float inv = 1.f / den;
#pragma unroll
          for (int cls = 0; cls < 8; ++cls) if (cls < ncls) dst[cls] = e[cls] * inv;
        }
      }
    }
    if (stats && last_n_blk >= 0) flush_stats(last_n_blk);
    if (issuer) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------ weight-gradient kernel
// One CTA = one (128 x BLOCK_N) tile of C and one slice of the reduction dimension P; partial products are added to C
// with fp32 atomics.  Operands are MN-major: each 64-row (P) x 64-column box lands as an 8 KB SWIZZLE_128B chunk.
template <int BLOCK_N, bool X3 = false>
__global__ void __launch_bounds__(256, 1)
gemm_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2, const TcParams p) {
  static_assert(!X3 || BLOCK_N <= 128, "the tf32x3 variant uses tiles up to 128 x 128");
  pdl_launch_dependents();
  using Cfg = TcCfg<BLOCK_N, X3>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = BLOCK_N / 64;             // 64-column chunks of the accumulator (epilogue) and of the bf16 B operand
  constexpr int kBoxBytes = Cfg::kBox;
  constexpr int kKRows = Cfg::kKRows;
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
  const int64_t total_kb = (p.K + kKRows - 1) / kKRows;
  const int64_t kb_begin = (int64_t)blockIdx.y * p.kb_per_split;
  const int64_t kb_end = min(total_kb, kb_begin + p.kb_per_split);
  const int num_k = (int)i64max(0, kb_end - kb_begin);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmB2); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&tmem_full[0], 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t tmem_base = *tmem_ptr;

  if (num_k > 0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int prow = (int)((kb_begin + kb) * kKRows);
          if (X3) {          // fp32 operands: 32-column chunks; tmA2 / tmB2 are the lo tensors
#pragma unroll
            for (int c = 0; c < Cfg::kAChunks; ++c) {
              tma_load_2d(smem_a + stage * Cfg::kABytes + c * kBoxBytes, &tmA, &full_bar[stage], m_blk * kBlockM + c * 32, prow, kEvictFirst);
              tma_load_2d(smem_a + stage * Cfg::kABytes + (Cfg::kAChunks + c) * kBoxBytes, &tmA2, &full_bar[stage], m_blk * kBlockM + c * 32, prow, kEvictFirst);
            }
#pragma unroll
            for (int c = 0; c < Cfg::kBChunks; ++c) {
              tma_load_2d(smem_b + stage * Cfg::kBBytes + c * kBoxBytes, &tmB, &full_bar[stage], n_blk * BLOCK_N + c * 32, prow, kEvictFirst);
              tma_load_2d(smem_b + stage * Cfg::kBBytes + (Cfg::kBChunks + c) * kBoxBytes, &tmB2, &full_bar[stage], n_blk * BLOCK_N + c * 32, prow, kEvictFirst);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            continue;
          }
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_2d(smem_a + stage * Cfg::kABytes + c * kBoxBytes, &tmA, &full_bar[stage], m_blk * kBlockM + c * 64, prow, kEvictFirst);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            const int nc = n_blk * (BLOCK_N / 64) + c;           // 64-column chunk of the (possibly concatenated) B operand
            if (nc < p.split_kb) tma_load_2d(smem_b + stage * Cfg::kBBytes + c * kBoxBytes, &tmB, &full_bar[stage], nc * 64, prow, kEvictFirst);
            else tma_load_2d(smem_b + stage * Cfg::kBBytes + c * kBoxBytes, &tmB2, &full_bar[stage], (nc - p.split_kb) * 64, prow, kEvictFirst);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = X3 ? (make_idesc_tf32(BLOCK_N) | (1u << 15) | (1u << 16)) : make_idesc(BLOCK_N, true, true);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          constexpr uint32_t kSbo = X3 ? 512u : 1024u, kLay = X3 ? 1u : 2u;
          const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes), kBoxBytes, kSbo, kLay);
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes), kBoxBytes, kSbo, kLay);
          if (X3) {
            const uint64_t alo = make_smem_desc(smem_u32(smem_a + stage * Cfg::kABytes + Cfg::kAChunks * kBoxBytes), kBoxBytes, kSbo, kLay);
            const uint64_t blo = make_smem_desc(smem_u32(smem_b + stage * Cfg::kBBytes + Cfg::kBChunks * kBoxBytes), kBoxBytes, kSbo, kLay);
#pragma unroll
            for (int k = 0; k < kKRows / 8; ++k) {   // 8 k-rows x 128 B = 1024 B per kind::tf32 instruction
              const uint64_t o = (uint64_t)(k * 64);
              umma_tf32(tmem_base, alo + o, bdesc + o, idesc, (kb | k) != 0);
              umma_tf32(tmem_base, adesc + o, blo + o, idesc, 1);
              umma_tf32(tmem_base, adesc + o, bdesc + o, idesc, 1);
            }
          } else
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


// ------------------------------------------------------------------------------------------------ fused pointwise backward
// Both contractions of the folded pointwise backward (engine.py, SeparableConv2D pointwise half + BatchNormalization backward)
// from ONE pass over [g | z] and d:
//   dd[P,Cin]     = [g | z] * wab^T + bias       (data gradient;   unet_gemm_tc a_trans=0 with A2)
//   G [Cin,2C]   += d^T [g | z]                  (weight gradient; unet_gemm_tc a_trans=1 with B2)
// A 128-pixel x 64-channel bf16 box landed by TMA with SWIZZLE_128B is at once a K-major operand (rows = pixels, K = channels:
// the data gradient's A) and an MN-major operand (K = pixels, MN = channels: both operands of the weight gradient), so each
// stage {g, z, d chunks} feeds both MMAs and [g | z] is read from HBM once instead of twice.  C = 64 (2C = 128 = the MMA's M
// for the weight gradient, computed transposed: G^T[2C,Cin] = [g | z]^T d), Cin in {64, 128}.
//   warp 0   TMA producer (wab once, then one stage per 128 pixels)      warp 1   MMA issuer (8 + 8 tcgen05.mma per stage)
//   warp 2   TMEM allocator                                              warps 4-7  epilogue: dd tile per stage (TMEM -> +bias ->
//   bf16 -> swizzled staging -> TMA store), and once at the end G^T (TMEM -> fp32 atomics, coalesced along 2C)
template <int CIN> struct PbCfg {
  static constexpr int kPix = 128;
  static constexpr int kChunk = kPix * 128;                        // 128 pixel rows x 64 bf16
  static constexpr int kDChunks = CIN / 64;
  static constexpr int kStageBytes = (2 + kDChunks) * kChunk;      // g | z | d
  static constexpr int kStages = CIN == 64 ? 3 : 2;
  static constexpr int kWBytes = 2 * CIN * 128;                    // wab: two K-major chunks of [CIN rows x 64 k]
  static constexpr int kOutBytes = kDChunks * kChunk;              // dd staging
  static constexpr int kTmemCols = CIN == 64 ? 256 : 512;          // G^T: CIN columns, dd: 2 x CIN columns
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kWBytes + kOutBytes + CIN * 4 + 256;
};

struct PbParams { int64_t P; const float* bias; float* G; int64_t ldG; int num_blocks; };

template <int CIN>
__global__ void __launch_bounds__(256, 1)
pw_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmZ,
                    const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmO, const PbParams p) {
  pdl_launch_dependents();
  using Cfg = PbCfg<CIN>;
  constexpr int S = Cfg::kStages, kChunk = Cfg::kChunk, DC = Cfg::kDChunks;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  uint8_t* smem_w = stages + S * Cfg::kStageBytes;
  uint8_t* out_tile = smem_w + Cfg::kWBytes;
  float* s_bias = reinterpret_cast<float*>(out_tile + Cfg::kOutBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + CIN);
  uint64_t* full_bar = bars;             // [S]
  uint64_t* empty_bar = bars + S;        // [S]
  uint64_t* w_full = bars + 2 * S;
  uint64_t* d2_full = bars + 2 * S + 1;  // [2]
  uint64_t* d2_empty = bars + 2 * S + 3; // [2]
  uint64_t* d1_full = bars + 2 * S + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmZ); tma_prefetch_desc(&tmD); tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(w_full, 1); mbar_init(d1_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, Cfg::kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();            // everything above touched only shared memory / TMEM / kernel parameters
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t d1_tmem = tmem_base;                    // G^T accumulator [128 x CIN]
  const uint32_t d2_tmem = tmem_base + CIN;              // dd accumulators [2][128 x CIN]

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_full, Cfg::kWBytes);
      for (int c = 0; c < 2; ++c) tma_load_2d(smem_w + c * (CIN * 128), &tmW, w_full, c * 64, 0, kEvictLast);
      int s = 0; uint32_t ph = 0;
      for (int blk = blockIdx.x; blk < p.num_blocks; blk += gridDim.x) {
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
        uint8_t* st = stages + s * Cfg::kStageBytes;
        const int row = blk * Cfg::kPix;
        tma_load_2d(st, &tmG, &full_bar[s], 0, row, kEvictFirst);
        tma_load_2d(st + kChunk, &tmZ, &full_bar[s], 0, row, kEvictFirst);
        for (int c = 0; c < DC; ++c) tma_load_2d(st + (2 + c) * kChunk, &tmD, &full_bar[s], c * 64, row, kEvictFirst);
        if (++s == S) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_nt = make_idesc(CIN, false, false);    // dd:  A K-major (pixels x channels), B K-major (wab)
      constexpr uint32_t idesc_tn = make_idesc(CIN, true, true);      // G^T: A MN-major ([g|z]), B MN-major (d)
      mbar_wait(w_full, 0);
      int s = 0; uint32_t ph = 0; int it = 0;
      for (int blk = blockIdx.x; blk < p.num_blocks; blk += gridDim.x, ++it) {
        const int buf = it & 1; const uint32_t bph = (it >> 1) & 1;
        mbar_wait(&full_bar[s], ph);
        mbar_wait(&d2_empty[buf], bph ^ 1);
        tc_fence_after();
        const uint32_t st = smem_u32(stages + s * Cfg::kStageBytes);
        const uint32_t dd_t = d2_tmem + buf * CIN;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                  // K = [g channels | z channels]
          const uint64_t adesc = make_smem_desc(st + c * kChunk, 0, 1024);
          const uint64_t bdesc = make_smem_desc(smem_u32(smem_w) + c * (CIN * 128), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(dd_t, adesc + 2 * k, bdesc + 2 * k, idesc_nt, (c | k) != 0);
        }
        umma_commit(&d2_full[buf]);
        const uint64_t gz = make_smem_desc(st, kChunk, 1024);                  // M = 2C = 128: chunks g, z
        const uint64_t dm = make_smem_desc(st + 2 * kChunk, kChunk, 1024);     // N = CIN: chunks of d
#pragma unroll
        for (int k = 0; k < Cfg::kPix / 16; ++k)       // 16 pixel rows x 128 B = 2048 B per UMMA_K
          umma_bf16(d1_tmem, gz + (uint64_t)(k * 128), dm + (uint64_t)(k * 128), idesc_tn, (it | k) != 0);
        umma_commit(&empty_bar[s]);
        if (++s == S) { s = 0; ph ^= 1; }
      }
      umma_commit(d1_full);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;                            // TMEM lane quarter
    const int row = q * 32 + lane;
    const int gtid = (warp - 4) * 32 + lane;
    const bool issuer = gtid == 0;
    for (int i = gtid; i < CIN; i += 128) s_bias[i] = p.bias ? __ldg(p.bias + i) : 0.f;
    named_bar_sync(1, 128);
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t my_row_s = smem_u32(out_tile) + (uint32_t)row * 128u;
    int it = 0;
    for (int blk = blockIdx.x; blk < p.num_blocks; blk += gridDim.x, ++it) {
      const int buf = it & 1; const uint32_t bph = (it >> 1) & 1;
      mbar_wait(&d2_full[buf], bph);
      tc_fence_after();
      if (issuer) bulk_wait_read<0>();                 // the previous store has read the staging tile
      named_bar_sync(1, 128);
      const uint32_t t_addr = d2_tmem + buf * CIN + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < DC; ++c) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c * 64 + half * 32, r);
          tmem_ld_wait();
          float v[32];
          const uint32_t pb_s = smem_u32(s_bias + c * 64 + half * 32);
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 h4 = lds128f(pb_s + i * 4);
            v[i] = __uint_as_float(r[i]) + h4.x; v[i + 1] = __uint_as_float(r[i + 1]) + h4.y;
            v[i + 2] = __uint_as_float(r[i + 2]) + h4.z; v[i + 3] = __uint_as_float(r[i + 3]) + h4.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 o = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            sts128(my_row_s + c * kChunk + ((((uint32_t)(half * 4 + j)) ^ sw) << 4), o);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&d2_empty[buf]);
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (issuer) {
        for (int c = 0; c < DC; ++c) tma_store_2d(&tmO, out_tile + c * kChunk, c * 64, blk * Cfg::kPix);
        bulk_commit();
      }
    }
    // G^T[r, ci] (r = column of [g | z], ci = input channel) -> G[ci, r]: a warp's 32 lanes hit 32 consecutive floats
    mbar_wait(d1_full, 0);
    tc_fence_after();
    if (blockIdx.x < p.num_blocks) {
#pragma unroll
      for (int half = 0; half < CIN / 32; ++half) {
        uint32_t r[32];
        tmem_ld_32x32(d1_tmem + ((uint32_t)(q * 32) << 16) + half * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(p.G + (int64_t)(half * 32 + i) * p.ldG + row, __uint_as_float(r[i]));
      }
    }
    if (issuer) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ================================================================================================ host side
// bf16 2-D tensor [outer, inner] with row pitch `ld` elements; box = {64, box_rows}; SWIZZLE_128B; OOB reads give zero
// (f32: fp32 elements, box = {32, box_rows} — the same 128-byte rows)
static int make_tmap(CUtensorMap* map, const void* base, int64_t inner, int64_t outer, int64_t ld, int box_rows, const char* who,
                     bool f32 = false, bool atom32 = false) {
  PFN_encodeTiled fn = get_encode_fn();
  UNET_REQUIRE(fn, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const int es = f32 ? 4 : 2;
  UNET_REQUIRE(aligned16(base) && ((ld * es) % 16 == 0), UNET_EALIGN, "%s: TMA operand needs a 16B-aligned base and a 16B-multiple row pitch", who);
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * es};
  cuuint32_t box[2] = {f32 ? 32u : 64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return UNET_OK;
}

static void fill_params(TcParams& p, const unet_gemm_args* a) {
  p.M = a->M; p.N = a->N; p.K = a->K; p.C = a->C; p.ldc = a->ldc;
  p.epilogue = a->epilogue; p.out_bf16 = a->out_dtype == UNET_BF16; p.accumulate = a->accumulate;
  p.scale = a->scale; p.shift = a->shift; p.colsum = a->colsum; p.colsq = a->colsq;
  p.convt_H = a->convt_H; p.convt_W = a->convt_W; p.convt_cout = a->N / 4;
  p.drop_on = 0; p.keep = 1.f; p.inv_keep = 1.f; p.seed = 0; p.ctot = 0; p.c0 = 0; p.seed_dev = nullptr;
  p.head_w = a->head_w; p.head_b = a->head_b; p.head_out = a->head_out; p.head_classes = a->head_classes;
  if (a->epilogue == UNET_EPI_CONVT && a->drop.rate > 0.f) {
    p.drop_on = 1; p.keep = 1.f - a->drop.rate; p.inv_keep = 1.f / (1.f - a->drop.rate);
    p.seed = a->drop.seed; p.ctot = a->drop.ctot; p.c0 = a->drop.c0; p.seed_dev = a->drop.seed_dev;
  }
}

// C as a TMA store target.  Row-major C[M,N] (ptr, ldc): 2-D {N, M}, box {128 B of columns, 128 rows}.
// Conv2DTranspose: destination [Nimg, 2H, 2W, ldc] seen as 5-D {co, j, q = image*H + i, b, a} so that one box of
// {128 B of channels, bj input columns, 128/bj input rows} lands on the pixels (2i+a, 2j+b) of the upsampled image.
static int make_c_tmap(CUtensorMap* map, const unet_gemm_args* a, const char* who) {
  PFN_encodeTiled fn = get_encode_fn();
  UNET_REQUIRE(fn, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  const bool bf = a->out_dtype == UNET_BF16;
  const cuuint64_t es = bf ? 2 : 4;
  const cuuint32_t cwid = bf ? 64u : 32u;
  UNET_REQUIRE((a->ldc * es) % 16 == 0, UNET_EALIGN, "%s: C row pitch must be a multiple of 16 bytes", who);
  const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r;
  if (a->epilogue == UNET_EPI_CONVT) {
    const cuuint64_t W = a->convt_W, H = a->convt_H, cout = a->N / 4, nimg = a->M / ((int64_t)a->convt_H * a->convt_W);
    const cuuint32_t bj = (cuuint32_t)(W < (cuuint64_t)kBlockM ? W : kBlockM), bq = kBlockM / bj;
    cuuint64_t dims[5] = {cout, W, nimg * H, 2, 2};
    cuuint64_t strides[4] = {2 * a->ldc * es, 4 * W * a->ldc * es, a->ldc * es, 2 * W * a->ldc * es};
    cuuint32_t box[5] = {cwid, bj, bq, 1u, 1u};
    cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    r = fn(map, dt, 5, a->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)a->N, (cuuint64_t)a->M};
    cuuint64_t strides[1] = {a->ldc * es};
    cuuint32_t box[2] = {cwid, (cuuint32_t)kBlockM};
    cuuint32_t estr[2] = {1u, 1u};
    r = fn(map, dt, 2, a->C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled(C) failed with CUresult %d", who, (int)r);
  return UNET_OK;
}

template <int BLOCK_N, bool OUT_BF16, bool X3 = false>
static int launch_nt(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmC, TcParams& p, cudaStream_t st,
                     const CUtensorMap* tmB2p = nullptr) {
  using Cfg = NtCfg<BLOCK_N, X3>;
  const CUtensorMap& tmB2 = tmB2p ? *tmB2p : tmB;
  static SmemAttrOnce once;
  if (cudaError_t e = ensure_dynamic_smem(once, gemm_tc_nt_kernel<BLOCK_N, OUT_BF16, X3>, Cfg::kSmemBytes))
    return set_cuda_error(e, "gemm_tc: cudaFuncSetAttribute");
  p.num_m_tiles = (int)ceil_div(p.M, kBlockM);
  p.num_n_tiles = (int)ceil_div(p.N, BLOCK_N);
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const unsigned grid = (unsigned)i64min(tiles, sm_count());
  launch_pdl(gemm_tc_nt_kernel<BLOCK_N, OUT_BF16, X3>, grid, 384, Cfg::kSmemBytes, st, tmA, tmA2, tmB, tmB2, tmC, p);
  UNET_LAUNCH_CHECK("gemm_tc_nt");
  return UNET_OK;
}

template <int BLOCK_N, bool X3 = false>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmB2, TcParams& p, cudaStream_t st,
                        const CUtensorMap* tmA2p = nullptr) {
  using Cfg = TcCfg<BLOCK_N, X3>;
  const CUtensorMap& tmA2 = tmA2p ? *tmA2p : tmA;
  static SmemAttrOnce once;
  if (cudaError_t e = ensure_dynamic_smem(once, gemm_tc_wgrad_kernel<BLOCK_N, X3>, Cfg::kSmemBytes))
    return set_cuda_error(e, "gemm_tc: cudaFuncSetAttribute");
  p.num_m_tiles = (int)ceil_div(p.M, kBlockM);
  p.num_n_tiles = (int)ceil_div(p.N, BLOCK_N);
  const int64_t tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
  const int64_t total_kb = ceil_div(p.K, Cfg::kKRows);
  int64_t splits = i64max(1, ((int64_t)sm_count() * 2) / tiles);
  splits = i64min(splits, i64max(1, total_kb / 8));
  splits = i64min(splits, 65535);
  p.kb_per_split = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, p.kb_per_split);
  dim3 grid((unsigned)tiles, (unsigned)splits);
  launch_pdl(gemm_tc_wgrad_kernel<BLOCK_N, X3>, grid, 256, Cfg::kSmemBytes, st, tmA, tmA2, tmB, tmB2, p);
  UNET_LAUNCH_CHECK("gemm_tc_wgrad");
  return UNET_OK;
}

static int pick_block_n(int64_t N) { return N >= 256 ? 256 : (N > 64 ? 128 : 64); }

template <int CIN>
static int launch_pw_bwd_fused(const CUtensorMap& tmG, const CUtensorMap& tmZ, const CUtensorMap& tmD, const CUtensorMap& tmW,
                               const CUtensorMap& tmO, PbParams& p, cudaStream_t st) {
  using Cfg = PbCfg<CIN>;
  static SmemAttrOnce once;
  if (cudaError_t e = ensure_dynamic_smem(once, pw_bwd_fused_kernel<CIN>, Cfg::kSmemBytes))
    return set_cuda_error(e, "pw_bwd_fused: cudaFuncSetAttribute");
  const unsigned grid = (unsigned)i64min(p.num_blocks, sm_count());
  launch_pdl(pw_bwd_fused_kernel<CIN>, grid, 256, Cfg::kSmemBytes, st, tmG, tmZ, tmD, tmW, tmO, p);
  UNET_LAUNCH_CHECK("pw_bwd_fused");
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_pw_bwd_fused(const void* g, int64_t ldg, const void* z, int64_t ldz, const void* d, int64_t ldd,
                                 const void* wab, int64_t ldw, const float* bias, void* dd, int64_t lddd,
                                 float* G, int64_t ldG, int64_t P, int Cin, int C, void* stream) {
  UNET_REQUIRE(g && z && d && wab && dd && G, UNET_EINVAL, "pw_bwd_fused: null pointer");
  UNET_REQUIRE(P > 0 && P < ((int64_t)1 << 31) - 128, UNET_EINVAL, "pw_bwd_fused: bad P %lld", (long long)P);
  UNET_REQUIRE(C == 64 && (Cin == 64 || Cin == 128), UNET_EUNSUPPORTED,
               "pw_bwd_fused: needs C == 64 and Cin in {64, 128} (got C=%d Cin=%d); use unet_gemm_tc twice", C, Cin);
  UNET_REQUIRE(ldg >= C && ldz >= C && ldd >= Cin && ldw >= 2 * C && lddd >= Cin && ldG >= 2 * C, UNET_EINVAL, "pw_bwd_fused: ld too small");
  UNET_REQUIRE((lddd * 2) % 16 == 0 && aligned16(dd) && aligned16(G), UNET_EALIGN, "pw_bwd_fused: dd / G need 16B alignment");
  CUtensorMap tmG, tmZ, tmD, tmW, tmO;
  if (int e = make_tmap(&tmG, g, C, P, ldg, 128, "pw_bwd_fused(g)")) return e;
  if (int e = make_tmap(&tmZ, z, C, P, ldz, 128, "pw_bwd_fused(z)")) return e;
  if (int e = make_tmap(&tmD, d, Cin, P, ldd, 128, "pw_bwd_fused(d)")) return e;
  if (int e = make_tmap(&tmW, wab, 2 * C, Cin, ldw, Cin, "pw_bwd_fused(wab)")) return e;
  {
    PFN_encodeTiled fn = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)P};
    cuuint64_t strides[1] = {(cuuint64_t)lddd * 2};
    cuuint32_t box[2] = {64u, 128u};
    cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(&tmO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dd, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "pw_bwd_fused: cuTensorMapEncodeTiled(dd) failed with CUresult %d", (int)r);
  }
  PbParams p{};
  p.P = P; p.bias = bias; p.G = G; p.ldG = ldG; p.num_blocks = (int)ceil_div(P, 128);
  cudaStream_t st = (cudaStream_t)stream;
  return Cin == 64 ? launch_pw_bwd_fused<64>(tmG, tmZ, tmD, tmW, tmO, p, st) : launch_pw_bwd_fused<128>(tmG, tmZ, tmD, tmW, tmO, p, st);
}

extern "C" int unet_gemm_tc(const unet_gemm_args* a, void* stream) {
  if (int e = gemm_validate(a, "gemm_tc")) return e;
  if (a->in_dtype == UNET_F32) {
    // fp32 mode on the tensor cores: every operand is a (hi, lo) pair from unet_split_tf32, three kind::tf32 MMAs per k-step
    if (a->a_trans) {      // weight gradient C[M,N] += A[K,M]^T * B[K,N]: both operands MN-major, split over K
      UNET_REQUIRE(a->A_lo && a->B_lo && a->lda_lo >= a->M && a->ldb_lo >= a->N, UNET_EUNSUPPORTED,
                   "gemm_tc: fp32 operands need the lo parts of their tf32 split (A_lo, B_lo with pitches lda_lo >= M, ldb_lo >= N)");
      UNET_REQUIRE(a->b_trans == 0 && a->accumulate == 1 && !a->A2 && !a->B2 && a->epilogue == UNET_EPI_NONE && a->out_dtype == UNET_F32,
                   UNET_EUNSUPPORTED, "gemm_tc(fp32): a_trans=1 needs b_trans=0, accumulate=1, fp32 C and no epilogue");
      UNET_REQUIRE(a->M % 4 == 0 && a->N % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc(fp32): wgrad needs M%%4==0 and N%%8==0");
      TcParams p{};
      fill_params(p, a);
      p.split_kb = 1 << 30;
      const int bn = a->N > 64 ? 128 : 64;
      CUtensorMap tmA, tmAl, tmB, tmBl;
      if (int e = make_tmap(&tmA, a->A, a->M, a->K, a->lda, 32, "gemm_tc(wgrad A)", true, true)) return e;
      if (int e = make_tmap(&tmAl, a->A_lo, a->M, a->K, a->lda_lo, 32, "gemm_tc(wgrad A lo)", true, true)) return e;
      if (int e = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, 32, "gemm_tc(wgrad B)", true, true)) return e;
      if (int e = make_tmap(&tmBl, a->B_lo, a->N, a->K, a->ldb_lo, 32, "gemm_tc(wgrad B lo)", true, true)) return e;
      cudaStream_t st = (cudaStream_t)stream;
      return bn == 128 ? launch_wgrad<128, true>(tmA, tmB, tmBl, p, st, &tmAl) : launch_wgrad<64, true>(tmA, tmB, tmBl, p, st, &tmAl);
    }
    UNET_REQUIRE(a->A_lo && a->B_lo && a->lda_lo >= a->K && a->ldb_lo >= a->K, UNET_EUNSUPPORTED,
                 "gemm_tc: fp32 operands need the lo parts of their tf32 split (A_lo, B_lo with pitches lda_lo, ldb_lo >= K)");
    UNET_REQUIRE(!a->a_trans && a->b_trans == 1 && !a->accumulate && !a->A2 && !a->B2, UNET_EUNSUPPORTED,
                 "gemm_tc: the fp32 (tf32x3) path takes C = A * B^T with B given as [N,K] and no operand concatenation");
    UNET_REQUIRE(a->out_dtype == UNET_F32 && a->epilogue != UNET_EPI_HEAD, UNET_EUNSUPPORTED, "gemm_tc: the fp32 path writes fp32 and has no fused head");
    UNET_REQUIRE(a->N % 8 == 0 && a->K % 4 == 0, UNET_EUNSUPPORTED, "gemm_tc(fp32): N must be a multiple of 8 and K of 4");
    UNET_REQUIRE(a->ldc % 4 == 0 && aligned16(a->C), UNET_EALIGN, "gemm_tc: C needs a 16B-aligned base and ldc%%4==0");
    UNET_REQUIRE((!a->scale || aligned16(a->scale)) && (!a->shift || aligned16(a->shift)), UNET_EALIGN, "gemm_tc: scale / shift must be 16B aligned");
    if (a->epilogue == UNET_EPI_CONVT) {
      UNET_REQUIRE(a->M < (int64_t)1 << 31 && (a->N / 4) % 32 == 0, UNET_EUNSUPPORTED, "gemm_tc(fp32): CONVT needs M < 2^31 and Cout%%32==0");
      const int w = a->convt_W;
      UNET_REQUIRE(w > 0 && (w >= kBlockM ? w % kBlockM == 0 : kBlockM % w == 0), UNET_EUNSUPPORTED,
                   "gemm_tc: CONVT needs the input width to divide 128 or be a multiple of it (got %d)", w);
      if (a->drop.rate > 0.f)
        UNET_REQUIRE(a->drop.ctot % 4 == 0 && a->drop.c0 % 4 == 0, UNET_EUNSUPPORTED, "gemm_tc: CONVT dropout needs ctot and c0 to be multiples of 4");
    }
    TcParams p{};
    fill_params(p, a);
    p.split_kb = 1 << 30;
    const int bn = a->N > 64 ? 128 : 64;
    CUtensorMap tmA, tmAl, tmB, tmBl, tmC;
    if (int e = make_tmap(&tmA, a->A, a->K, a->M, a->lda, kBlockM, "gemm_tc(A hi)", true)) return e;
    if (int e = make_tmap(&tmAl, a->A_lo, a->K, a->M, a->lda_lo, kBlockM, "gemm_tc(A lo)", true)) return e;
    if (int e = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, bn, "gemm_tc(B hi)", true)) return e;
    if (int e = make_tmap(&tmBl, a->B_lo, a->K, a->N, a->ldb_lo, bn, "gemm_tc(B lo)", true)) return e;
    if (int e = make_c_tmap(&tmC, a, "gemm_tc(C)")) return e;
    cudaStream_t st = (cudaStream_t)stream;
    return bn == 128 ? launch_nt<128, false, true>(tmA, tmAl, tmB, tmC, p, st, &tmBl) : launch_nt<64, false, true>(tmA, tmAl, tmB, tmC, p, st, &tmBl);
  }
  UNET_REQUIRE(a->in_dtype == UNET_BF16, UNET_EUNSUPPORTED, "gemm_tc: operands must be bf16, or fp32 with their tf32 split");
  UNET_REQUIRE(a->N % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: N must be a multiple of 8 (got %lld)", (long long)a->N);
  UNET_REQUIRE(a->C == nullptr || (a->ldc % 4 == 0 && aligned16(a->C)), UNET_EALIGN, "gemm_tc: C needs a 16B-aligned base and ldc%%4==0");
  if (a->epilogue == UNET_EPI_HEAD)
    UNET_REQUIRE(a->N <= 64 && a->out_dtype == UNET_BF16 && !a->a_trans, UNET_EUNSUPPORTED,
                 "gemm_tc: the fused output head needs N <= 64 and bf16 activations");
  UNET_REQUIRE(!a->scale || aligned16(a->scale), UNET_EALIGN, "gemm_tc: scale must be 16B aligned");
  UNET_REQUIRE(!a->shift || aligned16(a->shift), UNET_EALIGN, "gemm_tc: shift must be 16B aligned");
  if (a->epilogue == UNET_EPI_CONVT)
    UNET_REQUIRE(a->M < (int64_t)1 << 31, UNET_EUNSUPPORTED, "gemm_tc: CONVT needs M < 2^31");
  if (a->epilogue == UNET_EPI_CONVT && a->drop.rate > 0.f)
    UNET_REQUIRE(a->drop.ctot % 4 == 0 && a->drop.c0 % 4 == 0, UNET_EUNSUPPORTED,
                 "gemm_tc: CONVT dropout needs ctot and c0 to be multiples of 4 (mask groups of four elements)");
  if (a->epilogue == UNET_EPI_CONVT) {
    UNET_REQUIRE((a->N / 4) % 64 == 0, UNET_EUNSUPPORTED, "gemm_tc: CONVT needs Cout%%64==0 (got %lld)", (long long)(a->N / 4));
    const int w = a->convt_W;   // a 128-row tile must be a whole box of input pixels: W | 128 or 128 | W
    UNET_REQUIRE(w > 0 && (w >= kBlockM ? w % kBlockM == 0 : kBlockM % w == 0), UNET_EUNSUPPORTED,
                 "gemm_tc: CONVT needs the input width to divide 128 or be a multiple of it (got %d)", w);
  }
  cudaStream_t st = (cudaStream_t)stream;
  TcParams p{};
  fill_params(p, a);
  CUtensorMap tmA, tmB;
  const int bn = pick_block_n(a->N);

  if (!a->a_trans) {
    UNET_REQUIRE(a->b_trans == 1, UNET_EUNSUPPORTED, "gemm_tc: forward/dgrad needs B given as [N,K] (b_trans=1)");
    UNET_REQUIRE(!a->accumulate, UNET_EUNSUPPORTED, "gemm_tc: accumulate is only implemented for a_trans=1");
    UNET_REQUIRE(a->K % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: K must be a multiple of 8 (got %lld)", (long long)a->K);
    UNET_REQUIRE(!a->B2, UNET_EUNSUPPORTED, "gemm_tc: B2 (N concatenation) is only implemented for a_trans=1");
    CUtensorMap tmA2;
    p.split_kb = 1 << 30;
    if (a->A2) {     // A = [A | A2] along K
      UNET_REQUIRE(a->k_split > 0 && a->k_split < a->K && a->k_split % kBlockK == 0 && a->lda2 >= a->K - a->k_split, UNET_EINVAL,
                   "gemm_tc: k_split must be a multiple of %d inside (0,K) and lda2 >= K - k_split", kBlockK);
      if (int e = make_tmap(&tmA, a->A, a->k_split, a->M, a->lda, kBlockM, "gemm_tc(A)")) return e;
      if (int e = make_tmap(&tmA2, a->A2, a->K - a->k_split, a->M, a->lda2, kBlockM, "gemm_tc(A2)")) return e;
      p.split_kb = (int)(a->k_split / kBlockK);
    } else {
      if (int e = make_tmap(&tmA, a->A, a->K, a->M, a->lda, kBlockM, "gemm_tc(A)")) return e;
      tmA2 = tmA;
    }
    if (int e = make_tmap(&tmB, a->B, a->K, a->N, a->ldb, bn, "gemm_tc(B)")) return e;
    CUtensorMap tmC = tmA;       // placeholder descriptor when nothing is stored (HEAD without C)
    if (a->C)
      if (int e = make_c_tmap(&tmC, a, "gemm_tc(C)")) return e;
    const bool ob = a->out_dtype == UNET_BF16;
    if (bn == 256) return ob ? launch_nt<256, true>(tmA, tmA2, tmB, tmC, p, st) : launch_nt<256, false>(tmA, tmA2, tmB, tmC, p, st);
    if (bn == 128) return ob ? launch_nt<128, true>(tmA, tmA2, tmB, tmC, p, st) : launch_nt<128, false>(tmA, tmA2, tmB, tmC, p, st);
    return ob ? launch_nt<64, true>(tmA, tmA2, tmB, tmC, p, st) : launch_nt<64, false>(tmA, tmA2, tmB, tmC, p, st);
  }
  // weight gradient: C[M,N] += A[K,M]^T * B[K,N]
  UNET_REQUIRE(a->b_trans == 0 && a->accumulate == 1, UNET_EUNSUPPORTED, "gemm_tc: a_trans=1 needs b_trans=0 and accumulate=1");
  UNET_REQUIRE(a->M % 8 == 0, UNET_EUNSUPPORTED, "gemm_tc: wgrad M must be a multiple of 8");
  UNET_REQUIRE(!a->A2, UNET_EUNSUPPORTED, "gemm_tc: A2 (K concatenation) is only implemented for a_trans=0");
  if (int e = make_tmap(&tmA, a->A, a->M, a->K, a->lda, 64, "gemm_tc(wgrad A)")) return e;
  CUtensorMap tmB2;
  p.split_kb = 1 << 30;
  if (a->B2) {       // B = [B | B2] along N
    UNET_REQUIRE(a->n_split > 0 && a->n_split < a->N && a->n_split % 64 == 0 && a->ldb2 >= a->N - a->n_split, UNET_EINVAL,
                 "gemm_tc: n_split must be a multiple of 64 inside (0,N) and ldb2 >= N - n_split");
    if (int e = make_tmap(&tmB, a->B, a->n_split, a->K, a->ldb, 64, "gemm_tc(wgrad B)")) return e;
    if (int e = make_tmap(&tmB2, a->B2, a->N - a->n_split, a->K, a->ldb2, 64, "gemm_tc(wgrad B2)")) return e;
    p.split_kb = (int)(a->n_split / 64);
  } else {
    if (int e = make_tmap(&tmB, a->B, a->N, a->K, a->ldb, 64, "gemm_tc(wgrad B)")) return e;
    tmB2 = tmB;
  }
  if (bn == 256) return launch_wgrad<256>(tmA, tmB, tmB2, p, st);
  if (bn == 128) return launch_wgrad<128>(tmA, tmB, tmB2, p, st);
  return launch_wgrad<64>(tmA, tmB, tmB2, p, st);
}
