// Inference conv_block in ONE kernel: SeparableConv2D (depthwise 3x3 + pointwise 1x1) + folded BatchNormalization +
// ReLU (reference model/u_net.py:5-26).  The depthwise result never touches HBM: CUDA-core warps compute it from a
// TMA-staged halo patch straight into the 128B-swizzled shared-memory tile that tcgen05.mma reads as its A operand.
//
// One CTA per SM, persistent over 8x16-pixel output patches (128 GEMM rows).  640 threads:
//   warp 0        TMA producer: per 64-channel block k, one 4-D box {64 ch, 18 cols, 10 rows, image} of x (zero OOB fill =
//                 'same' padding and ragged patches) + the {64 k, BLOCK_N} slice of the pointwise kernel
//   warp 1        MMA issuer: tcgen05.mma.cta_group::1.kind::f16, 128 x BLOCK_N x 16, accumulating over the channel blocks
//   warp 2        TMEM allocator (2 accumulator stages)
//   warps 4-11    depthwise producers, two groups of 4 warps, each owning one A-tile buffer and every second (tile, channel
//                 block) step: thread = (2 adjacent patch columns, 4 channels); slides down the 10 halo rows with two open
//                 partial sums per output in registers (4 conflict-free 8-byte shared loads + 2 x 9 packed FFMA2 per row),
//                 packs bf16 and stores into the A tile (K-major, SWIZZLE_128B) -> fence.proxy.async -> mbarrier
//   warps 12-19   two epilogue groups (one per accumulator stage): tcgen05.ld of the whole 64-column chunk -> accumulator stage
//                 released -> packed scale/shift (FFMA2 / FADD2), ReLU on the bf16 conversion (cvt.rn.relu.bf16x2.f32) ->
//                 swizzled staging tile -> one 4-D TMA store per 64-channel chunk into the (possibly channel-sliced) NHWC
//                 destination (r02 ncu: with one 64-channel block per patch the epilogue, not the depthwise producers, set the
//                 pace — 73 % of the epilogue warps' samples inside ~330 instructions per patch; the packed forms, the early
//                 release and carried patch coordinates took 0.951 -> 0.891 ms at 64->64 and 1.45 -> 1.385 ms at 128->64);
//                 optionally MaxPooling2D((2,2)) (u_net.py:69) of the staged tile -> a second 4-D TMA store of the 4x8 pooled
//                 patch, so the encoder's skip tensor is not read back for pooling
#include "common.cuh"
#include "ptx.cuh"

namespace unet {

// shared with gemm_tc.cu (same encodings)
__device__ __forceinline__ uint64_t fs_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((1024u >> 4) & 0x3fffu) << 32;   // SBO = 8 rows x 128 B
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}
__host__ __device__ constexpr uint32_t fs_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint32_t fs_max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void fs_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void fs_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fs_tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int kPH = 8, kPW = 16;                        // output patch = 128 pixels
constexpr int kXRows = kPH + 2, kXCols = kPW + 2;
constexpr int kXBytes = kXRows * kXCols * 128;          // 23040
constexpr int kXStage = 23 * 1024;                      // padded so that the B tile behind it is 1024 B aligned
constexpr int kTileBytes = 128 * 128;
constexpr int kMaxCin = 256;
constexpr int kPoolBytes = (kPH / 2) * (kPW / 2) * 128;   // pooled patch: 32 pixels x 128 B

template <int BLOCK_N> struct FsCfg {
  static constexpr int kBBytes = BLOCK_N * 128;
  static constexpr int kStageBytes = kXStage + kBBytes;
  // the stage ring bounds the number of patches in flight per SM (a stage is held from its TMA until the MMA that read its
  // pointwise-kernel slice retires): 4 stages where they fit in 227 KB
  static constexpr int kStages = BLOCK_N == 64 ? 4 : 3;
  // A-tile buffers per producer group.  ncu (r02) shows 40 % of the producers' stall samples waiting for their A tile to be
  // released, so a second buffer per group was tried (BLOCK_N = 64, paid for with the fourth TMA stage): SLOWER — 0.973 vs
  // 0.950 ms at 64->64 and 1.63 vs 1.45 ms at 128->64 (512x512, batch 64).  The wait is the pipeline's slack, not its limit;
  // the patches in flight (TMA stages) matter more.  Kept at one.
  static constexpr int kABufs = 1;
  static constexpr int kHeadFloats = BLOCK_N == 64 ? 8 * 64 + 8 : 0;     // fused output head: w[class][64] + bias[8], per group
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + 2 * kABufs * kTileBytes /*A*/ + 2 * kTileBytes /*staging*/ +
                                    2 * kPoolBytes /*pooled staging*/ +
                                    2 * (2 * BLOCK_N + kHeadFloats) * 4 /*scale,shift[,head] per group*/ +
                                    9 * kMaxCin * 4 /*dw weights*/ + 256;
};

struct FsParams {
  int H, W, Cin, Cout;
  int tiles_h, tiles_w, total_tiles, num_k;
  const float* wd9c; const float* scale; const float* shift;
  int store_y;                                                       // 0: the activation itself is not needed (head only)
  int store_pool;                                                    // also store the 2x2 max-pooled activation (tmP)
  const float* head_w; const float* head_b; float* head_out; int head_classes;   // optional fused 1x1 output head (Cout <= 64)
};

template <bool RELU> __device__ __forceinline__ uint32_t fs_pack(float2 v) {     // bf16x2 {lo = v.x, hi = v.y}
  uint32_t d;
  if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v.y), "f"(v.x));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(v.y), "f"(v.x));
  return d;
}

template <int BLOCK_N, bool RELU, bool HEAD>
__global__ void __launch_bounds__(640, 1)
sepconv_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmP, const FsParams p) {
  pdl_launch_dependents();
  using Cfg = FsCfg<BLOCK_N>;
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stages = smem;
  constexpr int NA = Cfg::kABufs;
  uint8_t* a_tiles = stages + S * Cfg::kStageBytes;                  // [group][NA][16 KB]
  uint8_t* out_tiles = a_tiles + 2 * NA * kTileBytes;                // [group][16 KB]
  uint8_t* pool_tiles = out_tiles + 2 * kTileBytes;                  // [group][4 KB]
  float* s_par = reinterpret_cast<float*>(pool_tiles + 2 * kPoolBytes);  // [group][2][BLOCK_N]
  float* s_wd = s_par + 2 * (2 * BLOCK_N + Cfg::kHeadFloats);        // [9][Cin]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wd + 9 * kMaxCin);
  uint64_t* ld_full = bars;              // [S]
  uint64_t* x_empty = bars + S;          // [S]
  uint64_t* a_full = bars + 2 * S;            // [group][NA]
  uint64_t* a_empty = bars + 2 * S + 2 * NA;  // [group][NA]
  uint64_t* tmem_full = bars + 2 * S + 4 * NA;
  uint64_t* tmem_empty = bars + 2 * S + 4 * NA + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * S + 4 * NA + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k = p.num_k;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmY); tma_prefetch_desc(&tmP); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&ld_full[i], 1); mbar_init(&x_empty[i], 5); }   // 4 producer warps + the MMA commit
    for (int i = 0; i < 2 * NA; ++i) { mbar_init(&a_full[i], 4); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 2 * BLOCK_N); tmem_relinquish(); }
  pdl_wait();            // above: shared memory, TMEM and kernel parameters only
  for (int i = threadIdx.x; i < 9 * p.Cin; i += blockDim.x) s_wd[i] = p.wd9c[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int wb = tile % p.tiles_w; const int t2 = tile / p.tiles_w;
        const int hb = t2 % p.tiles_h; const int n = t2 / p.tiles_h;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&x_empty[s], ph ^ 1);
          mbar_expect_tx(&ld_full[s], kXBytes + Cfg::kBBytes);
          uint8_t* st = stages + s * Cfg::kStageBytes;
          tma_load_4d(st, &tmX, &ld_full[s], kb * 64, wb * kPW - 1, hb * kPH - 1, n, 0x1000000000000000ull);
          tma_load_2d(st + kXStage, &tmB, &ld_full[s], kb * 64, 0, kEvictLast);
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = fs_idesc(BLOCK_N);
      int s = 0; uint32_t ph = 0; int step = 0; int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          // step = global (tile, channel block) counter: group step & 1 produced it, into that group's buffer (step >> 1) % NA
          const int ai = (step & 1) * NA + ((step >> 1) % NA);
          const uint32_t aph = (uint32_t)((step >> 1) / NA) & 1u;
          mbar_wait(&ld_full[s], ph);            // pointwise-kernel slice has landed
          mbar_wait(&a_full[ai], aph);           // depthwise tile is complete and visible to the async proxy
          tc_fence_after();
          const uint64_t adesc = fs_smem_desc(smem_u32(a_tiles + ai * kTileBytes));
          const uint64_t bdesc = fs_smem_desc(smem_u32(stages + s * Cfg::kStageBytes + kXStage));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&a_empty[ai]);
          umma_commit(&x_empty[s]);
          if (kb == num_k - 1) umma_commit(&tmem_full[acc]);
          if (++s == S) { s = 0; ph ^= 1; }
          ++step;
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================================================================== depthwise producers
    // Two groups of 4 warps; group g owns A-tile buffer g and produces every second (tile, channel block) step on its own, so
    // a thread covers TWO adjacent patch columns (4 shared loads per row for 2 outputs instead of 6) and, when the number of
    // channel blocks is 1 or 2, always meets the same channel block: its 9 taps stay in registers for the whole kernel.
    const int g = (warp - 4) >> 2;
    const int t = ((warp - 4) & 3) * 32 + lane;
    const int cp = t >> 4, cg = t & 15;                   // column pair (patch columns 2cp, 2cp+1), 4-channel group
    const uint32_t x_off = (uint32_t)(2 * cp) * 128u + (uint32_t)cg * 8u;
    const uint32_t at_base = smem_u32(a_tiles + g * NA * kTileBytes);
    float2 k9[9][2];                            // fp32 pairs: every FMA below is a packed FFMA2
    int k9_kb = -1;
    int mine = 0;                               // steps this group has produced
    int step = 0;                               // global (tile, channel block) counter; this group takes step % 2 == g
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < num_k; ++kb, ++step) {
        if ((step & 1) != g) continue;
        const int s = step % S; const uint32_t ph = (uint32_t)(step / S) & 1u;
        if (k9_kb != kb) {
          k9_kb = kb;
          const int c = kb * 64 + cg * 4;
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            if (c < p.Cin) {
              const float4 w4 = lds128f(smem_u32(s_wd + i * p.Cin + c));
              k9[i][0] = make_float2(w4.x, w4.y); k9[i][1] = make_float2(w4.z, w4.w);
            } else { k9[i][0] = k9[i][1] = make_float2(0.f, 0.f); }
          }
        }
        const int ab = mine % NA;                // this group's buffer for this step
        const uint32_t aph = (uint32_t)(mine / NA) & 1u;
        const uint32_t at_s = at_base + (uint32_t)ab * kTileBytes;
        mbar_wait(&ld_full[s], ph);
        mbar_wait(&a_empty[g * NA + ab], aph ^ 1);
        const uint32_t xs = smem_u32(stages + s * Cfg::kStageBytes) + x_off;
        const float2 z2 = make_float2(0.f, 0.f);
        float2 prev0[2] = {z2, z2}, cur0[2] = {z2, z2}, prev1[2] = {z2, z2}, cur1[2] = {z2, z2};
        // software-pipelined by one row: the shared loads of row r+1 are issued before the FMAs and the A-tile stores of row r
        // (the explicit ld/st.shared keep program order, so without this every row pays the full shared-load latency)
        uint2 n0 = lds64(xs), n1 = lds64(xs + 128), n2 = lds64(xs + 256), n3 = lds64(xs + 384);
#pragma unroll
        for (int r = 0; r < kXRows; ++r) {
          const uint2 r0 = n0, r1 = n1, r2 = n2, r3 = n3;          // halo columns 2cp .. 2cp+3 of this row
          if (r + 1 < kXRows) {
            n0 = lds64(xs + (r + 1) * kXCols * 128);
            n1 = lds64(xs + (r + 1) * kXCols * 128 + 128);
            n2 = lds64(xs + (r + 1) * kXCols * 128 + 256);
            n3 = lds64(xs + (r + 1) * kXCols * 128 + 384);
          }
          float2 q0[2], q1[2], q2[2], q3[2];
          q0[0] = make_float2(bf16lo_to_f32(r0.x), __uint_as_float(r0.x & 0xffff0000u)); q0[1] = make_float2(bf16lo_to_f32(r0.y), __uint_as_float(r0.y & 0xffff0000u));
          q1[0] = make_float2(bf16lo_to_f32(r1.x), __uint_as_float(r1.x & 0xffff0000u)); q1[1] = make_float2(bf16lo_to_f32(r1.y), __uint_as_float(r1.y & 0xffff0000u));
          q2[0] = make_float2(bf16lo_to_f32(r2.x), __uint_as_float(r2.x & 0xffff0000u)); q2[1] = make_float2(bf16lo_to_f32(r2.y), __uint_as_float(r2.y & 0xffff0000u));
          q3[0] = make_float2(bf16lo_to_f32(r3.x), __uint_as_float(r3.x & 0xffff0000u)); q3[1] = make_float2(bf16lo_to_f32(r3.y), __uint_as_float(r3.y & 0xffff0000u));
          // the FMAs of this halo row ordered by the column they read (the second operand): order-pinned FFMA2 runs whose
          // shared operand comes from the operand-reuse cache (see fma2v); summation order unchanged
          float2 o0[2], o1[2], pn0[2], pn1[2], cn0[2], cn1[2];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (r >= 2) o0[j] = fma2v(k9[6][j], q0[j], prev0[j]);
            if (r >= 1 && r <= kXRows - 2) pn0[j] = fma2v(k9[3][j], q0[j], cur0[j]);
            if (r <= kXRows - 3) cn0[j] = mul2v(k9[0][j], q0[j]);
            if (r >= 2) { o0[j] = fma2v(k9[7][j], q1[j], o0[j]); o1[j] = fma2v(k9[6][j], q1[j], prev1[j]); }
            if (r >= 1 && r <= kXRows - 2) { pn0[j] = fma2v(k9[4][j], q1[j], pn0[j]); pn1[j] = fma2v(k9[3][j], q1[j], cur1[j]); }
            if (r <= kXRows - 3) { cn0[j] = fma2v(k9[1][j], q1[j], cn0[j]); cn1[j] = mul2v(k9[0][j], q1[j]); }
            if (r >= 2) { o0[j] = fma2v(k9[8][j], q2[j], o0[j]); o1[j] = fma2v(k9[7][j], q2[j], o1[j]); }
            if (r >= 1 && r <= kXRows - 2) { pn0[j] = fma2v(k9[5][j], q2[j], pn0[j]); pn1[j] = fma2v(k9[4][j], q2[j], pn1[j]); }
            if (r <= kXRows - 3) { cn0[j] = fma2v(k9[2][j], q2[j], cn0[j]); cn1[j] = fma2v(k9[1][j], q2[j], cn1[j]); }
            if (r >= 2) o1[j] = fma2v(k9[8][j], q3[j], o1[j]);
            if (r >= 1 && r <= kXRows - 2) pn1[j] = fma2v(k9[5][j], q3[j], pn1[j]);
            if (r <= kXRows - 3) cn1[j] = fma2v(k9[2][j], q3[j], cn1[j]);
          }
          if (r >= 2) {
            const uint32_t m = (uint32_t)((r - 2) * kPW + 2 * cp);      // GEMM rows m (column 2cp) and m + 1
            const uint32_t hi = ((uint32_t)cg & 1u) << 3;
            sts64(at_s + m * 128u + ((((uint32_t)cg >> 1) ^ (m & 7u)) << 4) + hi,
                  make_uint2(pack_bf16x2(o0[0].x, o0[0].y), pack_bf16x2(o0[1].x, o0[1].y)));
            sts64(at_s + (m + 1u) * 128u + ((((uint32_t)cg >> 1) ^ ((m + 1u) & 7u)) << 4) + hi,
                  make_uint2(pack_bf16x2(o1[0].x, o1[0].y), pack_bf16x2(o1[1].x, o1[1].y)));
          }
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (r >= 1 && r <= kXRows - 2) { prev0[j] = pn0[j]; prev1[j] = pn1[j]; }
            if (r <= kXRows - 3) { cur0[j] = cn0[j]; cur1[j] = cn1[j]; }
          }
        }
        fs_fence_proxy_async();                  // generic-proxy stores -> visible to tcgen05 (async proxy)
        __syncwarp();
        if (lane == 0) { mbar_arrive(&a_full[g * NA + ab]); mbar_arrive(&x_empty[s]); }
        ++mine;
      }
    }
  } else if (warp >= 12) {
    // ===================================================================== epilogue groups
    const int g = (warp - 12) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int gtid = (warp - 12 - 4 * g) * 32 + lane;
    const bool issuer = gtid == 0;
    uint8_t* buf = out_tiles + g * kTileBytes;
    float* par_scale = s_par + g * (2 * BLOCK_N + Cfg::kHeadFloats);
    float* par_shift = par_scale + BLOCK_N;
    float* par_head = par_shift + BLOCK_N;
    constexpr bool kHeadCapable = HEAD;           // the fused output head is its own instantiation (BLOCK_N = 64)
    constexpr bool head = HEAD;
    for (int i = gtid; i < BLOCK_N; i += 128) {
      par_scale[i] = (i < p.Cout && p.scale) ? __ldg(p.scale + i) : 1.f;
      par_shift[i] = (i < p.Cout && p.shift) ? __ldg(p.shift + i) : 0.f;
    }
    if (kHeadCapable && head) {
      for (int i = gtid; i < 8 * 64; i += 128) {
        const int cls = i >> 6, k = i & 63;
        par_head[i] = (cls < p.head_classes && k < p.Cout) ? __ldg(p.head_w + (int64_t)k * p.head_classes + cls) : 0.f;
      }
      if (gtid < 8) par_head[8 * 64 + gtid] = (gtid < p.head_classes && p.head_b) ? __ldg(p.head_b + gtid) : 0.f;
    }
    fs_bar_sync(1 + g, 128);
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t my_row_s = smem_u32(buf) + (uint32_t)row * 128u;
    // patch coordinates advance by a fixed stride (2 * gridDim.x tiles): carried additions instead of two divisions per tile
    const int stride = 2 * (int)gridDim.x;
    const int st_w = stride % p.tiles_w, st_h = (stride / p.tiles_w) % p.tiles_h, st_n = stride / (p.tiles_w * p.tiles_h);
    int tile = blockIdx.x + g * gridDim.x;
    int wb = tile % p.tiles_w, hb = (tile / p.tiles_w) % p.tiles_h, n = tile / (p.tiles_w * p.tiles_h);
    int it = g;
    for (; tile < p.total_tiles; tile += stride, it += 2) {
      mbar_wait(&tmem_full[g], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + g * BLOCK_N;
      float hacc[kHeadCapable ? 8 : 1];
#pragma unroll
      for (int i = 0; i < (kHeadCapable ? 8 : 1); ++i) hacc[i] = 0.f;
      float2 h1a = make_float2(0.f, 0.f), h1b = h1a;
#pragma unroll
      for (int c = 0; c < BLOCK_N / 64; ++c) {
        if (c * 64 >= p.Cout) break;
        // the whole 64-column chunk of this row goes to registers at once, so the accumulator stage is handed back to the MMA
        // issuer before any arithmetic (the next patch's GEMM overlaps this epilogue)
        // (the head instantiation carries up to 8 class accumulators as well: it takes the chunk in two halves)
        uint32_t r[HEAD ? 1 : 2][32];
        const bool last_chunk = c == BLOCK_N / 64 - 1 || (c + 1) * 64 >= p.Cout;
        tmem_ld_32x32(t_addr + c * 64, r[0]);
        if (!HEAD) tmem_ld_32x32(t_addr + c * 64 + 32, r[HEAD ? 0 : 1]);
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has read `buf`
        fs_bar_sync(1 + g, 128);
        tmem_ld_wait();
        if (!HEAD && last_chunk) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[g]);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          if (HEAD && half == 1) {
            tmem_ld_32x32(t_addr + c * 64 + 32, r[0]);
            tmem_ld_wait();
            if (last_chunk) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty[g]);
            }
          }
          const float4* sc4 = reinterpret_cast<const float4*>(par_scale + c * 64 + half * 32);
          const float4* sh4 = reinterpret_cast<const float4*>(par_shift + c * 64 + half * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) {           // 8 channels = one 16-byte chunk of the staged row
            float2 v[4];
            const float4 h0 = sh4[2 * j], h1 = sh4[2 * j + 1];
            const uint32_t* rr = &r[HEAD ? 0 : half][8 * j];
            if (p.scale) {
              const float4 s0 = sc4[2 * j], s1 = sc4[2 * j + 1];
              v[0] = fma2(make_float2(__uint_as_float(rr[0]), __uint_as_float(rr[1])), make_float2(s0.x, s0.y), make_float2(h0.x, h0.y));
              v[1] = fma2(make_float2(__uint_as_float(rr[2]), __uint_as_float(rr[3])), make_float2(s0.z, s0.w), make_float2(h0.z, h0.w));
              v[2] = fma2(make_float2(__uint_as_float(rr[4]), __uint_as_float(rr[5])), make_float2(s1.x, s1.y), make_float2(h1.x, h1.y));
              v[3] = fma2(make_float2(__uint_as_float(rr[6]), __uint_as_float(rr[7])), make_float2(s1.z, s1.w), make_float2(h1.z, h1.w));
            } else {                              // scale folded into the pointwise kernel: half the broadcast shared loads
              v[0] = add2(make_float2(__uint_as_float(rr[0]), __uint_as_float(rr[1])), make_float2(h0.x, h0.y));
              v[1] = add2(make_float2(__uint_as_float(rr[2]), __uint_as_float(rr[3])), make_float2(h0.z, h0.w));
              v[2] = add2(make_float2(__uint_as_float(rr[4]), __uint_as_float(rr[5])), make_float2(h1.x, h1.y));
              v[3] = add2(make_float2(__uint_as_float(rr[6]), __uint_as_float(rr[7])), make_float2(h1.z, h1.w));
            }
            if (kHeadCapable && head) {           // 1x1 output convolution from the fp32 activations (not re-rounded to bf16)
              if (RELU) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = make_float2(fmaxf(v[i].x, 0.f), fmaxf(v[i].y, 0.f));
              }
              if (p.head_classes == 1) {          // the reference's default: two open packed chains, no per-class branches
                const float4* hw4 = reinterpret_cast<const float4*>(par_head + half * 32 + 8 * j);
                const float4 w0 = hw4[0], w1 = hw4[1];
                h1a = fma2(v[0], make_float2(w0.x, w0.y), h1a);
                h1b = fma2(v[1], make_float2(w0.z, w0.w), h1b);
                h1a = fma2(v[2], make_float2(w1.x, w1.y), h1a);
                h1b = fma2(v[3], make_float2(w1.z, w1.w), h1b);
              } else {
#pragma unroll
                for (int cls = 0; cls < 8; ++cls) {
                  if (cls < p.head_classes) {
                    const float4* hw4 = reinterpret_cast<const float4*>(par_head + cls * 64 + half * 32 + 8 * j);
                    const float4 w0 = hw4[0], w1 = hw4[1];
                    float2 a2 = fma2(v[0], make_float2(w0.x, w0.y), make_float2(hacc[cls], 0.f));
                    a2 = fma2(v[1], make_float2(w0.z, w0.w), a2);
                    a2 = fma2(v[2], make_float2(w1.x, w1.y), a2);
                    a2 = fma2(v[3], make_float2(w1.z, w1.w), a2);
                    hacc[cls] = a2.x + a2.y;
                  }
                }
              }
            }
            if (p.store_y) {                      // ReLU rides on the bf16 conversion (cvt.rn.relu.bf16x2.f32)
              const uint4 o = make_uint4(fs_pack<RELU>(v[0]), fs_pack<RELU>(v[1]), fs_pack<RELU>(v[2]), fs_pack<RELU>(v[3]));
              sts128(my_row_s + ((((uint32_t)(half * 4 + j)) ^ sw) << 4), o);
            }
          }
        }
        fs_fence_proxy_async();
        fs_bar_sync(1 + g, 128);
        if (p.store_pool) {
          // 2x2 max over the staged tile (row = h*16 + w, 16-byte chunk j of row r at ((j ^ (r & 7)) << 4)): 32 pooled pixels
          // x 8 chunks = 2 items per thread, bf16x2 max, into the swizzled pooled tile
          const uint32_t ps = smem_u32(pool_tiles + g * kPoolBytes);
#pragma unroll
          for (int it2 = 0; it2 < 2; ++it2) {
            const int item = gtid + it2 * 128;
            const int pp = item >> 3, j = item & 7;                 // pooled pixel (4 x 8), chunk
            const int r00 = ((pp >> 3) * 2) * kPW + (pp & 7) * 2;
            uint4 m4 = lds128u(smem_u32(buf) + (uint32_t)r00 * 128u + (((uint32_t)j ^ ((uint32_t)r00 & 7u)) << 4));
#pragma unroll
            for (int q = 1; q < 4; ++q) {
              const int rr = r00 + (q & 1) + (q >> 1) * kPW;
              const uint4 v4 = lds128u(smem_u32(buf) + (uint32_t)rr * 128u + (((uint32_t)j ^ ((uint32_t)rr & 7u)) << 4));
              m4.x = fs_max_bf16x2(m4.x, v4.x); m4.y = fs_max_bf16x2(m4.y, v4.y);
              m4.z = fs_max_bf16x2(m4.z, v4.z); m4.w = fs_max_bf16x2(m4.w, v4.w);
            }
            sts128(ps + (uint32_t)pp * 128u + (((uint32_t)j ^ ((uint32_t)pp & 7u)) << 4), m4);
          }
          fs_fence_proxy_async();
          fs_bar_sync(1 + g, 128);
        }
        if (issuer && (p.store_y || p.store_pool)) {
          if (p.store_y) fs_tma_store_4d(&tmY, buf, c * 64, wb * kPW, hb * kPH, n);
          if (p.store_pool) fs_tma_store_4d(&tmP, pool_tiles + g * kPoolBytes, c * 64, wb * (kPW / 2), hb * (kPH / 2), n);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (kHeadCapable && head) {
        const int hh = hb * kPH + (row >> 4), ww = wb * kPW + (row & 15);
        if (hh < p.H && ww < p.W) {
          const int ncls = p.head_classes;
          float* dst = p.head_out + (((int64_t)n * p.H + hh) * p.W + ww) * ncls;
          if (ncls == 1) {
            dst[0] = 1.f / (1.f + expf(-((h1a.x + h1a.y) + (h1b.x + h1b.y) + par_head[8 * 64])));
          } else {
            float mx = -INFINITY, e[8], lg[8], den = 0.f;
#pragma unroll
            for (int cls = 0; cls < 8; ++cls) if (cls < ncls) { lg[cls] = hacc[cls] + par_head[8 * 64 + cls]; mx = fmaxf(mx, lg[cls]); }
#pragma unroll
            for (int cls = 0; cls < 8; ++cls) if (cls < ncls) { e[cls] = expf(lg[cls] - mx); den += e[cls]; }
            const float inv = 1.f / den;
#pragma unroll
            for (int cls = 0; cls < 8; ++cls) if (cls < ncls) dst[cls] = e[cls] * inv;
          }
        }
      }
      wb += st_w; if (wb >= p.tiles_w) { wb -= p.tiles_w; ++hb; }
      hb += st_h; if (hb >= p.tiles_h) { hb -= p.tiles_h; ++n; }
      n += st_n;
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 2 * BLOCK_N);
}

static int fs_tmap_4d(CUtensorMap* map, const void* base, int64_t ld, int N, int H, int W, int C, int box_w, int box_h,
                      CUtensorMapSwizzle swz, const char* who) {
  PFN_encodeTiled fn = get_encode_fn();
  UNET_REQUIRE(fn, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled is not available from this driver", who);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return UNET_OK;
}

template <int BLOCK_N, bool RELU, bool HEAD = false>
static int fs_launch(const CUtensorMap& tmX, const CUtensorMap& tmB, const CUtensorMap& tmY, const CUtensorMap& tmP, const FsParams& p, cudaStream_t st) {
  using Cfg = FsCfg<BLOCK_N>;
  static SmemAttrOnce once;
  if (cudaError_t e = ensure_dynamic_smem(once, sepconv_fused_kernel<BLOCK_N, RELU, HEAD>, Cfg::kSmemBytes))
    return set_cuda_error(e, "sepconv_fused: cudaFuncSetAttribute");
  const unsigned grid = (unsigned)i64min(p.total_tiles, sm_count());
  launch_pdl(sepconv_fused_kernel<BLOCK_N, RELU, HEAD>, grid, 640, Cfg::kSmemBytes, st, tmX, tmB, tmY, tmP, p);
  UNET_LAUNCH_CHECK("sepconv_fused");
  return UNET_OK;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_sepconv_fused_fwd(const void* x, int64_t ldx, const float* wd9c, const void* wp_t, int64_t ldw,
                                      const float* scale, const float* shift, int relu, void* y, int64_t ldy,
                                      int N, int H, int W, int Cin, int Cout,
                                      const float* head_w, const float* head_b, float* head_out, int head_classes,
                                      void* pooled, int64_t ldp, void* stream) {
  UNET_REQUIRE(!pooled || (y && H % 2 == 0 && W % 2 == 0 && ldp >= Cout && ldp % 8 == 0 && aligned16(pooled)), UNET_EINVAL,
               "sepconv_fused: pooled needs y, even H and W, ldp >= Cout, ldp%%8==0 and a 16B-aligned base");
  UNET_REQUIRE(x && wd9c && wp_t && (y || head_out) && N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, UNET_EINVAL, "sepconv_fused: bad argument");
  UNET_REQUIRE(ldx >= Cin && (!y || ldy >= Cout) && ldw >= Cin, UNET_EINVAL, "sepconv_fused: leading dimension too small");
  UNET_REQUIRE(!head_out || (head_w && head_classes >= 1 && head_classes <= 8 && Cout <= 64), UNET_EINVAL,
               "sepconv_fused: the fused head needs head_w, 1 <= classes <= 8 and Cout <= 64");
  UNET_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cin <= kMaxCin && Cout <= 128, UNET_EUNSUPPORTED,
               "sepconv_fused: needs Cin%%8==0, Cout%%8==0, Cin <= %d, Cout <= 128 (got %d -> %d)", kMaxCin, Cin, Cout);
  UNET_REQUIRE(ldx % 8 == 0 && (!y || (ldy % 8 == 0 && aligned16(y))) && ldw % 8 == 0 && aligned16(x) && aligned16(wp_t) && aligned16(wd9c),
               UNET_EALIGN, "sepconv_fused: operands need 16B-aligned bases and ld%%8==0");
  UNET_REQUIRE((!scale || aligned16(scale)) && (!shift || aligned16(shift)), UNET_EALIGN, "sepconv_fused: scale/shift must be 16B aligned");
  CUtensorMap tmX, tmB, tmY, tmP;
  if (int e = fs_tmap_4d(&tmX, x, ldx, N, H, W, Cin, kXCols, kXRows, CU_TENSOR_MAP_SWIZZLE_NONE, "sepconv_fused(x)")) return e;
  if (y) { if (int e = fs_tmap_4d(&tmY, y, ldy, N, H, W, Cout, kPW, kPH, CU_TENSOR_MAP_SWIZZLE_128B, "sepconv_fused(y)")) return e; }
  else tmY = tmX;        // placeholder descriptor: nothing is stored through it
  if (pooled) { if (int e = fs_tmap_4d(&tmP, pooled, ldp, N, H / 2, W / 2, Cout, kPW / 2, kPH / 2, CU_TENSOR_MAP_SWIZZLE_128B, "sepconv_fused(pooled)")) return e; }
  else tmP = tmX;
  const int bn = Cout > 64 ? 128 : 64;
  {
    PFN_encodeTiled fn = get_encode_fn();
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)ldw * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(wp_t), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    UNET_REQUIRE(r == CUDA_SUCCESS, UNET_EDRIVER, "sepconv_fused(w): cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  }
  FsParams p{};
  p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.tiles_h = (int)ceil_div(H, kPH); p.tiles_w = (int)ceil_div(W, kPW);
  const int64_t tiles = (int64_t)N * p.tiles_h * p.tiles_w;
  UNET_REQUIRE(tiles < ((int64_t)1 << 31), UNET_EUNSUPPORTED, "sepconv_fused: too many tiles");
  p.total_tiles = (int)tiles; p.num_k = (int)ceil_div(Cin, 64);
  p.wd9c = wd9c; p.scale = scale; p.shift = shift;
  p.store_y = y != nullptr;
  p.store_pool = pooled != nullptr;
  p.head_w = head_w; p.head_b = head_b; p.head_out = head_out; p.head_classes = head_classes;
  cudaStream_t st = (cudaStream_t)stream;
  if (head_out) return relu ? fs_launch<64, true, true>(tmX, tmB, tmY, tmP, p, st) : fs_launch<64, false, true>(tmX, tmB, tmY, tmP, p, st);
  if (relu) return bn == 128 ? fs_launch<128, true>(tmX, tmB, tmY, tmP, p, st) : fs_launch<64, true>(tmX, tmB, tmY, tmP, p, st);
  return bn == 128 ? fs_launch<128, false>(tmX, tmB, tmY, tmP, p, st) : fs_launch<64, false>(tmX, tmB, tmY, tmP, p, st);
}
