// GPU pre/post-processing around model.predict for the inference / benchmark CLIs:
//   pre : cv2.imread image (uint8 HWC, BGR) -> float32 / 255 -> cv2.resize(INTER_LINEAR) to the model size
//         (reference scripts/inference.py:98-110, scripts/benchmark.py:95-110; the division happens BEFORE the resize)
//   post: probability mask -> cv2.resize(INTER_LINEAR) back to the original size -> (> threshold) * 255 (inference.py:147-160)
// cv2's INTER_LINEAR for float data: half-pixel centres, no antialiasing, source index clamped at the borders,
// horizontal pass then vertical pass, weights (1-f, f) in float.  One thread per output pixel.
#include "common.cuh"

namespace unet {

struct LinCoord { int i0, i1; float w0, w1; };

// cv2: f = (float)((d + 0.5) * scale - 0.5); s = floor(f); f -= s; clamp
__device__ __forceinline__ LinCoord lin_coord(int d, double scale, int src) {
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= src - 1) { s = src - 1; f = 0.f; }
  LinCoord c;
  c.i0 = s; c.i1 = min(s + 1, src - 1); c.w0 = 1.f - f; c.w1 = f;
  return c;
}

template <int CH>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ img, int H0, int W0, int64_t row_stride, float* __restrict__ out, int h, int w,
                     float divisor, double sy, double sx) {
  pdl_enter();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)h * w) return;
  const int y = (int)(t / w), x = (int)(t % w);
  const LinCoord cy = lin_coord(y, sy, H0), cx = lin_coord(x, sx, W0);
  const uint8_t* r0 = img + cy.i0 * row_stride;
  const uint8_t* r1 = img + cy.i1 * row_stride;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const float a00 = __fdiv_rn((float)r0[cx.i0 * CH + c], divisor), a01 = __fdiv_rn((float)r0[cx.i1 * CH + c], divisor);
    const float a10 = __fdiv_rn((float)r1[cx.i0 * CH + c], divisor), a11 = __fdiv_rn((float)r1[cx.i1 * CH + c], divisor);
    const float h0 = __fadd_rn(__fmul_rn(a00, cx.w0), __fmul_rn(a01, cx.w1));     // horizontal pass (no FMA contraction)
    const float h1 = __fadd_rn(__fmul_rn(a10, cx.w0), __fmul_rn(a11, cx.w1));
    out[t * CH + c] = __fadd_rn(__fmul_rn(h0, cy.w0), __fmul_rn(h1, cy.w1));      // vertical pass
  }
}

__global__ void __launch_bounds__(256)
postprocess_mask_kernel(const float* __restrict__ prob, int h, int w, int64_t ldp, uint8_t* __restrict__ mask, int H0, int W0,
                        float threshold, double sy, double sx) {
  pdl_enter();
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= (int64_t)H0 * W0) return;
  const int y = (int)(t / W0), x = (int)(t % W0);
  const LinCoord cy = lin_coord(y, sy, h), cx = lin_coord(x, sx, w);
  const float* r0 = prob + ((int64_t)cy.i0 * w) * ldp;
  const float* r1 = prob + ((int64_t)cy.i1 * w) * ldp;
  const float h0 = __fadd_rn(__fmul_rn(r0[cx.i0 * ldp], cx.w0), __fmul_rn(r0[cx.i1 * ldp], cx.w1));
  const float h1 = __fadd_rn(__fmul_rn(r1[cx.i0 * ldp], cx.w0), __fmul_rn(r1[cx.i1 * ldp], cx.w1));
  const float v = __fadd_rn(__fmul_rn(h0, cy.w0), __fmul_rn(h1, cy.w1));
  mask[t] = v > threshold ? 255 : 0;
}

}  // namespace unet

using namespace unet;

extern "C" int unet_preprocess_u8(const uint8_t* img, int H0, int W0, int C, int64_t row_stride_bytes, float* out, int h, int w,
                                  float divisor, void* stream) {
  UNET_REQUIRE(img && out && H0 > 0 && W0 > 0 && h > 0 && w > 0 && divisor > 0.f, UNET_EINVAL, "preprocess_u8: bad argument");
  UNET_REQUIRE(C == 1 || C == 3 || C == 4, UNET_EUNSUPPORTED, "preprocess_u8: 1, 3 or 4 channels (got %d)", C);
  UNET_REQUIRE(row_stride_bytes >= (int64_t)W0 * C, UNET_EINVAL, "preprocess_u8: row stride too small");
  const unsigned grid = (unsigned)ceil_div((int64_t)h * w, 256);
  const double sy = (double)H0 / h, sx = (double)W0 / w;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 1) launch_pdl(preprocess_u8_kernel<1>, grid, 256, 0, st, img, H0, W0, row_stride_bytes, out, h, w, divisor, sy, sx);
  else if (C == 3) launch_pdl(preprocess_u8_kernel<3>, grid, 256, 0, st, img, H0, W0, row_stride_bytes, out, h, w, divisor, sy, sx);
  else launch_pdl(preprocess_u8_kernel<4>, grid, 256, 0, st, img, H0, W0, row_stride_bytes, out, h, w, divisor, sy, sx);
  UNET_LAUNCH_CHECK("preprocess_u8");
  return UNET_OK;
}

extern "C" int unet_postprocess_mask(const float* prob, int h, int w, int64_t ld, uint8_t* mask, int H0, int W0, float threshold,
                                     void* stream) {
  UNET_REQUIRE(prob && mask && h > 0 && w > 0 && H0 > 0 && W0 > 0 && ld >= 1, UNET_EINVAL, "postprocess_mask: bad argument");
  const unsigned grid = (unsigned)ceil_div((int64_t)H0 * W0, 256);
  launch_pdl(postprocess_mask_kernel, grid, 256, 0, (cudaStream_t)stream, prob, h, w, ld, mask, H0, W0, threshold, (double)h / H0, (double)w / W0);
  UNET_LAUNCH_CHECK("postprocess_mask");
  return UNET_OK;
}
