/*
 * unet_b200.h — C-ABI of libunet_b200.so (sm_100a only).
 *
 * The reference (planck-epoch/unet-image-segmentation) has no FFI of its own:
 * its hot path is a Keras layer graph (model/u_net.py:5-116) plus tensor
 * algebra in utils/metrics.py:6-62 and utils/loss.py:9-48, executed by
 * TensorFlow kernels.  Each entry point below replaces the TensorFlow kernel(s)
 * behind one reference call site; the citation names that call site.
 *
 * Conventions
 *   - every function returns 0 (UNET_OK) on success, a negative UNET_E* code for
 *     a rejected argument, or a positive cudaError_t; unet_last_error() returns
 *     a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers owned by the caller unless the name says
 *     `host`; the library allocates nothing and never synchronises.
 *   - activations are NHWC.  `ld*` is the distance in ELEMENTS between two
 *     consecutive pixels (>= C), so a tensor may be a channel slice of a wider
 *     buffer (this is how Concatenate, u_net.py:96, becomes zero-copy).
 *   - `dtype` selects the storage type of activations (UNET_F32 / UNET_BF16);
 *     all arithmetic accumulates in fp32.  Parameters, statistics, losses and
 *     probabilities are always fp32.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - all launches are CUDA-graph capturable.
 */
#ifndef UNET_B200_H_
#define UNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNET_OK            0
#define UNET_EINVAL       -1   /* bad argument (null pointer, non-positive dim, bad enum) */
#define UNET_EUNSUPPORTED -2   /* legal request this build has no kernel for            */
#define UNET_EALIGN       -3   /* pointer / leading dimension violates an alignment rule */
#define UNET_EDRIVER      -4   /* driver entry point (tensor-map encode) unavailable     */

typedef enum { UNET_F32 = 0, UNET_BF16 = 1 } unet_dtype;

/* GEMM epilogues (fused producers of the ops that follow a contraction in u_net.py) */
typedef enum {
  UNET_EPI_NONE      = 0, /* C = acc                                  (dgrad)                          */
  UNET_EPI_AFFINE    = 1, /* C = acc*scale[n] + shift[n]              (folded BatchNormalization, u_net.py:23; or bias) */
  UNET_EPI_AFFINE_RELU = 2, /* C = max(acc*scale[n]+shift[n], 0)      (BN + Activation('relu'), u_net.py:23-25) */
  UNET_EPI_STATS     = 3, /* C = acc, and colsum[n] += C, colsq[n] += C*C over rows (training BN batch statistics) */
  UNET_EPI_CONVT     = 4, /* Conv2DTranspose(k=2,s=2) pixel-shuffle store + bias (+ dropout), u_net.py:88-98 */
  UNET_EPI_HEAD      = 5  /* y = max(acc*scale+shift,0) (stored only if C != NULL) and, from the same registers, the output head
                             Conv2D(classes,1,sigmoid|softmax) (u_net.py:105-112): head_out[m,:] = act(y[m,:] . head_w + head_b).
                             Inference only; tensor-core path only; needs N <= 64 (one column tile) and classes <= 8. */
} unet_epilogue;

/* stateless dropout mask: elements 4k .. 4k+3 (linear NHWC offsets in the ctot-wide tensor) share two 32-bit words
   a = mix(k ^ seed-mix), b = a * 0xC2B2AE3D; b ^= b >> 16 (csrc/common.cuh::dropout_words); element 4k+j keeps iff its 16-bit
   field (a.lo, a.hi, b.lo, b.hi) < floor((1-rate) * 65536).  rate == 0 disables.  Tensors of up to 2^34 elements. */
typedef struct {
  float    rate;      /* Dropout(rate), u_net.py:78,98 */
  uint32_t seed;
  int64_t  ctot;      /* channels of the logical tensor the mask is defined on (concat buffer width) */
  int64_t  c0;        /* channel offset of this view inside that tensor */
  const uint32_t* seed_dev; /* optional DEVICE word added to `seed` inside the kernel, so that a captured CUDA graph
                               draws a fresh mask on every replay (unet_step_advance increments it); NULL = 0 */
} unet_dropout;

typedef struct {
  /* C[M,N] (+)= A[M,K] * B[K,N] */
  int64_t M, N, K;
  const void* A; int64_t lda;      /* A row-major [M,K] (a_trans=0) or [K,M] (a_trans=1, wgrad)          */
  const void* B; int64_t ldb;      /* B row-major [K,N] (b_trans=0) or [N,K] (b_trans=1)                   */
  void*       C; int64_t ldc;      /* row-major [M,N]                                                     */
  int a_trans, b_trans;
  int in_dtype;                    /* dtype of A and B                                                     */
  int out_dtype;                   /* dtype of C; must be UNET_F32 when accumulate=1                       */
  int accumulate;                  /* 1: C += (atomic, split-K allowed) — used for weight gradients        */
  int epilogue;                    /* unet_epilogue                                                        */
  const float* scale;              /* [N] AFFINE*: per-column scale (NULL = 1)                             */
  const float* shift;              /* [N] AFFINE*: per-column shift;  CONVT: bias[Cout]  (NULL = 0)        */
  double* colsum; double* colsq;   /* [N] STATS accumulators (caller zeroes)                               */
  /* CONVT: A rows are pixels (n,i,j) of an [Nimg,H,W,K] tensor, N = 4*Cout ordered (a,b,co); C points at
     channel 0 of the [Nimg,2H,2W,*] destination with pixel stride ldc.                                    */
  int convt_H, convt_W;
  unet_dropout drop;               /* CONVT only */
  /* HEAD only: */
  const float* head_w;             /* [N, head_classes] (Keras output_mask kernel (1,1,N,classes))                          */
  const float* head_b;             /* [head_classes] or NULL                                                              */
  float*       head_out;           /* [M, head_classes] fp32 probabilities                                                */
  int          head_classes;
  /* Operand concatenation (tensor-core path only; used to fold BatchNormalization backward into the pointwise
     data / weight gradients, see unet_bn_bwd_coef).  split == 0 or NULL second operand: none.
       a_trans=0: A = [A | A2] along K — columns [0,k_split) from A (pitch lda), [k_split,K) from A2 (pitch lda2);
       a_trans=1: B = [B | B2] along N — columns [0,n_split) from B (pitch ldb), [n_split,N) from B2 (pitch ldb2).
     The split must be a multiple of 64. */
  const void* A2; int64_t lda2; int64_t k_split;
  const void* B2; int64_t ldb2; int64_t n_split;
  /* fp32 operands on the tensor cores (unet_gemm_tc, in_dtype = UNET_F32, a_trans = 0, b_trans = 1): A and B are the fp32
     tensors themselves (kind::tf32 ignores the low 13 mantissa bits: they act as hi = trunc(x)) and A_lo / B_lo the `lo` parts
     written by unet_split_tf32 (same shapes; row pitches lda_lo / ldb_lo).  The kernel issues lo*hi + hi*lo + hi*hi per k-step
     (fp32-grade products, error ~2^-20); NULL elsewhere. */
  const void* A_lo; int64_t lda_lo; const void* B_lo; int64_t ldb_lo;
} unet_gemm_args;

/* fp32 -> tf32 split for the tensor-core path of fp32 mode: lo = tf32_rna(x - trunc13(x)), where trunc13 clears the 13 mantissa
   bits kind::tf32 ignores, so that x itself is the `hi` operand.  hi (optional, may be NULL) receives x unchanged (useful with
   transpose).  src: [rows, cols] with row pitch ld; hi / lo: contiguous [rows, cols], or [cols, rows] when transpose != 0. */
int unet_split_tf32(const float* src, int64_t ld, int64_t rows, int64_t cols, float* hi, float* lo, int transpose, void* stream);

/* ---- library ---- */
int         unet_version(void);
int         unet_sm_arch(void);                /* 100: the only architecture this library is built for */
const char* unet_last_error(void);
int         unet_device_check(int device);     /* UNET_OK iff the device is compute capability 10.x    */

/* ---- SeparableConv2D, depthwise half (u_net.py:14-20) ---- */
/* y[n,i,j,c] = sum_{a,b} x'[n,i+a-1,j+b-1,c] * w[a,b,c], zero 'same' padding.
   x' = x, or max(x*in_scale[c]+in_shift[c],0) when in_scale != NULL (BN+ReLU of the producer fused into the load, so the
   producer's activation is never materialised; padding is zero in x' space: the TMA-strip kernel re-imposes it after the
   transform).  flip=1 correlates with the 180-degree rotated kernel (= gradient w.r.t. input).
   drop (rate>0) multiplies the OUTPUT by the dropout mask (used by the input-gradient of a dropped tensor).
   colsum (fp32 [C], accumulated, may be NULL): column sums over (n,i,j) of the stored outputs — the rank-1 term of the
   folded BatchNormalization backward (unet_bn_bwd_wgrad_combine); TMA-strip path only, not with dropout. */
int unet_dwconv3x3_fwd(const void* x, int64_t ldx, const float* w9c, void* y, int64_t ldy,
                       int N, int H, int W, int C, int dtype, int flip,
                       const float* in_scale, const float* in_shift,
                       const unet_dropout* drop, float* colsum, void* stream);
/* dw[a,b,c] += sum_{n,i,j} x[n,i+a-1,j+b-1,c] * dy[n,i,j,c]   (dw fp32 [3,3,C], accumulated atomically) */
int unet_dwconv3x3_bwd_weight(const void* x, int64_t ldx, const void* dy, int64_t lddy, float* dw9c,
                              int N, int H, int W, int C, int dtype, void* stream);

/* Both gradients of the depthwise half from ONE pass over dy (TMA strips; dy is read once, x needs no halo):
     dx = dy (*) rot180(w) [* dropout mask],   dw[a,b,c] += sum_{n,i,j} x[n,i+a-1,j+b-1,c] * dy[n,i,j,c].
   relu_mask=1: x is the post-ReLU output y of the producing conv_block (u_net.py:22-25), so (x > 0) is its ReLU mask and
   the stored dx becomes dL/d(BN output) of that block; bn_sums (fp32 [2,C], accumulated, may be NULL) then receives
   sum(dx) and sum(dx * y) over (n,i,j) — the two reductions BatchNormalization backward needs — from the stored values.
   drop (rate>0) multiplies dx by the Dropout mask for channels >= drop_c_from (a multiple of the 128-byte channel block;
   the channels below are left to another reader of dx, e.g. unet_convt_bwd_gather).
   x_scale/x_shift (fp32 [C], may be NULL): x is the producer's PRE-BatchNormalization output z and the activation
   y = max(z*x_scale + x_shift, 0) is formed on load (the producer's BN+ReLU pass never materialises y); not with dropout.
   up_out (may be NULL; relu_mask=0, even H and W): dx is the gradient of a skip-concat buffer whose first up_c channels
   came out of Conv2DTranspose(k=2,s=2) (u_net.py:88-96).  Those channels are stored un-pixel-shuffled into up_out
   [N*H/2*W/2, 4*up_c] — row (n, i/2, j/2), column block (i%2, j%2): the operand of the transposed convolution's weight /
   data gradient GEMMs, i.e. what unet_convt_bwd_gather would produce from dx — instead of into dx, and up_colsum
   (fp32 [up_c], accumulated, may be NULL) receives their per-channel sums (the Conv2DTranspose bias gradient).
   UNET_EUNSUPPORTED unless C % (8/sizeof(T)) == 0, C >= 8 and all views are 16-byte aligned. */
int unet_dwconv3x3_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, const float* w9c,
                       void* dx, int64_t lddx, float* dw9c, int N, int H, int W, int C, int dtype,
                       int relu_mask, float* bn_sums, const unet_dropout* drop, int drop_c_from,
                       const float* x_scale, const float* x_shift,
                       void* up_out, int up_c, float* up_colsum, void* stream);

/* ---- first conv_block, fused (enc1_block1_sepconv on the RGB image, u_net.py:14-20,63-66; Cin = 3, Cout = 64 only) ---- */
/* out = pw(dw(x)) [* scale + shift, ReLU if relu]; x contiguous [N,H,W,3]; with colsum/colsq also the BN batch statistics
   of the stored values.  UNET_EUNSUPPORTED for other channel counts (callers then use dwconv3x3_fwd + gemm). */
int unet_stem_fwd(const void* x, const float* wd9c, const float* wp, void* out, int64_t ldo,
                  int N, int H, int W, int Cin, int Cout, int dtype,
                  const float* scale, const float* shift, int relu, double* colsum, double* colsq,
                  float* d_out /* optional fp32 [N*H*W,3] workspace: the depthwise output is materialised there (kept for
                                  unet_stem_bwd_folded) and the pointwise half runs as a barrier-free stream over it */, void* stream);
/* given dz = gradient w.r.t. pw(dw(x)): dwp[3,64] += d^T dz (d = dw(x) recomputed on chip), dwd9c[3,3,3] += x (*) (dz Wp^T) */
int unet_stem_bwd(const void* x, const void* dz, int64_t lddz, const float* wd9c, const float* wp,
                  float* dwd9c, float* dwp, int N, int H, int W, int Cin, int Cout, int dtype, void* stream);
/* streaming backward of the first block with BatchNormalization backward folded in (no dz tensor): given g = dy*[y>0] (ptr,ldg),
   the saved pre-BN z (contiguous [M,64]), coef = [A|B|K] from unet_bn_bwd_coef and d3 from unet_stem_fwd:
   dz = A*g + B*z + K in registers; dwp[3,64] += d3^T dz; dd[M,3] (dtype) = dz Wp^T.  The depthwise weight gradient follows
   as unet_dwconv3x3_bwd_weight(x, dd). */
int unet_stem_bwd_folded(const void* g, int64_t ldg, const void* z, const float* coef, const float* d3,
                         const float* wp, float* dwp, void* dd, int64_t M, int dtype, void* stream);

/* ---- whole conv_block for inference (u_net.py:5-26): y = act((dw3x3(x) . Wp) * scale + shift), depthwise result kept on chip ---- */
/* bf16 only.  x: [N,H,W,Cin] view (ldx); wp_t: pointwise kernel TRANSPOSED, bf16 [Cout, Cin] (ldw); scale/shift: folded
   BatchNormalization (or NULL / bias); y: [N,H,W,Cout] view (ldy).  Cin <= 256, Cout <= 128, both multiples of 8.
   Optional fused output head (Cout <= 64; u_net.py:105-112): head_out[N*H*W, classes] = sigmoid|softmax(y . head_w + head_b);
   y may then be NULL (the last activation of the network is never written).
   pooled (may be NULL; needs y, even H and W): additionally MaxPooling2D((2,2)) (u_net.py:69) of the stored activation,
   [N,H/2,W/2,Cout] view (ldp) — the encoder's skip tensor is then not read back by unet_maxpool2x2_fwd. */
int unet_sepconv_fused_fwd(const void* x, int64_t ldx, const float* wd9c, const void* wp_t, int64_t ldw,
                           const float* scale, const float* shift, int relu, void* y, int64_t ldy,
                           int N, int H, int W, int Cin, int Cout,
                           const float* head_w, const float* head_b, float* head_out, int head_classes,
                           void* pooled, int64_t ldp, void* stream);

/* ---- dense contractions: SeparableConv2D pointwise half, Conv2DTranspose, their gradients ---- */
/* fp32-exact CUDA-core path (any shape, either dtype) */
int unet_gemm_simt(const unet_gemm_args* args, void* stream);
/* tcgen05/TMEM/TMA path: bf16 operands, fp32 accumulation.  Requirements: in_dtype = BF16; lda/ldb multiples of 8;
   a_trans=0: B must be given as [N,K] (b_trans=1);  a_trans=1 (weight gradient): b_trans=0, accumulate=1. */
int unet_gemm_tc(const unet_gemm_args* args, void* stream);

/* ---- BatchNormalization (u_net.py:23; Keras defaults momentum .99, eps 1e-3) ---- */
/* inference: scale = gamma/sqrt(var+eps), shift = beta - mean*scale  (moving statistics) */
int unet_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 float* scale, float* shift, int C, void* stream);
/* training: from colsum/colsq over `count` rows -> batch mean / biased var; scale/shift as above with batch stats;
   saves mean and rstd for backward; moving <- momentum*moving + (1-momentum)*batch.  gamma/beta NULL = 1/0. */
int unet_bn_finalize(const double* colsum, const double* colsq, int64_t count,
                     const float* gamma, const float* beta, float eps, float momentum,
                     float* moving_mean, float* moving_var,
                     float* scale, float* shift, float* save_mean, float* save_rstd, int C, void* stream);
/* y = max(z*scale+shift,0) (relu=1) written at (y,ldy), optionally dropped out; optional 2x2/2 max-pool of the
   un-dropped activation to `pooled` (contiguous [N,H/2,W/2,C]).  Activation + MaxPooling2D + Dropout,
   u_net.py:25,69,78,98.  z contiguous [N,H,W,C]. */
int unet_bn_act(const void* z, const float* scale, const float* shift, int relu,
                void* y, int64_t ldy, void* pooled,
                int N, int H, int W, int C, int dtype, const unet_dropout* drop, void* stream);
/* backward of y = relu(bn(z)): given dy (ptr,lddy) and saved z (contiguous),
   pass 1: dgamma[c] = sum g*xhat, dbeta[c] = sum g, g = dy*[y>0]   (accumulated into the fp32 outputs)
   pass 2: dz = scale*(g - dbeta/M - xhat*dgamma/M)                 (scale = gamma*rstd)
   `drop` (rate>0): dy is first multiplied by the dropout mask of the forward pass (y was stored dropped-out).
   save_mean == NULL: no normalisation (use_batch_norm=False); z is the pre-activation, dbeta is the bias gradient. */
int unet_bn_bwd_reduce(const void* dy, int64_t lddy, const void* z,
                       const float* scale, const float* shift, const float* save_mean, const float* save_rstd,
                       float* dgamma, float* dbeta, int64_t M, int C, int dtype, int relu,
                       const unet_dropout* drop, void* stream);
int unet_bn_bwd_apply(const void* dy, int64_t lddy, const void* z,
                      const float* scale, const float* shift, const float* save_mean, const float* save_rstd,
                      const float* dgamma, const float* dbeta, void* dz,
                      int64_t M, int C, int dtype, int relu, const unet_dropout* drop, void* stream);

/* BatchNormalization backward folded into the pointwise contractions (no dz tensor).  With g = dy*[y>0] produced — together
   with sums[0,c] = sum(g), sums[1,c] = sum(g*y) — by unet_dwconv3x3_bwd(relu_mask=1) (or any producer that has y in
   registers), BN backward is per-channel affine: dz = A*g + B*z + K.  This call turns the sums into
     dgamma[c] += sum(g*xhat) = (sums[1]-beta*sums[0])/gamma,  dbeta[c] += sums[0],  coef = [A | B | K] (fp32 [3,C]),
   and, when w (pointwise kernel, fp32 [Cin,C]) is given, the operands of the folded data gradient
     wab (bf16 [Cin,2C]) = [w diag(A) | w diag(B)],  bias[i] = sum_c K[c]*w[i,c]
   so that  dd = [g | z] * wab^T + bias  (unet_gemm_tc, A2 = z, UNET_EPI_AFFINE)  and
            dW += combine(d^T [g | z], coef, colsum(d))  (unet_gemm_tc a_trans, B2 = z;  unet_bn_bwd_wgrad_combine).
   ill_conditioned (device int, optional): set to 1 when some |gamma[c]| < |beta[c]|/16 (incl. gamma == 0), where recovering
   sum(g*xhat) from bf16-rounded activations loses precision; the caller then uses unet_bn_bwd_reduce/_apply instead. */
int unet_bn_bwd_coef(const float* sums, const float* gamma, const float* beta, const float* save_mean,
                     const float* save_rstd, int64_t count, float* dgamma, float* dbeta, float* coef,
                     const float* w, int Cin, int C, void* wab, float* bias, int* ill_conditioned, void* stream);
/* dw[i,c] += G[i,c]*A[c] + G[i,C+c]*B[c] + sd[i]*K[c];  G fp32 [Cin,2C] = d^T [g | z], coef = [A|B|K], sd[i] = sum_m d[m,i] */
int unet_bn_bwd_wgrad_combine(const float* G, const float* coef, const float* sd, float* dw, int Cin, int C, void* stream);
/* Both contractions above from ONE pass over [g | z] and d (tcgen05; bf16; C == 64, Cin in {64,128}; else UNET_EUNSUPPORTED
   and the caller issues the two unet_gemm_tc calls):  dd[P,Cin] = [g | z] * wab^T + bias  and  G[Cin,2C] += d^T [g | z].
   The same TMA-staged 128-pixel x 64-channel tiles serve as the K-major operand of the first product and the MN-major
   operands of the second, so [g | z] is read from HBM once (SeparableConv2D pointwise backward, u_net.py:14-23). */
int unet_pw_bwd_fused(const void* g, int64_t ldg, const void* z, int64_t ldz, const void* d, int64_t ldd,
                      const void* wab, int64_t ldw, const float* bias, void* dd, int64_t lddd,
                      float* G, int64_t ldG, int64_t P, int Cin, int C, void* stream);

/* ---- MaxPooling2D((2,2)) (u_net.py:69) ---- */
int unet_maxpool2x2_fwd(const void* x, int64_t ldx, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* dy_total[n,i,j,c] = dskip[n,i,j,c] (may be NULL) + dpool routed to the first max of each window, where the
   pooled activation is recomputed as relu(z*scale+shift) (scale NULL: the stored tensor is the activation).
   (H,W) are the un-pooled dims.  bn_sums (fp32 [2,C], accumulated, may be NULL; needs scale/shift): dy is additionally
   multiplied by the ReLU mask [y>0] and sum(dy), sum(dy*y) are accumulated (see unet_bn_bwd_coef).
   skip_drop (rate>0): dskip is first multiplied by the Dropout mask of the concat buffer it is a slice of (u_net.py:96-98). */
int unet_maxpool2x2_bwd(const void* z, int64_t ldz, const float* scale, const float* shift,
                        const void* dpool, const void* dskip, int64_t lddskip, void* dy,
                        int N, int H, int W, int C, int dtype, float* bn_sums, const unet_dropout* skip_drop, void* stream);

/* ---- Conv2DTranspose backward helper: un-pixel-shuffle dU[N,2H,2W,Cout] (ptr,ld) into G[N*H*W, 4*Cout]
        (columns (a,b,co)) and accumulate dbias[co] += sum dU; drop (rate>0): dU is first multiplied by the Dropout mask of
        the concat buffer it is a slice of ---- */
int unet_convt_bwd_gather(const void* du, int64_t lddu, void* g, float* dbias,
                          int N, int H, int W, int Cout, int dtype, const unet_dropout* drop, void* stream);

/* ---- output head: Conv2D(num_classes,1,activation) (u_net.py:105-112) + Dice/IoU sums (utils/metrics.py:29-31) ---- */
/* probs[M,C] fp32 = sigmoid (C==1) or softmax (C>1) of x[M,K]*w[K,C]+b.  If y_true != NULL also accumulates, per
   image n and class c, sums[n][c][0..2] += (sum t*p, sum t, sum p)   (double).  hw = pixels per image.
   x_scale/x_shift (fp32 [K], may be NULL; binary bf16 head with 64 contiguous channels only): x is dec1_block2's PRE-BatchNormalization
   tensor and max(x*x_scale + x_shift, 0) is formed on load, so that block's BN+ReLU pass never runs (same for unet_head_bwd). */
int unet_head_fwd(const void* x, int64_t ldx, const float* w, const float* b, float* probs,
                  const float* y_true, double* sums, int64_t M, int64_t hw, int K, int C, int dtype,
                  const float* x_scale, const float* x_shift, void* stream);
/* loss finalize (utils/loss.py:9-45, utils/metrics.py:33-38,58-62): kind 0 = dice, 1 = iou.
   out[0] = loss, out[1] = mean dice_coef, out[2] = mean iou_coef; coef[n][c][0..1] = (ca, cb) with
   dLoss/dp = ca*t + cb (scaled by grad_scale). */
int unet_seg_loss_finalize(const double* sums, int NC_pairs, float smooth, int kind, float grad_scale,
                           float* out3, float* coef, void* stream);
/* backward through activation + 1x1 conv: dx[M,K] (dtype), dw[K,C] +=, db[C] +=.
   bn_sums (fp32 [2,K], accumulated, may be NULL): x is the post-ReLU output of dec1_block2, so dx is multiplied by [x>0] and
   sum(dx), sum(dx*x) are accumulated for the folded BatchNormalization backward (see unet_bn_bwd_coef). */
int unet_head_bwd(const void* x, int64_t ldx, const float* w, const float* probs, const float* y_true,
                  const float* coef, void* dx, int64_t lddx, float* dw, float* db,
                  int64_t M, int64_t hw, int K, int C, int dtype, float* bn_sums,
                  const float* x_scale, const float* x_shift, void* stream);
/* standalone (I,T,P) sums over [NB,HW,C] fp32 pairs: dice_coef / iou_coef as metrics on arbitrary arrays */
int unet_seg_sums(const float* y_true, const float* y_pred, double* sums, int64_t NB, int64_t hw, int C, void* stream);

/* ---- tf.keras.metrics.MeanIoU (train.py:231, benchmark.py:237,269): counts[t*C+p] += 1 with int truncation ---- */
int unet_confusion_matrix_update(const float* y_true, const float* y_pred, int64_t n, int num_classes,
                                 unsigned long long* counts, void* stream);
/* thresholded variant used by benchmark.py:260: p = (prob > thr) */
int unet_confusion_matrix_update_thr(const float* y_true, const float* prob, float thr, int64_t n,
                                     unsigned long long* counts /*[4]*/, void* stream);
/* calculate_sample_iou (benchmark.py:159-170) on thresholded predictions (:260): per-sample 2x2 counts in one launch,
 * counts[nb*4 + t*2 + p] += 1; I = counts[3], T = counts[2]+counts[3], P = counts[1]+counts[3] per sample, and the sum over
 * samples is the MeanIoU(2) confusion matrix of the batch (:269) */
int unet_sample_confusion_thr(const float* y_true, const float* prob, float thr, int64_t NB, int64_t per_sample,
                              unsigned long long* counts /*[NB][4]*/, void* stream);

/* ---- AdamW, Keras form (train.py:226): w -= lr*wd*w; m,v update; w -= alpha*m/(sqrt(v)+eps) ---- */
/* hyper (device, fp32[8]): lr, wd, beta1, beta2, eps, step t (as float), grad_scale, unused — read on device so a
   captured graph follows ReduceLROnPlateau without re-capture. */
int unet_adamw_step(float* w, const float* g, float* m, float* v, int64_t n, const float* hyper, void* stream);
/* end of one optimizer step, on device: hyper[5] (t) += 1 and *counter += 1 (the dropout seed_dev word). Either may be NULL. */
int unet_step_advance(float* hyper, uint32_t* counter, void* stream);

/* ---- parameter staging for the tensor-core path: dst[r,c] = bf16(src[r,c]); dst_t[c,r] = bf16(src[r,c]).
        col_scale (fp32 [C], may be NULL) multiplies column c first: the folded BatchNormalization scale of inference
        (u_net.py:23) goes into the pointwise kernel, so the GEMM epilogue only adds the shift ---- */
int unet_cast_transpose_bf16(const float* src, void* dst, void* dst_t, int R, int C, const float* col_scale, void* stream);
/* the same for n matrices that live in one fp32 buffer, in one launch (the engine restages every dense kernel after each
   optimizer step).  table: DEVICE int64 [n][6] = {offset of the matrix in `base` (floats), dst address or 0, dst_t address or 0,
   R, C, index of its first 32x32 tile}; matrices in ascending tile order; total_tiles = sum of ceil(R/32)*ceil(C/32). */
int unet_cast_transpose_bf16_batched(const float* base, const int64_t* table, int n, int64_t total_tiles, void* stream);
/* dst = cast(src) elementwise between fp32/bf16 (n elements) */
int unet_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);

/* ---- pre/post-processing of the inference / benchmark CLIs (scripts/inference.py:98-110,147-160; scripts/benchmark.py:95-110) ---- */
/* img: DEVICE uint8 [H0,W0,C] as cv2.imread returns it (BGR kept) -> out[h,w,C] fp32 = cv2.resize(img / divisor, INTER_LINEAR) */
int unet_preprocess_u8(const uint8_t* img, int H0, int W0, int C, int64_t row_stride_bytes, float* out, int h, int w,
                       float divisor, void* stream);
/* prob: fp32 [h,w] with element stride ld (class channel of an NHWC tensor) -> mask[H0,W0] u8 = (cv2.resize(prob) > threshold) * 255 */
int unet_postprocess_mask(const float* prob, int h, int w, int64_t ld, uint8_t* mask, int H0, int W0, float threshold,
                          void* stream);

/* host helper: the 32-bit word holding element idx's 16-bit mask field (idx & 1 selects the half), for reproducing dropout on
   the host in tests */
uint32_t unet_host_dropout_hash(uint64_t idx, uint32_t seed);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H_ */
