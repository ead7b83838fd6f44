/*
 * pbt.h — C-ABI of libpbt.so, the B200 (sm_100a) native library behind the
 * patch-based-training hot path (GeneratorJ forward/backward, masked patch
 * sampler, full-frame inference).
 *
 * The reference (Mega-Gorilla/Video-to-Video_Few-Shot-Patch-Based-Training) is
 * pure Python and has NO FFI layer of its own; every entry point below states
 * the reference call site (file:line, relative to the reference root) whose
 * arithmetic it replaces.  INTEGRATION.md shows the ctypes binding a reference
 * maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types cross this boundary
 *  - every device pointer is owned by the caller (PyTorch allocates); the
 *    library allocates nothing and keeps no state
 *  - all work is enqueued on the caller's stream (a cudaStream_t passed as
 *    void*); no call synchronises; all are CUDA-graph capturable
 *  - return value: 0 = ok, negative = pbt_status error (pbt_error_string)
 *  - there is no CPU fallback: without a CUDA device every compute call
 *    returns PBT_ERR_CUDA
 *
 * Activation layout "P8" (planar, 8-channel interleaved, 16-bit elements):
 *      act[n][c/8][y][x][c%8]        C is a multiple of 8
 * i.e. each group of 8 channels is one H*W plane of 16-byte pixels.  A view
 * onto a channel range of a wider tensor is expressed by offsetting `ptr` by
 * whole planes and keeping the parent's `img_stride`.
 * fp32 tensors in "P8F" layout use the same indexing with 4-byte elements.
 */
#ifndef PBT_H_
#define PBT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBT_ABI_VERSION 2

typedef enum {
  PBT_OK = 0,
  PBT_ERR_ARG = -1,      /* bad shape / alignment / enum */
  PBT_ERR_CUDA = -2,     /* CUDA runtime or driver error (incl. no device) */
  PBT_ERR_UNSUPPORTED = -3,
  PBT_ERR_SMEM = -4      /* configuration does not fit shared/tensor memory */
} pbt_status;

typedef enum { PBT_BF16 = 0, PBT_FP16 = 1 } pbt_dtype;
typedef enum { PBT_ACT_NONE = 0, PBT_ACT_RELU = 1, PBT_ACT_LEAKY02 = 2 } pbt_act_kind;

/* 16-bit P8 activation view. */
typedef struct {
  void*   ptr;         /* device pointer, 16-byte aligned; NULL = absent */
  int32_t n, c, h, w;  /* c multiple of 8 */
  int64_t img_stride;  /* elements between consecutive images */
} pbt_act_t;

int         pbt_abi_version(void);
const char* pbt_error_string(int status);
/* last CUDA error text seen by this thread's most recent failing call */
const char* pbt_last_cuda_error(void);

/* ------------------------------------------------------------------------
 * Convolution as implicit GEMM on tcgen05 tensor cores (TMA-fed, TMEM fp32
 * accumulators).  Replaces every nn.Conv2d of GeneratorJ:
 *   src/models/generator.py:41,49 (ResNetBlock), :126 (conv11), :133,136
 *   (smoothers), :141 (output, fused as `head`), :171 (_make_conv_block),
 *   :200 (_make_upconv_block).
 * One entry point serves forward and dgrad (dgrad = the same stride-1 conv
 * with tap-flipped, channel-transposed packed weights).
 *
 * The conv is stride 1 with explicit top/left padding; output spatial size ==
 * input spatial size.  Stride-2 layers are expressed by the caller as 2x2
 * stride-1 convs over a space-to-depth input (see pbt_norm_apply s2d output).
 *
 * Packed weights (`wpack`, 16-bit): for channel block cb (blk_c channels,
 * last block may be shorter), tap t = ky*kw+kx:
 *      w[cb][t][k/8][cout][k%8]      k = channel inside the block
 * `blk_c` must be 16, 32 or 64 and a multiple of 16; cin multiple of 16.
 * ---------------------------------------------------------------------- */
typedef struct {
  pbt_act_t   in;            /* [n, cin, h, w] */
  const void* wpack;
  int32_t     cout;          /* multiple of 16, <= 256 */
  int32_t     kh, kw, pad_t, pad_l;
  int32_t     blk_c;         /* channels per K block: 16 / 32 / 64 */
  int32_t     tiles_per_cta; /* T in {1,2,3}: x-adjacent 8x16 pixel tiles sharing one haloed smem region; (8T+kw-1) <= 32 */
  int32_t     dtype;         /* pbt_dtype of in / wpack / out / mask */
  /* epilogue, applied in this order */
  const float* bias;         /* [cout] or NULL */
  int32_t      act;          /* pbt_act_kind */
  const float* post_scale;   /* [cout] affine after act (eval BatchNorm), or NULL */
  const float* post_shift;
  pbt_act_t    mask;         /* v = mask>0 ? v : 0 (ReLU backward); ptr NULL = off */
  const float* addend32;     /* P8F [n,cout,h,w] added before store, or NULL */
  float*       out32;        /* P8F fp32 store, or NULL */
  pbt_act_t    out;          /* 16-bit store, ptr NULL = off */
  float*       stats_partial;/* [n][tiles][2][cout] per-tile (sum, sumsq) of the stored values, or NULL */
  /* fused 1x1 head: tanh(head_w . v + head_b) -> fp32 NCHW [n,3,h,w]  (src/models/generator.py:141-144) */
  const float* head_w;       /* [3][cout] or NULL */
  const float* head_b;       /* [3] */
  float*       head_out;
  int32_t      head_tanh;    /* 1 = apply tanh */
  int32_t      upsample2x;   /* 1: `in` is [n,cin,h/2,w/2]; the conv consumes its bilinear x2 (align_corners=True) upsample,
                              * interpolated inside the kernel (src/models/generator.py:13 fused into :200); 3x3 pad 1 only */
  /* normalise-on-load (inference): `pre` (optional, ptr NULL = absent) supplies the FIRST pre.c input channels as the RAW
   * output of the previous conv; the kernel applies that layer's InstanceNorm + activation, act(x*pre_scale+pre_shift)
   * with pre_scale/pre_shift fp32 [n][pre.c], inside shared memory before the tensor core reads the tile
   * (src/models/generator.py:36-40 fused into the next conv).  `in` supplies the remaining channels (in.ptr NULL = none).
   * pre.c must be a multiple of blk_c; not combinable with upsample2x. */
  pbt_act_t    pre;
  const float* pre_scale;
  const float* pre_shift;
  int32_t      pre_act;
  int32_t      ctas_per_sm;  /* 0 or 2: default configuration (8 epilogue warps, two co-resident CTAs per SM); 4: request the
                              * small-footprint configuration for short CTAs (4 epilogue warps, four co-resident CTAs); used when
                              * tiles_per_cta*cout <= 128, the rings fit 56 KB and no on-load transform is active, else default */
  int32_t      cta_pair;     /* 1: CTA-pair configuration (clusters of two CTAs run one M=256 cta_group::2 MMA stream, each CTA
                              * stages half of the weight columns); `wpack` must be in the pair layout (pack mode bit 2).
                              * Needs blk_c 32 and cout % 32 == 0 (or blk_c 16, tiles_per_cta 2 with ctas_per_sm = 4); no upsample2x */
  int32_t      concurrent;   /* 1: kernels of other streams are expected to run next to this launch: never use persistent CTAs
                              * (they would hold every SM slot / tensor-memory column until the launch ends) */
  int32_t      batch_tiles;  /* 1: the tiles_per_cta (>= 2) tiles of a CTA are the SAME 8x16 tile of consecutive IMAGES instead of
                              * x-adjacent tiles of one image: on patch-sized maps (20x20 ... 40x40) the packed weights are then
                              * streamed from L2 once per T images and no tile columns are wasted.  `stats_partial` is indexed
                              * with pbt_conv_num_tiles(h, w, 1) tiles per image.  Excludes upsample2x / pre / cta_pair */
  int32_t      up_raw_channels; /* with upsample2x: the first up_raw_channels channels of `in` (multiple of blk_c) are the RAW output of the
                              * previous conv; act(x*pre_scale + pre_shift) (pre_scale / pre_shift fp32 [n][up_raw_channels], pre_act) is
                              * applied to the staged low-res tile before the interpolation, so the normalised tensor is never written */
  int32_t      tap_pairs;    /* 1 (first layer: 16-channel input whose real channels all sit in the first 8-channel plane, blk_c 16): one
                              * K = 16 MMA covers that plane at two horizontally adjacent taps (descriptor LBO = one pixel), 4 MMAs
                              * per 7-tap row instead of 7, and only the first plane is loaded.  `wpack` in pack mode bit 4 */
  int32_t      valid_h, valid_w; /* 0 = whole map; else outputs with y >= valid_h or x >= valid_w are stored as ZERO and left out of
                              * `stats_partial`: a 4x4 / pad 1 / stride 1 conv of the PatchGAN critic (src/models/discriminator.py:105-133)
                              * shrinks the map by one pixel; it runs on the fixed grid with (pad_t, pad_l) = (1, 1) and the zero
                              * border doubles as the next layer's padding */
  int32_t      debug_flags;  /* bring-up only; 0 in production */
  void*        debug_buf;    /* bring-up only: int64[grid][8] per-CTA phase timestamps, NULL in production */
} pbt_conv_desc_t;

/* One launch packs every conv parameter (nn.Conv2d layout [co][ci][kh][kw], fp32 or fp16 master copy) into the
 * operand layout above, optionally as the space-to-depth form of a stride-2 3x3 kernel (mode bit 0) and/or as
 * the dgrad kernel: channels transposed, taps flipped (mode bit 1).  `jobs` is a DEVICE array. */
typedef struct {
  const void* w;        /* source parameter */
  void*       dst;      /* packed 16-bit destination: taps * k_pad * n_out elements */
  int32_t     co, ci, kh, kw;   /* source dims */
  int32_t     mode;     /* bit 0: space-to-depth, bit 1: dgrad, bit 2: CTA-pair layout [cb][half][tap][k/8][n/2][8],
                         * bit 3: with bit 0, the source is a 4x4 stride-2 pad-1 kernel (critic): 3x3 stride-1 pad-1 kernel
                         * over the space-to-depth input, `reserved` = channels per phase of that input (>= ci, 0 = ci);
                         * bit 4: tap-pair layout of a first-layer kernel with ci <= 8: taps = kh * ceil(kw/2), K = 16 =
                         * (channels 0-7 at tap 2j, channels 0-7 at tap 2j+1), zero where 2j+1 == kw */
  int32_t     k_pad;    /* destination K channels (multiple of 16) */
  int32_t     n_out;    /* destination N rows (multiple of 16) */
  int32_t     n_keep;   /* rows taken from the source, remaining rows are zero */
  int32_t     blk_c;    /* channels per K block */
  int32_t     dtype;    /* pbt_dtype of dst */
  int32_t     src_is_half;
  int32_t     reserved;
} pbt_pack_job_t;
int pbt_pack_weights(const pbt_pack_job_t* jobs_dev, int32_t n_jobs, int64_t max_elems, void* stream);

/* ------------------------------------------------------------------------
 * Fused optimiser tail of the G-only step: clip_grad_norm_ (lightning_model.py:245-248) + Adam with L2 weight decay
 * (torch.optim.Adam semantics; lightning_model.py:326-329, config/optimizer/default.yaml:2-10) over a DEVICE table
 * of fp32 tensors in two launches.  state = device float[3 + 32*n_jobs]: [0] squared gradient norm of the last call,
 * [1] step count (advanced by one per call, so the call is CUDA-graph capturable), [2] skipped steps: when the gradient
 * norm is inf/nan (fp16 overflow in the backward sweep) nothing is updated and the step does not count (AMP semantics),
 * [3..] scratch: per-block partial sums of the norm, added up in a fixed order (no atomics), so that identical
 * gradients give bit-identical updates - data-parallel replicas stay bit-identical.  max_norm <= 0 disables clipping.
 * norm_out (optional) receives the un-clipped total gradient norm.
 * ---------------------------------------------------------------------- */
typedef struct {
  void*       param;       /* fp32 [count], updated in place */
  const void* grad;        /* fp32 [count] */
  float*      exp_avg;     /* fp32 [count] first moment */
  float*      exp_avg_sq;  /* fp32 [count] second moment */
  int64_t     count;
} pbt_optim_job_t;
int pbt_clip_adam_step(const pbt_optim_job_t* jobs_dev, int32_t n_jobs, int64_t max_elems, float* state, double max_norm,
                       double lr, double beta1, double beta2, double eps, double weight_decay, float* norm_out, void* stream);

/* ------------------------------------------------------------------------
 * Tiled inference mode (generator.py:427-565 `process_large_image`).  All tensors fp32, one frame at a time.
 * boxes_dev: int32 [n_tiles][4] = (y_start, y_end, x_start, x_end) as built by `_get_valid_patch_positions` (:359-398).
 * ---------------------------------------------------------------------- */
/* out[b] = frame window b, centred in a zero patch when the window is smaller than `patch` (:470-497) */
int pbt_tile_gather(const float* src /*[c][h][w]*/, int32_t channels, int32_t h, int32_t w, const int32_t* boxes_dev,
                    int32_t n_tiles, int32_t patch, float* out /*[n_tiles][c][patch][patch]*/, void* stream);
/* acc[:, y0+i, x0+j] += proc[b, :, i, j] * wt ; wsum[y0+i, x0+j] += wt, wt = weight_table[weight_index[b]][i][j] (:536-541);
 * accumulation by fp32 atomics (the order of overlapping tiles is not fixed) */
int pbt_tile_blend(const float* proc /*[n_tiles][3][patch][patch]*/, const int32_t* boxes_dev, const int32_t* weight_index_dev,
                   const float* weight_table /*[n_shapes][patch][patch]*/, int32_t n_tiles, int32_t patch, int32_t h, int32_t w,
                   float* acc /*[3][h][w]*/, float* wsum /*[h][w]*/, void* stream);
/* out = rgb*(1-mask) + (acc / (wsum > 1e-8 ? wsum : 1))*mask   (:553-560) */
int pbt_tile_finish(const float* acc, const float* wsum, const float* rgb /*[3][h][w]*/, const float* mask /*[h][w]*/,
                    int32_t h, int32_t w, float* out /*[3][h][w]*/, void* stream);

/* L1 reconstruction loss of the G-only step (lightning_model.py:267-268), value and gradient in one launch:
 * *loss = weight * mean|y - target| ;  gy = weight / count * sign(y - target).  All fp32, `count` elements. */
int pbt_l1_loss_fwd_bwd(const float* y, const float* target, int64_t count, float weight, float* loss, float* gy, void* stream);

/* number of stats tiles per image for a given geometry (tiles_x*tiles_y) */
int pbt_conv_num_tiles(int32_t h, int32_t w, int32_t tiles_per_cta);
int pbt_conv_fwd(const pbt_conv_desc_t* d, void* stream);

/* ------------------------------------------------------------------------
 * Weight gradient: dW[t][ci][co] += sum_pixels X[pix+t][ci] * dY[pix][co]
 * (autograd of the convs above; fires at lightning_model.py:241).
 * Output `dw` is fp32 [kh*kw][cin][cout], ACCUMULATED into (caller zeroes),
 * scaled by *inv_scale (device scalar, NULL = 1).
 * ---------------------------------------------------------------------- */
typedef struct {
  pbt_act_t   x;             /* conv input  [n, cin, h, w] */
  pbt_act_t   dy;            /* grad of conv output [n, cout, h, w] */
  int32_t     kh, kw, pad_t, pad_l;
  int32_t     dtype;
  float*      dw;            /* fp32 [kh*kw][cin][cout] */
  const float* inv_scale;
  int32_t     debug_flags;
} pbt_wgrad_desc_t;
int pbt_conv_wgrad(const pbt_wgrad_desc_t* d, void* stream);

/* ------------------------------------------------------------------------
 * Normalisation statistics (nn.InstanceNorm2d eps 1e-5 affine=False,
 * src/models/generator.py:36,87,178,204; nn.BatchNorm2d(64) :135).
 * partial: [n][tiles][2][c] from pbt_conv_fwd.
 * instance mode: scale[n][c] = rstd, shift[n][c] = -mean*rstd
 * batch mode   : statistics pooled over n; scale = gamma*rstd, shift = beta -
 *                mean*scale written for every n; running_mean/var updated
 *                (momentum, unbiased var) when non-NULL.
 * mean_out/rstd_out ([n][c], optional) are kept for the backward pass.
 * ---------------------------------------------------------------------- */
int pbt_norm_finalize(const float* partial /* consumed: folded in place */, int32_t n, int32_t tiles, int32_t c,
                      int64_t count_per_image, float eps, int32_t batch_mode,
                      const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum,
                      float* scale, float* shift, float* mean_out, float* rstd_out,
                      void* stream);

/* y = act(x*scale[n][c] + shift[n][c]) (+ residual32 | residual16); any of the outputs may be absent.
 * Replaces the InstanceNorm/BatchNorm + LeakyReLU/ReLU + residual add modules
 * (src/models/generator.py:36-58,93,99,103,116,120,135).
 * out_s2d (optional) receives the same values space-to-depth'ed:
 *   [n, 4*c, h/2, w/2], plane = ((y&1)*2+(x&1))*(c/8) + c/8-index. */
typedef struct {
  pbt_act_t    x;
  const float* scale;       /* [n][c] (or [c] when per_channel) ; NULL = identity */
  const float* shift;
  int32_t      per_channel; /* 1: scale/shift indexed [c] only */
  int32_t      act;
  const float* residual32;  /* P8F or NULL */
  pbt_act_t    out;         /* 16-bit y */
  pbt_act_t    out_relu;    /* 16-bit relu(y) */
  float*       out32;       /* P8F y */
  pbt_act_t    out_s2d;     /* 16-bit y, space-to-depth */
  int32_t      dtype;
  pbt_act_t    residual16;  /* 16-bit residual addend (ptr NULL = none; not together with residual32): the inference pass keeps
                             * the residual stream in 16 bits, like the reference's own `.half()` CUDA inference (generator.py:185) */
  /* fused InstanceNorm finalize (training passes, <= 64 tiles per image): when `partial` (the conv's [n][tiles][2][c] per-tile
   * sums) is given, scale = rstd and shift = -mean*rstd are computed inside this launch from it (count pixels per image, eps)
   * and WRITTEN to `scale` / `shift` for later consumers, replacing a separate pbt_norm_finalize launch */
  const float* partial;
  int32_t      tiles;
  int64_t      count;
  float        eps;
} pbt_norm_apply_desc_t;
int pbt_norm_apply(const pbt_norm_apply_desc_t* d, void* stream);

/* Bilinear x2 upsample, align_corners=True (src/models/generator.py:13), on a P8 view.
 * scale/shift ([n][c], optional) + act: the taps are normalised and activated on load, i.e. the kernel
 * computes upsample(act(in*scale+shift)) without materialising the normalised low-resolution tensor. */
int pbt_upsample2x(const pbt_act_t* in, const pbt_act_t* out, const float* scale, const float* shift, int32_t act,
                   int32_t dtype, void* stream);
/* its transpose (autograd): gin[h,w] (16-bit and/or fp32 P8F) = sum of gout[2h,2w] contributions */
int pbt_upsample2x_bwd(const pbt_act_t* gout, const pbt_act_t* gin16, float* gin32, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------
 * Backward of (norm -> act [-> + residual]) blocks.
 *  g = ga (16-bit P8, optionally s2d-indexed) + gb16 + gb32 (any may be absent)
 *  xhat = x*scale+shift ; gact = g * act'(xhat)
 *  reduce: sums[n][c] = (sum gact, sum gact*xhat) over h*w   (batch mode: over n too)
 *  apply : dx = k[n][c] * (gact - s1/cnt - xhat*s2/cnt), k = rstd (instance) or gamma*rstd (batch)
 * ---------------------------------------------------------------------- */
typedef struct {
  pbt_act_t    x;           /* raw conv output saved by the forward pass */
  const float* scale;       /* rstd and -mean*rstd, so that xhat = x*scale + shift (for InstanceNorm these are */
  const float* shift;       /* exactly the arrays given to pbt_norm_apply)                                    */
  int32_t      per_channel;
  int32_t      act;
  pbt_act_t    ga;          /* grad wrt y, 16-bit */
  int32_t      ga_is_s2d;   /* ga is [n,4c,h/2,w/2] space-to-depth of the logical grad */
  pbt_act_t    gb16;        /* second 16-bit addend or absent */
  const float* gb32;        /* fp32 P8F addend or NULL */
  float*       sums;        /* [n][2][c] (instance) or [2][c] (batch) fp32, zeroed by caller */
  const float* kmul;        /* [n][c] or [c]: rstd or gamma*rstd */
  int64_t      count;       /* elements per statistic (h*w or n*h*w) */
  int32_t      batch_mode;
  pbt_act_t    dx;          /* 16-bit output of apply */
  int32_t      dtype;
  int32_t      relu_mask_x; /* 1: dx is additionally zeroed where x <= 0 (x is itself a ReLU output, e.g. BatchNorm input) */
} pbt_norm_bwd_desc_t;
int pbt_norm_bwd_reduce(const pbt_norm_bwd_desc_t* d, void* stream);
int pbt_norm_bwd_apply(const pbt_norm_bwd_desc_t* d, void* stream);
/* reduce + apply in ONE launch for per-image statistics (batch_mode = 0): a CTA per (image, 8-channel plane) sums its
 * slice and re-reads it from L2 for the apply.  Meant for patch-sized maps with n * c/8 >= the SM count; writes (not
 * accumulates) d->sums. */
int pbt_norm_bwd_fused(const pbt_norm_bwd_desc_t* d, void* stream);

/* Backward of the fused head (tanh(1x1 conv), src/models/generator.py:141-144) and of the
 * ReLU in front of it:  gz = gy*(1-y^2)*gscale ; dW[3][c] += gz (x) s ; db += gz ;
 * gs = (s>0) ? W^T gz : 0 -> 16-bit P8 ; dbias_prev[c] += sum gs (bias grad of the conv that made s).
 * gy, y: fp32 NCHW [n,3,h,w]; s: P8 [n,c,h,w]; grads fp32 accumulated (caller zeroes),
 * left multiplied by gscale (undo with inv_scale). */
int pbt_head_bwd(const float* gy, const float* y, const pbt_act_t* s, const float* head_w,
                 const float* gscale, int32_t head_tanh,
                 float* dw, float* db, const pbt_act_t* gs, float* dbias_prev,
                 int32_t dtype, void* stream);

/* per-channel sum over n,h,w of a 16-bit P8 tensor, accumulated into out[c] times *inv_scale (bias grads) */
int pbt_channel_sum(const pbt_act_t* g, float* out, const float* inv_scale, int32_t dtype, void* stream);

/* ------------------------------------------------------------------------
 * Layout / dtype conversion at the module boundary
 * (GeneratorJ.forward takes NCHW fp32/fp16: src/models/generator.py:210).
 * ---------------------------------------------------------------------- */
/* x NCHW (fp32 if src_is_half==0 else fp16) [n,c,h,w] -> P8 `out` (out.c >= c, extra channels zeroed) */
int pbt_nchw_to_p8(const void* x, int32_t src_is_half, int32_t n, int32_t c, int32_t h, int32_t w,
                   const pbt_act_t* out, int32_t dtype, void* stream);
/* P8 (first c channels) -> NCHW fp32 */
int pbt_p8_to_nchw_f32(const pbt_act_t* in, int32_t c, float* out, float mul, int32_t dtype, void* stream);
/* P8F fp32 -> NCHW fp32 */
int pbt_p8f_to_nchw_f32(const float* in, int32_t n, int32_t c_total, int32_t c, int32_t h, int32_t w, float* out, void* stream);

/* uint8 HWC frame -> P8 in [-1,1] : ToTensor + Normalize(0.5,0.5) (generator.py:91-95,584-616) */
int pbt_u8hwc_to_p8(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t c,
                    const pbt_act_t* out, int32_t dtype, void* stream);
/* fp32 NCHW [-1,1] -> uint8 HWC : clamp, (x+1)*127.5, round (generator.py:643-647) */
int pbt_nchw_to_u8hwc(const float* y, int32_t n, int32_t c, int32_t h, int32_t w, uint8_t* out, void* stream);
/* uint8 HWC -> fp32 CHW exactly as torchvision ToTensor + Normalize(0.5, 0.5) (src/data/dataset.py:34-38) */
int pbt_u8hwc_to_norm_chw(const uint8_t* img, int32_t h, int32_t w, int32_t c, float* out, void* stream);

/* ------------------------------------------------------------------------
 * Patch sampler gather (StyleTransferDataset._cut_patch + default_collate +
 * torch.cat, src/data/dataset.py:209-232,262-273; lightning_model.py:211-221).
 * For every patch b and source s:
 *    out_s[b][:, 0:hx-hn, 0:xx-xn] = src_s[img_b][:, hn:hx, xn:xx], rest 0,
 *    hn=max(0,y-P/2) hx=min(y+P/2,H-1) xn=max(0,x-P/2) xx=min(x+P/2,W-1)
 * src_ptrs: device array [n_src][n_images] of const float* (CHW fp32, 3 channels... `ch` each)
 * img_hw  : device array [n_images][2] int32 (H, W)
 * pos     : device array [n_patches][3] int32 (img, y, x)
 * outs    : host array [n_src] of float* ; out_ch_off/out_ch_total give the
 *           channel slot of each source inside its output tensor so that
 *           'pre' and the guides land concatenated (dim=1) in one tensor.
 * ---------------------------------------------------------------------- */
int pbt_patch_gather(const float* const* src_ptrs, int32_t n_src, int32_t n_images, int32_t ch,
                     const int32_t* img_hw, const int32_t* pos, int32_t n_patches, int32_t patch,
                     float* const* outs, const int32_t* out_ch_off, const int32_t* out_ch_total,
                     void* stream);

/* 7x7 box-sum != 0 of a binary mask (dilation), src/data/dataset.py:157-170: out[y][x] = 1 if any
 * mask pixel > 0 in the 7x7 window. mask: uint8 [h][w] (already thresholded 0/255). */
int pbt_mask_dilate7(const uint8_t* mask, int32_t h, int32_t w, uint8_t* out, void* stream);

/* The residual trunk of GeneratorJ (reference src/models/generator.py:18-58,107-110,223-224) on patch-sized maps as ONE launch:
 * r_{b+1} = r_b + IN(convB(relu(IN(convA(relu(r_b)))))) for b = 0 .. n_blocks-1, 128 -> 128 channels, 3x3 pad 1, InstanceNorm
 * (biased variance, eps) over each image.  One CTA owns one image and keeps the running activation in shared memory (row pitch
 * w+2: the pad pixels are the zero halo); requires 128 channels and h*(w+2) <= 512 (pbt_res_trunk_supported).  It writes
 * everything the layer-by-layer path saves for the backward sweep:
 *   a[0]           in : relu(r_0)                                  a[b+1]   out: relu(r_{b+1}) (b+1 < n_blocks)
 *   raw_a[b]       out: convA output, 16 bit                       hmid[b]  out: relu(IN(raw_a[b]))
 *   raw_b[b]       out: convB output, 16 bit                       scale_x / shift_x [n][128] out: rstd, -mean*rstd of raw_x[b]
 *   residual32     in/out: r_0 -> r_{n_blocks-1} fp32 [n][16][h][w][8] (the last block's sum only goes to last16)
 *   last16         out: r_{n_blocks} in 16 bit (may be a channel view)
 * w_a / w_b: pbt_pack_weights output (forward form, blk_c 32, k_pad 128, n_out 128).  Conv biases are not applied: a constant
 * per channel is removed by the InstanceNorm that follows. */
#define PBT_TRUNK_MAX_BLOCKS 16
typedef struct {
  int32_t n_blocks, dtype;
  float eps;
  int32_t reserved;
  pbt_act_t a[PBT_TRUNK_MAX_BLOCKS];
  pbt_act_t raw_a[PBT_TRUNK_MAX_BLOCKS];
  pbt_act_t hmid[PBT_TRUNK_MAX_BLOCKS];
  pbt_act_t raw_b[PBT_TRUNK_MAX_BLOCKS];
  const void* w_a[PBT_TRUNK_MAX_BLOCKS];
  const void* w_b[PBT_TRUNK_MAX_BLOCKS];
  float* scale_a[PBT_TRUNK_MAX_BLOCKS];
  float* shift_a[PBT_TRUNK_MAX_BLOCKS];
  float* scale_b[PBT_TRUNK_MAX_BLOCKS];
  float* shift_b[PBT_TRUNK_MAX_BLOCKS];
  float* residual32;
  pbt_act_t last16;
} pbt_res_trunk_desc_t;
int pbt_res_trunk_supported(int32_t channels, int32_t h, int32_t w);   /* 1 / 0 */
int pbt_res_trunk_fwd(const pbt_res_trunk_desc_t* d, void* stream);

/* Perceptual (VGG feature) term of the generator loss, reference src/models/perception.py:93-143 and
 * lightning_model.py:270-275: loss = mean((features(generated) - features(target))^2) over the concatenated taps.
 * `f` is ONE feature tensor of 2*n_pairs images - [0, n_pairs) from the generated patches, [n_pairs, 2*n_pairs) from the
 * targets - so both feature passes are a single batch through the conv kernels.  flags: 1 = `g` already holds the gradient
 * arriving from the layers behind `f` (it is added to), 2 = `f` is a ReLU output (the gradient is masked with f > 0),
 * 4 = `f` is a tap: *loss += loss_mul * sum (f_gen - f_tgt)^2 and g += grad_mul * (f_gen - f_tgt).  `g` (n_pairs images,
 * same c/h/w; may be NULL when only the value is wanted) is dL/df of the generated half in the caller's fixed gradient scale.
 * `partial`: scratch of >= 4096 floats; `counter`: one uint32, zero before the first call (the kernel returns it to zero);
 * the sum is formed in a fixed order (no float atomics), so *loss is reproducible. */
int pbt_feature_mse(const pbt_act_t* f, int32_t n_pairs, float grad_mul, int32_t flags, const pbt_act_t* g, float* partial,
                    uint32_t* counter, float* loss, float loss_mul, int32_t dtype, void* stream);
/* 2x2 stride-2 max pooling of a P8 tensor (torchvision VGG `features[4]`, floor mode) and its transpose: the gradient of a
 * window goes to its first maximum in row-major order; dx may cover only the first dx->n images of x */
int pbt_maxpool2(const pbt_act_t* x, const pbt_act_t* y, int32_t dtype, void* stream);
int pbt_maxpool2_bwd(const pbt_act_t* x, const pbt_act_t* dy, const pbt_act_t* dx, int32_t dtype, void* stream);

/* zero everything outside the valid window [0,valid_h) x [0,valid_w) of a 16-bit P8 tensor (critic maps that shrink by one
 * pixel per 4x4 stride-1 layer live on a fixed grid; the zero border is the next layer's padding) */
int pbt_zero_border(const pbt_act_t* t, int32_t valid_h, int32_t valid_w, void* stream);
/* space-to-depth P8 [n, 4*cpp, h, w] (channel = phase*cpp + c, phase = (y&1)*2 + (x&1)) -> NCHW fp32 [n, c, 2h, 2w], values
 * multiplied by *mul_dev (device scalar, NULL = 1): the critic's gradient w.r.t. its input patches */
int pbt_p8s2d_to_nchw_f32(const pbt_act_t* in, int32_t cpp, int32_t c, float* out, const float* mul_dev, int32_t dtype, void* stream);

/* 7x7 erosion of thresholded masks (generator.py:327-351 `_process_mask`, :627-631): out[i][y][x] = 1.0f where all 49
 * pixels of the zero-padded 7x7 window are set, else 0.0f.  mask: uint8 [n][h][w] (0 / non-zero), out: float [n][h][w]. */
int pbt_mask_erode7(const uint8_t* mask, int32_t n, int32_t h, int32_t w, float* out, void* stream);
/* mask composite + uint8 conversion of the frame loop (generator.py:562-563,643-647):
 * out[n][h][w][3] = round(clamp((clamp(rgb*(1-m) + y*m, -1, 1) + 1)*127.5, 0, 255)) with rgb = ToTensor+Normalize of the
 * first three channels of frame uint8 [n][h][w][c]; y: float NCHW [n][3][h][w]; mask: float [n][h][w] or NULL (then
 * out is the plain conversion of y and `frame` is not read). */
int pbt_composite_to_u8(const float* y, const uint8_t* frame, int32_t c, const float* mask, int32_t n, int32_t h, int32_t w,
                        uint8_t* out, void* stream);

/* ------------------------------------------------------------------------
 * Loss / optimiser tail helpers.
 * ---------------------------------------------------------------------- */
/* max |g| over count fp32 values -> *out (device); used for dynamic fp16 gradient scaling */
int pbt_absmax_f32(const float* g, int64_t count, float* out, void* stream);
/* scale[0] = 2^k with amax*2^k in [target/2, target], scale[1] = 1/scale[0] (scale[0]=1 when amax==0);
 * adjust (optional device float) multiplies the target: the overflow back-off maintained by pbt_grad_scale_feedback */
int pbt_make_grad_scale(const float* amax, float target, float* scale2, const float* adjust, void* stream);
/* AMP-style feedback for the fp16 gradient scale, device-side (graph capturable).  probe = the most downstream fp32
 * gradient of the sweep; adjust = device float[3]: [0] target multiplier for the next sweep (/16 when the probe holds
 * inf/nan, x2 after 256 clean sweeps, within [2^-20, 1]), [1] clean sweeps since the last change, [2] overflow count */
int pbt_grad_scale_feedback(const float* probe, int64_t count, float* adjust, void* stream);

/* ------------------------------------------------------------------------
 * Host-side sampler bookkeeping (no CUDA): order-statistics tree that replaces
 * `valid_indices_left[img].pop(center_idx)` (src/data/dataset.py:254-256), which
 * is the c-th smallest not-yet-used index.  `tree` is caller-owned int32[n+1].
 * ---------------------------------------------------------------------- */
void    pbt_ostree_reset(int32_t* tree, int32_t n);
/* returns the k-th (0-based) remaining index and removes it; -1 if k is out of range */
int32_t pbt_ostree_take(int32_t* tree, int32_t n, int32_t k);

#ifdef __cplusplus
}
#endif
#endif /* PBT_H_ */
