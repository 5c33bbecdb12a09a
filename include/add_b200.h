/*
 * add_b200.h — C ABI of libadd_b200.so, the B200 (sm_100a) kernels behind the ADD
 * dense-segmentation forward path.
 *
 * The reference (HankKung/Auto-Dynamic-DeepLab) has no native layer: its "FFI" for this path is
 * the set of ATen/cuDNN calls its Python modules dispatch to.  Each entry point below names the
 * reference call site(s) it replaces (file:line under /root/reference).  The Python drop-in
 * modules in auto-dynamic-deeplab_b200/ bind these symbols with ctypes.
 *
 * Conventions
 *  - Plain C: device pointers, ints, a cudaStream_t passed as void*.  No torch types.
 *  - The caller owns every buffer (inputs, outputs, workspaces).  The library never allocates,
 *    never synchronises, and is re-entrant; work is enqueued on `stream`.
 *  - Activations are NHWC ("channels_last") views described by add_tensor_t; a view may be a
 *    channel slice of a wider buffer (pix_stride > c), which is how concat / node-sum are fused.
 *  - Return value: ADD_OK (0) or a negative add_status_t.  add_status_string() explains it.
 *  - There is no CPU fallback: without a CUDA device every launch returns ADD_ERR_CUDA.
 */
#ifndef ADD_B200_H_
#define ADD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ADD_OK = 0,
  ADD_ERR_BAD_ARG = -1,      /* null pointer, negative size, bad enum                     */
  ADD_ERR_UNSUPPORTED = -2,  /* shape / dtype / alignment the kernels do not implement    */
  ADD_ERR_CUDA = -3,         /* cudaGetLastError() after launch was not cudaSuccess       */
  ADD_ERR_WORKSPACE = -4     /* caller-provided workspace too small                       */
} add_status_t;

typedef enum { ADD_F32 = 0, ADD_BF16 = 1 } add_dtype_t;

/* NHWC activation view.  Element (n,y,x,ch) lives at
 *   ptr + ((n*h + y)*w + x) * pix_stride + ch          (in elements of `dtype`)
 * so a channel slice of a wider concat buffer is {ptr = base + c_off, c, pix_stride = C_total}. */
typedef struct {
  void*   ptr;
  int32_t n, h, w, c;
  int32_t pix_stride;
  int32_t dtype;             /* add_dtype_t */
} add_tensor_t;

/* op flags */
#define ADD_RELU_IN    1u    /* apply ReLU to the input as it is loaded (ReLU→conv order)         */
#define ADD_RELU_OUT   2u    /* apply ReLU to the result before it is stored                      */
#define ADD_ACCUMULATE 4u    /* y += result instead of y = result (cell node sum, ADD.py:108)     */

const char* add_status_string(int status);
const char* add_last_cuda_error(void); /* text of the last CUDA error behind an ADD_ERR_CUDA on this thread */
int  add_version(void);               /* 10000*major + 100*minor + patch                       */
int  add_device_sm_count(void);       /* SMs of the current device, <0 on error                */
/* 1 (default): the tcgen05 kernels are launched with programmatic dependent launch — their prologue overlaps the
 * previous kernel's tail and they wait for it (griddepcontrol.wait) before touching activations; 0: plain launches. */
int  add_set_pdl(int on);
/* Tuning: persistent kernels launch pct % of their default CTA count (10..100; default 100); 1000 * c + p scales the
 * small convs of the persistent conv kernel by c % and every other persistent kernel by p %. */
int  add_set_persistent_grid_pct(int pct);

/* ---- layout / dtype edges --------------------------------------------------------------- */

/* NCHW fp32 (the reference's tensor layout at its Python API) → NHWC view (fp32 or bf16).
 * Channels c_src..y->c-1 of the destination are zero-filled (stem0 pads 3 → 4 channels). */
int add_nchw_to_nhwc(const float* src, int c_src, const add_tensor_t* y, void* stream);
/* NHWC view → NCHW fp32. */
int add_nhwc_to_nchw(const add_tensor_t* x, float* dst, void* stream);

/* ---- dense convolution (CUDA-core fp32 accumulate; any dtype mix) ------------------------ */
/* Replaces nn.Conv2d + folded eval-mode BatchNorm (+ReLU) call sites: ReLUConvBN
 * (operations.py:18-29), DilConv (operations.py:32-43), FactorizedReduce / DoubleFactorizedReduce
 * (operations.py:86-119, as two launches with pad = 0 and pad = -stride/2 into channel halves),
 * stems (ADD.py:154-169), low_level_conv (ADD.py:255-259), ASPP_train branches
 * (aspp_train.py:16-25), Decoder._conv (decoder.py:12-21), EDM.conv (ADD.py:508).
 *   y[n,oy,ox,co] (+)= act( bias[co] + sum_{ky,kx,ci} w[ky][kx][ci][co] *
 *                           relu?(x[n, oy*stride - pad + ky*dil, ox*stride - pad + kx*dil, ci]) )
 * w: fp32, layout [kh][kw][Cin][Cout] with the BN scale already folded in; bias: fp32[Cout] or NULL.
 * `pad` may be negative (FactorizedReduce's odd lattice).  Out-of-range taps read zero. */
int add_conv2d_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w, const float* bias,
                   int64_t bias_image_stride, int kh, int kw, int stride, int pad, int dil, uint32_t flags,
                   void* stream);
/* bias_image_stride: 0 = one bias vector; else image n uses bias + n*bias_image_stride (floats) — how the ASPP
 * image-pool branch (constant over an image) enters the 1x1 over the concatenation (aspp_train.py:49-59). */

/* ---- dense convolution on tcgen05 tensor cores (bf16 in, fp32 TMEM accumulate) ----------- */
/* Same contract as add_conv2d_fwd for bf16 activations; weights are pre-packed by
 * add_conv2d_tc_pack() into the UMMA shared-memory image the kernel streams with bulk copies. */
int64_t add_conv2d_tc_packed_bytes(int cin, int cout, int kh, int kw);
int add_conv2d_tc_pack(const float* w_hwio, int cin, int cout, int kh, int kw, void* packed_host);
int add_conv2d_tc_fwd(const add_tensor_t* x, const add_tensor_t* y, const void* w_packed,
                      const float* bias, int64_t bias_image_stride, int kh, int kw, int stride, int pad, int dil,
                      uint32_t flags, void* stream);

/* ---- first layer on the bf16 path: NCHW fp32 image -> 3x3 stride-2 conv 3->64 + folded BN (+ReLU) -> NHWC
 * bf16 in one kernel (ADD.py:154-158 `stem0` applied to the loader's NCHW tensor, eval.py:175): the layout
 * change, bf16 conversion and im2col happen on the way into shared memory; tcgen05 GEMM; TMA tile store.
 * w_packed: add_stem_tc_pack() of the fp32 [64][3][3][3] weights (BN scale folded); y: [n,(h-1)/2+1,(w-1)/2+1,64]. */
int64_t add_stem_tc_packed_bytes(void);
int add_stem_tc_pack(const float* w_oihw, void* packed_host);
int add_stem_conv3x3s2_nchw_fwd(const float* x_nchw, int n, int h, int w, const add_tensor_t* y,
                                const void* w_packed, const float* bias, uint32_t flags, void* stream);

/* A-operand strategy of add_conv2d_tc_fwd: 0 = one TMA tile per tap; 1 = halo rows resident in shared
 * memory, taps addressed through shifted UMMA descriptors (default); 2 = as 1 with descriptor base_offset. */
int add_conv2d_tc_set_halo_mode(int mode);

/* ---- SepConv half: ReLU → depthwise k×k → pointwise 1×1 → folded BN ---------------------- */
/* Replaces operations.py:51-54 and :55-58 (two calls make one SepConv, operations.py:46-62).
 * w_dw: fp32 [k][k][C]; w_pw: fp32 [Cin][Cout] (BN scale folded); bias fp32[Cout]. Stride 1, pad k/2. */
int add_sepconv_half_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw,
                         const float* w_pw, const float* bias, int k, uint32_t flags, void* stream);

/* Same contract on the tensor-core path (bf16 activations): CUDA-core depthwise (packed fp32x2 FMA,
 * TMA halo tile) whose bf16-rounded result is the A operand of a tcgen05 pointwise GEMM with a TMEM
 * accumulator.  w_pw_packed: add_conv2d_tc_pack() image of the [1][1][Cin][Cout] pointwise weights.
 * C % 8 == 0, C <= 256, Cout <= 256; y may be bf16 or fp32. */
int add_sepconv_half_tc_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw,
                            const void* w_pw_packed, const float* bias, int k, uint32_t flags,
                            void* stream);

/* Stand-alone depthwise k x k (stride 1, pad k/2; `nn.Conv2d(C, C, k, groups=C)`, operations.py:52,56), w_dw fp32
 * [k][k][C], ReLU-on-load / ReLU-on-store flags: for SepConv halves wider than add_sepconv_half_tc_fwd takes (C > 256);
 * the pointwise 1x1 + BN follows as an add_conv2d_tc_fwd over the bf16 result. */
int add_depthwise_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w_dw, int k, uint32_t flags, void* stream);

/* add_sepconv_half_tc_fwd scheduling: 1 = persistent warp-specialised pipeline (default), 0 = one tile per CTA. */
int add_sepconv_tc_set_mode(int mode);

/* ---- the non-convolutional OPS primitives (operations.py:7-11) ------------------------------------------ */
/* avg_pool_3x3 = nn.AvgPool2d(3, stride, padding=1, count_include_pad=False) (mode 0: divisor = taps inside the image);
 * max_pool_3x3 = nn.MaxPool2d(3, stride, padding=1) (mode 1).  y: [n, (h-1)/stride+1, (w-1)/stride+1, c]. */
int add_pool3x3_fwd(const add_tensor_t* x, const add_tensor_t* y, int mode, int stride, uint32_t flags, void* stream);
/* y (+)= scale * x[:, ::stride, ::stride, :] — skip_connect / Identity (scale 1, stride 1, operations.py:65-71) and
 * none / Zero (scale 0: IEEE x*0 like the reference's x.mul(0.), operations.py:74-83). */
int add_scale_fwd(const add_tensor_t* x, const add_tensor_t* y, float scale, int stride, uint32_t flags, void* stream);

/* ---- training-mode BatchNorm forward -------------------------------------------------------
 * The reference's SynchronizedBatchNorm (modeling/sync_batchnorm/batchnorm.py:48-78, 113-125) splits a training
 * forward into per-device [sum, square-sum] (:59-61), one reduce + broadcast over the devices (:90-111), the
 * mean / inv_std / running-statistics update (:113-125) and the normalisation (:68-75).  Same split here; the reduce is
 * ONE all-reduce of the packed [sum(C) | ssum(C) | count] vector done by the host side with torch.distributed.
 *
 * add_bn_stats_fwd: sums[0..C) = per-channel sum of x over (n,h,w), sums[C..2C) = sum of x^2 (fp32, deterministic
 *   two-stage reduction).  workspace >= add_bn_stats_workspace_bytes(n,h,w,c).  C % 4 == 0, C <= 1024.
 * add_bn_finalize: from (possibly all-reduced) sums and the element count (count_dev: device pointer to the fp32 count,
 *   e.g. the all-reduced sums + 2C; NULL -> `count`): mean = sum/n, sumvar = ssum - sum*mean,
 *   running_mean = (1-m) running_mean + m mean, running_var = (1-m) running_var + m sumvar/(n-1)   (either may be NULL),
 *   inv_std = clamp(sumvar/n, eps)^-1/2 when sync != 0 (batchnorm.py:125), (sumvar/n + eps)^-1/2 when sync == 0
 *   (F.batch_norm, the :50-53 path the reference takes on one device or under DDP).
 * add_bn_apply_fwd: y = (x - mean) * (inv_std * weight) + bias [ReLU with ADD_RELU_OUT]; weight / bias may be NULL
 *   (affine=False); with mean = running_mean and inv_std = (running_var + eps)^-1/2 it is eval-mode F.batch_norm. */
int64_t add_bn_stats_workspace_bytes(int n, int h, int w, int c);
int add_bn_stats_fwd(const add_tensor_t* x, float* sums, void* workspace, int64_t workspace_bytes, void* stream);
int add_bn_finalize(const float* sums, const float* count_dev, float count, int c, float eps, float momentum, int sync,
                    float* running_mean, float* running_var, float* mean, float* inv_std, void* stream);
int add_bn_apply_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* mean, const float* inv_std,
                     const float* weight, const float* bias, uint32_t flags, void* stream);

/* ---- bilinear resize, align_corners=False (F.interpolate: ADD.py:76,84,89,317; decoder.py:24) */
int add_bilinear_fwd(const add_tensor_t* x, const add_tensor_t* y, uint32_t flags, void* stream);
/* add_bilinear_fwd scheduling: 1 = bf16 upscales by >= 1.5x use a kernel whose threads own a SOURCE cell and produce
 * every output that interpolates it (default; results identical to the generic kernel), 0 = generic kernel everywhere. */
int add_bilinear_set_mode(int mode);

/* ---- batch compaction for per-image early exit (ADD.py:421-432 applied to a batch): whole-image
 * slabs dst[j] = src[idx[j]], j < count; idx lives on the device so the launch is graph-replayable. */
int add_gather_images(const void* src, void* dst, const int32_t* idx_dev, int count,
                      int64_t bytes_per_image, void* stream);
/* Same for NHWC views: dst image j = src image idx[j] restricted to the views' channel range (dst->n images are
 * written; src and dst may be channel slices of wider buffers with different pixel strides). */
int add_gather_images_view(const add_tensor_t* src, const add_tensor_t* dst, const int32_t* idx_dev, void* stream);

/* ---- global average pool (aspp_train.py:49, ADD.py:522): out[n][c] fp32 = mean_hw relu?(x).
 * Two deterministic stages (per-split partial sums in the workspace, fixed-order final sum). */
int64_t add_global_avgpool_workspace_bytes(int n, int h, int w, int c);
int add_global_avgpool_fwd(const add_tensor_t* x, float* out, uint32_t flags, void* workspace,
                           int64_t workspace_bytes, void* stream);

/* ---- ASPP image-pool branch as a per-image bias (aspp_train.py:49-59): GAP -> 1x1 (+BN, ReLU) -> align_corners
 * upsample from 1x1 (a broadcast) -> its 256 channels of the 1280->256 1x1.  Constant over the image, so
 *   bias_out[n][co] = b_out[co] + sum_d w_out_pool[d][co] * relu(b5[d] + sum_ci w5[ci][d] * pooled[n][ci])
 * (fp32; w5 [cin][depth], w_out_pool [depth][cout] = the pool-branch rows of the folded 1x1) feeds
 * add_conv2d*_fwd(bias_image_stride = cout) over the other four branches. */
int add_aspp_pool_bias_fwd(const float* pooled, int n, int cin, const float* w5, const float* b5, int depth,
                           const float* w_out_pool, const float* b_out, int cout, float* bias_out, void* stream);

/* ---- EDM tail (ADD.py:509-513,523-525): pooled[n][128] → Linear/ReLU ×2 → Linear → out[n] */
int add_edm_mlp_fwd(const float* pooled, int n, const float* w0, const float* b0, const float* w1,
                    const float* b1, const float* w2, const float* b2, float* out, void* stream);

/* ---- exit head: final bilinear (decoder.py:28) fused with its consumers -------------------- */
/* logits view x: [n,h,w,num_class] fp32 NHWC at decoder resolution; (H,W) = input image size.   */
/* (a) materialise the reference's return value: NCHW fp32 [n,num_class,H,W]                    */
int add_upsample_logits_nchw(const add_tensor_t* x, float* dst, int H, int W, void* stream);
/* (b) upsample → argmax (eval.py:183) → optional int64 prediction map and/or confusion matrix
 *     (utils/metrics.py:34-39).  gt: int64 [n,H,W] or NULL; pred_out: int64 [n,H,W] or NULL;
 *     cm_out: int64 [n][num_class*num_class] per-image matrices or NULL; cm_row_index: device int32 [n] or NULL —
 *     when given, image j's matrix is written to row cm_row_index[j] of cm_out (the early-exit runner's compacted
 *     batches scatter their results straight into the [N][nc*nc] result of the original batch);
 *     entropy_out: float [n] per-image normalized Shannon entropy (operations.py:161-170) or NULL.
 *     workspace: add_head_workspace_bytes() bytes.                                              */
int64_t add_head_workspace_bytes(int n, int H, int W, int num_class);
int add_upsample_argmax_fwd(const add_tensor_t* x, int H, int W, const int64_t* gt,
                            int64_t* pred_out, int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out,
                            void* workspace, int64_t workspace_bytes, void* stream);
/* Same with uint8 labels (the Cityscapes PNG bytes; 255 = ignore, like any value >= num_class): 1 byte per pixel
 * instead of the 8 of the int64 tensor the reference's Evaluator receives. */
int add_upsample_argmax_u8_fwd(const add_tensor_t* x, int H, int W, const uint8_t* gt_u8, int64_t* pred_out,
                               int64_t* cm_out, const int32_t* cm_row_index, float* entropy_out, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- loader / dump edges ------------------------------------------------------------------ */
/* Cityscapes label ids -> train ids (dataloaders/datasets/cityscapes.py:85-91: void ids -> 255, valid ids -> 0..18) as a
 * 256-entry device table, fused with the bottom / right pad of the evaluation transform (custom_transforms.py:344,
 * ConstantPad2d(..., 255)): dst[n][Hp][Wp] = lut[src[n][h][w]] inside the image, `fill` outside.  lut256_dev NULL =
 * identity (pad only). */
int add_encode_pad_labels_u8(const uint8_t* src, uint8_t* dst, int n, int h, int w, int Hp, int Wp,
                             const uint8_t* lut256_dev, int fill, void* stream);
/* full_image_eval_preprocess (custom_transforms.py:322-347): uint8 HWC [n][h][w][3] -> torchvision ToTensor + Normalize
 * (all float32: /255, - mean, / std, each rounded to nearest; bit-identical) -> zero pad to [n][3][Hp][Wp] fp32 (ZeroPad2d
 * AFTER the normalisation, so the padding is exactly 0).  (add_normalize_u8_hwc_to_nchw is the numpy Normalize class of
 * the training transforms, which evaluates - mean and / std in float64.) */
int add_normalize_pad_u8_hwc_to_nchw(const uint8_t* src, float* dst, int n, int h, int w, int Hp, int Wp, double mean0,
                                     double mean1, double mean2, double std0, double std1, double std2, void* stream);
/* decode_segmap (dataloaders/utils.py:14-51): class map (int64 if labels_are_int64 else uint8) -> uint8 RGB [n_pixels][3]
 * through a 256 x 3 device table (rows >= 19: the label value itself in all three channels, like the reference). */
int add_decode_segmap(const void* labels, int labels_are_int64, uint8_t* rgb, int64_t n_pixels, const uint8_t* lut768_dev,
                      void* stream);

/* Loader edge: uint8 HWC images [n][h][w][3] (PIL / Cityscapes PNG layout) -> normalised fp32 NCHW [n][3][h][w], the
 * tensor eval.py:175 copies to the device.  Same arithmetic as the reference's host transforms (Normalize then ToTensor,
 * dataloaders/custom_transforms.py:17-24, :39: /255 in float32, -mean and /std through float64), bit-identical. */
int add_normalize_u8_hwc_to_nchw(const uint8_t* src, float* dst, int n, int h, int w, double mean0, double mean1, double mean2,
                                 double std0, double std1, double std2, void* stream);

/* ---- label edge: uint8 labels (Cityscapes PNG depth, 255 = ignore) -> int64 [n] as the Evaluator path reads them.
 * The loader-side H2D then moves 1 byte per pixel instead of 8 (cityscapes.py:85-91 encodes ids into 0..18 / 255). */
int add_widen_labels_u8(const uint8_t* src, int64_t* dst, int64_t n, void* stream);

/* ---- Evaluator (utils/metrics.py:34-39): int64 confusion matrix, atomics-free ------------ */
int64_t add_confusion_workspace_bytes(int64_t n_pixels, int num_class);
int add_confusion_matrix(const int64_t* gt, const int64_t* pred, int64_t n_pixels, int num_class,
                         int64_t* cm_out, void* workspace, int64_t workspace_bytes, void* stream);
/* A/B switch (no effect on results): 0 = per-warp privatised histogram, intra-warp collisions resolved with match.any,
 * one-block finalize (default: the variant measured on the B200); 1 = thread-private 16-bit histograms in shared memory,
 * conflict-free, one CTA per SM, many-loads-in-flight finalize; 2 = the default histogram kernel with that finalize
 * (1 and 2 opt-in until pinned on hardware). */
int add_confusion_set_impl(int impl);

/* ---- confidence scalars on materialised NCHW fp32 logits (operations.py:161-180) ---------- */
/* out[0] = normalized Shannon entropy summed over batch and pixels / (H*W);
 * out[1] = fraction of pixels whose max softmax prob > threshold (also / (H*W)).               */
int64_t add_confidence_workspace_bytes(int n, int H, int W);
int add_confidence_nchw(const float* logits, int n, int num_class, int H, int W, float threshold,
                        float* out2, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- training step: backward kernels (train.py:216-247) — fp32 NHWC, deterministic ------------------ */
/* dx *= (x > 0): the ReLU in front of every conv (operations.py:21,33,47) */
int add_relu_mask_bwd(const add_tensor_t* x, const add_tensor_t* dx, void* stream);
/* conv weight gradient, layout [kh][kw][Cin][Cout] like add_conv2d_fwd's weights; flags: ADD_RELU_IN (the conv read relu(x)),
 * ADD_ACCUMULATE (dw += ...).  workspace: add_conv2d_wgrad_workspace_bytes(). */
int64_t add_conv2d_wgrad_workspace_bytes(int n, int ho, int wo, int cin, int cout, int kh, int kw);
int add_conv2d_wgrad(const add_tensor_t* x, const add_tensor_t* dy, float* dw, int kh, int kw, int stride, int pad, int dil,
                     uint32_t flags, void* workspace, int64_t workspace_bytes, void* stream);
/* conv input gradient for any stride (stride-1 convs use add_conv2d_fwd with flipped weights instead); flags: ADD_ACCUMULATE */
int add_conv2d_dgrad(const add_tensor_t* dy, const float* w, const add_tensor_t* dx, int kh, int kw, int stride, int pad,
                     int dil, uint32_t flags, void* stream);
/* depthwise weight gradient [k][k][C] (operations.py:52,56) */
int64_t add_depthwise_wgrad_workspace_bytes(int n, int h, int w, int c, int k);
int add_depthwise_wgrad(const add_tensor_t* x, const add_tensor_t* dy, float* dw, int k, uint32_t flags, void* workspace,
                        int64_t workspace_bytes, void* stream);
/* training-mode BatchNorm backward (F.batch_norm / SynchronizedBatchNorm2d, sync_batchnorm/batchnorm.py:59-75,113-125):
 * reduce -> sums double[2][C] = [sum dy', sum dy' * xhat] (dy' masked by y > 0 with ADD_RELU_OUT); the caller all-reduces
 * `sums` over the ranks for the synchronised layer; apply -> dx = gamma * inv_std * (dy' - sum1/M - xhat * sum2/M * var_term).
 * d gamma = sum2, d beta = sum1 (this rank's, before the all-reduce). */
int64_t add_bn_bwd_workspace_bytes(int n, int h, int w, int c);
int add_bn_bwd_reduce(const add_tensor_t* dy, const add_tensor_t* x, const float* mean, const float* inv_std, const float* gamma,
                      const float* beta, uint32_t flags, double* sums, void* workspace, int64_t workspace_bytes, void* stream);
int add_bn_bwd_apply(const add_tensor_t* dy, const add_tensor_t* x, const float* mean, const float* inv_std, const float* gamma,
                     const float* beta, const double* sums, double inv_count, const float* count_dev, const float* var_term,
                     uint32_t flags, const add_tensor_t* dx, void* stream);   /* count_dev (device fp32, or NULL): M read on the device */
/* adjoint of add_bilinear_fwd (F.interpolate bilinear, align_corners=False): per-axis tables built once per (in, out) size */
int add_bilinear_bwd_tables(int in_size, int out_size, int32_t* i0, int32_t* i1, float* l0, float* l1, int32_t* lo, int32_t* hi,
                            void* stream);
int add_bilinear_bwd(const add_tensor_t* dy, const add_tensor_t* dx, const int32_t* yi0, const int32_t* yi1, const float* yl0,
                     const float* yl1, const int32_t* ylo, const int32_t* yhi, const int32_t* xi0, const int32_t* xi1,
                     const float* xl0, const float* xl1, const int32_t* xlo, const int32_t* xhi, uint32_t flags, void* stream);
/* 3x3 pool backward (operations.py:9-10): mode 0 avg (count_include_pad=False), 1 max (first maximum) */
int add_pool3x3_bwd(const add_tensor_t* x, const add_tensor_t* dy, const add_tensor_t* dx, int mode, int stride, uint32_t flags,
                    void* stream);
/* nn.CrossEntropyLoss(weight, ignore_index) on fp32 logits (utils/loss.py:16-25; train.py:229-233).  Element (n, c, pixel) at
 * n*stride_n + c*stride_c + pixel*stride_pix (NCHW or NHWC with channels padded to c_store); loss_wsum[0] = mean loss,
 * [1] = weight sum of the valid pixels; dlogits (or NULL, same strides) = grad_scale * d loss / d logits */
int64_t add_ce_loss_workspace_bytes(int n, int h, int w);
int add_ce_loss_fwd_bwd(const float* logits, const int64_t* target, int n, int num_class, int h, int w, int64_t stride_n,
                        int64_t stride_c, int64_t stride_pix, int c_store, int64_t ignore_index, const float* class_weight,
                        float grad_scale, float* loss_wsum, float* dlogits, void* workspace, int64_t workspace_bytes, void* stream);
/* torch.optim.SGD(momentum, weight_decay, nesterov) (train.py:126-127) over a device table of
 * {float* param, const float* grad, float* momentum_buf, int64 numel} entries, one launch */
int add_sgd_nesterov(const void* table_dev, int n_tensors, int64_t max_numel, float lr, const float* lr_dev, float momentum,
                     float weight_decay, int nesterov, int first_step, void* stream);   /* lr_dev (or NULL): lr read on the device */

/* ---- SynchronizedBatchNorm2d exchange over NVLink peer memory (sync_batchnorm/batchnorm.py:90-111) -------------------------
 * One-shot all-reduce of a short vector in ONE kernel, no NCCL: each rank writes its vector into its symmetric buffer, signals
 * every peer, waits for all peers' signals and sums the ranks' vectors in rank order (bit-identical on all ranks).
 * buf_ptrs / signal_ptrs: HOST arrays of `world` peer-mapped addresses (symmetric buffer = 2 slots of cap_bytes; signal pad =
 * >= world zero-initialised uint32 words); epoch = 1, 2, 3, ... the same on every rank; status_dev: set to 1 when a peer did
 * not arrive within ~2 s (never hangs).  vec is fp32 (elem_bytes 4) or fp64 (8). */
int add_peer_allreduce(void* vec, int n, int elem_bytes, const uint64_t* buf_ptrs, const uint64_t* signal_ptrs, int rank, int world,
                       uint32_t epoch, uint32_t* epoch_dev, int64_t cap_bytes, int* status_dev, void* stream);
/* epoch_dev (or NULL): device call counter incremented by the kernel itself — the launch is then CUDA-graph replayable */

/* ---- supernet edge: MixedOp = sum_k w_k * op_k(x) over the eight PRIMITIVES (cell_level_search.py:10-29) -- */
/* The four parameter-free primitives of an edge in ONE kernel, x read once:
 *   y (+)= w[0] * (x * 0) + w[1] * BN(maxpool3x3(x)) + w[2] * BN(avgpool3x3(x)) + w[3] * x        (PRIMITIVES order)
 * w8_dev: device fp32 [8] (the softmaxed alphas of the edge; entries 4..7 belong to the conv primitives, which add
 * w_k * op_k(x) in their own epilogues); BN(affine=False) after each pool as per-channel (mean, inv_std) device vectors. */
int add_mixed_light_fwd(const add_tensor_t* x, const add_tensor_t* y, const float* w8_dev, const float* max_mean,
                        const float* max_inv_std, const float* avg_mean, const float* avg_inv_std, uint32_t flags, void* stream);
/* y = sum_k w[k] * y_k, the k weights read on the device; ys: HOST array of k descriptors (NULL entry = primitive skipped).
 * Backward: dys[k] (NULL = not needed) = w[k] * dy, dw[k] = <dy, y_k> (deterministic).  fp32. */
int add_weighted_sum_fwd(const add_tensor_t* const* ys, int k, const float* w_dev, const add_tensor_t* out, void* stream);
int64_t add_weighted_sum_workspace_bytes(int n, int h, int w, int c);
int add_weighted_sum_bwd(const add_tensor_t* dy, const add_tensor_t* const* ys, const add_tensor_t* const* dys, int k,
                         const float* w_dev, float* dw_dev, void* workspace, int64_t workspace_bytes, void* stream);
/* softmax over the last dimension of a [rows, k] fp32 matrix, k <= 32 (alphas / betas, model_net_search.py:294-310) */
int add_softmax_rows_fwd(const float* x, float* y, int rows, int k, void* stream);
int add_softmax_rows_bwd(const float* y, const float* dy, float* dx, int rows, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ADD_B200_H_ */
