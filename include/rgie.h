/* rgie.h -- C ABI of the B200-native (sm_100a) hot path of regressor-guided image editing.
 *
 * The reference (christophgebhardt/regressor-guided-image-editing) is pure Python/PyTorch and has NO plugin / FFI
 * layer: the drop-in boundary is its Python call surface.  This header is the C ABI that the Python
 * mirror of that surface (regressor_guided_image_editing_b200/) binds through ctypes; every entry point names the
 * reference function (path relative to /root/reference/src) whose arithmetic it replaces.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all `float*` / `int*` data pointers are DEVICE pointers unless the
 *     parameter name starts with `h_` (host).  Images are NCHW fp32 contiguous.
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*), never synchronises, and is CUDA-graph
 *     capturable (creation / destruction functions excepted: they allocate and may synchronise).
 *   - return value: RGIE_OK (0) or RGIE_ERR (1); rgie_last_error() returns the thread-local message.  Nothing throws
 *     across the ABI.  There is no CPU fallback: without a CUDA device the calls fail.
 *   - the caller owns all buffers it passes; the library owns only what rgie_*_create returns (packed weights,
 *     tap tables, activation workspace), released by the matching rgie_*_destroy.
 */
#ifndef RGIE_H_
#define RGIE_H_

#ifdef __cplusplus
extern "C" {
#endif

#define RGIE_OK 0
#define RGIE_ERR 1
#define RGIE_ABI_VERSION 1

int rgie_version(void);
const char* rgie_last_error(void);

/* ------------------------------------------------------------------------------------------------------------
 * Parametric filters: baselines/image_transformations/image_transformations.py:7-66 (`apply_params` stages) and the
 * functions they call (img_trans_torch_diff.py:6-19,60-64; kornia.enhance.adjust_saturation,
 * adjust_contrast_with_mean_subtraction, sharpness; kornia.filters.gaussian_blur2d; kornia.geometry.transform.scale).
 * Each stage includes the trailing clamp(0,1) of apply_params (:60).  `p` holds EFFECTIVE parameter values (after the
 * reference's clamps in image_transformations.py:98,120,195 and optimize_image_param.py:276-291), `p_stride` floats
 * between images (0 = one parameter set shared by the batch, as in the reference's B=1 semantics).
 * ------------------------------------------------------------------------------------------------------------ */
enum RgieFilter {
  RGIE_F_EXPOSURE = 0,   /* 1 param  */
  RGIE_F_SATURATION = 1, /* 1 param  */
  RGIE_F_TONE = 2,       /* 8 params */
  RGIE_F_COLOR = 3,      /* 24 params (3 x 8) */
  RGIE_F_CONTRAST = 4,   /* 1 param  */
  RGIE_F_SHARP = 5,      /* 1 param  */
  RGIE_F_BLUR = 6,       /* 1 param (sigma, 25x25 kernel, reflect) */
  RGIE_F_SCALE = 7,      /* 4 params (sx, sy, cx, cy) */
  RGIE_F_GAMMA = 8,      /* 1 param: clamp(x^gamma, 0, 1)                      image_transformations.py:176-185 */
  RGIE_F_BRIGHT = 9,     /* 1 param: clamp(x + p, 0, 1)                        :136-143 */
  RGIE_F_BW = 10,        /* 1 param: lerp(x, 0.27 r + 0.67 g + 0.06 b, p)      :156-163 */
  RGIE_F_HUE = 11,       /* 1 param: hsv hue shift, fmod(h + p, 2 pi)          :166-173 */
  RGIE_F_WB = 12,        /* 1 param: lerp(x, x * 0.5 / mean_HW(x), p)          :126-133 */
  RGIE_F_AFFINE = 13     /* 6 params: 2x3 matrix, bilinear warp, border padding :198-206 (d(image) by atomic scatter) */
};
int rgie_filter_param_count(int kind);
/* floats of scratch `ws` that rgie_filter_fwd / rgie_filter_bwd need for a [B,3,H,W] batch */
long rgie_filter_ws_floats(int B, int H, int W);
int rgie_filter_fwd(int kind, const float* in, float* out, const float* p, int p_stride, int B, int H, int W,
                    float* ws, void* stream);
/* gin [B,3,H,W] = d(in);  gp[b*gp_stride + i] = d(param i of image b)  (overwritten, deterministic reduction) */
int rgie_filter_bwd(int kind, const float* in, const float* gout, float* gin, const float* p, int p_stride,
                    float* gp, int gp_stride, int B, int H, int W, float* ws, void* stream);

/* The head of the reference's default filter list (optimize_image_param.py:227: exposure, saturation, tone, color --
 * image_transformations.py:16-58 applied in that order, each followed by clamp(0,1) :60) as ONE pixel pass each way: the
 * three intermediate images are never materialised; results are bit-identical to four rgie_filter_fwd calls.
 * `p` points at image 0's exposure value, followed by saturation, 8 tone and 24 colour values (34 floats, p_stride floats
 * between images).  bwd produces the 34 parameter gradients only: the input is the fixed original image (no d(in)). */
int rgie_filter_prefix_fwd(const float* in, float* out, const float* p, int p_stride, int B, int H, int W, void* stream);
int rgie_filter_prefix_bwd(const float* in, const float* gout, const float* p, int p_stride, float* gp, int gp_stride,
                           int B, int H, int W, float* ws, void* stream);

/* x (raw optimisation vector, optimize_image_param.py:262-292 `get_params_from_vector`) -> effective parameters for
 * the default filter list ['exposure','saturation','tone','color','contrast','sharp','blur','scale'] (41 floats):
 * saturation/sharp/blur clamp(min=0), scale clamp(min=1), centre clamp(0,input_size), contrast<0 -> 0.
 * bwd multiplies gp by the clamp masks in place (contrast<0 -> zero gradient: the reference substitutes a Python float) */
int rgie_params_default_fwd(const float* x, float* p, int B, float input_size, void* stream);
int rgie_params_default_bwd(const float* x, float* gp, int B, float input_size, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Antialiased bilinear resize: torchvision `transforms.Resize(480, antialias=True)` on a tensor
 * (baselines/models/EmotionPredictionModel.py:36-37) = ATen _upsample_bilinear2d_aa (+ its transpose for backward).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct RgieResize RgieResize;
int rgie_resize_create(int in_h, int in_w, int out_h, int out_w, RgieResize** out);
void rgie_resize_destroy(RgieResize* r);
/* planes = B*C; tmp: planes*in_h*out_w floats of scratch */
int rgie_resize_fwd(const RgieResize* r, const float* in, float* out, int planes, float* tmp, void* stream);
int rgie_resize_bwd(const RgieResize* r, const float* gout, float* gin, int planes, float* tmp, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Valence/arousal regressor: torchvision resnet50 (eval, BN folded) behind ReplicateAndCrop / MeanReplicatedCrops
 * (baselines/models/EmotionPredictionModel.py:10-54, utilities/ReplicateAndCrop.py:30-45,
 * utilities/MeanReplicatedCrops.py:18-27).  Forward + input-gradient backward only (no weight gradients: the
 * reference computes and discards them).
 *
 * h_tensors: host fp32 arrays in this order -- conv1.w[64,3,7,7], conv1.b[64]; then per bottleneck (layer1.0 ...
 * layer4.2): c1.w[Cm,Ci,1,1], c1.b, c2.w[Cm,Cm,3,3], c2.b, c3.w[Co,Cm,1,1], c3.b, and for the first block of each
 * layer ds.w[Co,Ci,1,1], ds.b; finally fc.w[num_classes,2048], fc.b.  Conv weights/biases have eval-mode BatchNorm
 * already folded in (w*gamma/sqrt(var+eps), beta-mean*gamma/sqrt(var+eps)).  n_tensors must be 2 + 2*(3*16+4) + 2.
 * precision: RGIE_PREC_FP32 (parity mode: fp32 storage, fp32-accurate GEMMs on the tcgen05 tensor cores through an exact
 *            bf16x3 operand split, csrc/gemm_tc32.cu) | RGIE_PREC_BF16 (tcgen05/TMEM/TMA, throughput mode) |
 *            RGIE_PREC_BF16_SIMT (bf16 storage on CUDA cores: on-device cross-check of the tcgen05 kernels) |
 *            RGIE_PREC_FP32_SIMT (fp32 on CUDA cores: cross-check of the fp32 tensor-core mode).
 * ------------------------------------------------------------------------------------------------------------ */
enum RgiePrecision { RGIE_PREC_FP32 = 0, RGIE_PREC_BF16 = 1, RGIE_PREC_BF16_SIMT = 2, RGIE_PREC_FP32_SIMT = 3 };
typedef struct RgieRegressor RgieRegressor;
int rgie_regressor_create(const float* const* h_tensors, int n_tensors, int num_classes, int crop_size, int max_crops,
                          int precision, RgieRegressor** out);
void rgie_regressor_destroy(RgieRegressor* r);
long rgie_regressor_workspace_bytes(const RgieRegressor* r);
/* img: [B,3,Hr,Wr] (already resized); offsets: int32 [B,reps,2] (top,left) -- the crop draws of
 * torchvision RandomCrop.get_params replayed by the caller; normalize: (x-0.5)/0.5 per crop.
 * logits: [B*reps, num_classes] fp32.  Activations needed by backward stay in the handle's workspace. */
int rgie_regressor_forward(RgieRegressor* r, const float* img, int B, int Hr, int Wr, const int* offsets, int reps,
                           int normalize, float* logits, void* stream);
/* same, with the crop offsets taken from a device-resident table: offsets + (*step_ptr) * off_step_stride
 * (lets a captured CUDA graph replay with a different crop draw every optimisation step) */
int rgie_regressor_forward_ex(RgieRegressor* r, const float* img, int B, int Hr, int Wr, const int* offsets,
                              const int* step_ptr, long off_step_stride, int reps, int normalize, float* logits,
                              void* stream);
/* dlogits: [B*reps, num_classes]; dimg: [B,3,Hr,Wr] gradient w.r.t. `img` of the preceding forward (overwritten).
 * LIFETIME: backward re-reads `offsets` (and `step_ptr`; `img` in input-transform mode 2) of the preceding forward call
 * through the pointers that call was given -- they must stay valid and unchanged until backward has run. */
int rgie_regressor_backward(RgieRegressor* r, const float* dlogits, float* dimg, void* stream);
/* `normalize` of the forward calls: 0 = crops as they are, 1 = (v - 0.5) / 0.5 (ReplicateAndCrop.py:26-28), 2 = the
 * handle's input transform: t = clamp(v * pre_scale + pre_shift, 0, 1); (t - mean_c) / std_c  -- the EmoNet ten-crop
 * pipeline (src/baselines/models/EmoNet.py:63-88: denorm, /255, ImageNet normalisation).  In mode 2 the backward pass
 * re-reads `img` of the forward call (clamp mask), which must still be alive.  A crop whose `left` offset has bit 30 set
 * is mirrored horizontally (EmoNet's flipped crops). */
int rgie_regressor_set_input_transform(RgieRegressor* r, float pre_scale, float pre_shift, const float* mean3,
                                       const float* std3);
/* per-GEMM timing of the last forward/backward (cudaEvent pairs around every row-shifted GEMM launch; used by bench.py
 * for the live roofline figure).  get_profile synchronises on the recorded events.  h_info[4*i..]: {0 fwd | 1 bwd,
 * Cout, K, m_tiles}; h_flops: algorithmic FLOPs (valid pixels only, padding excluded); h_bytes: algorithmic HBM bytes
 * (every distinct operand element read once, every output element written once). */
int rgie_regressor_set_profiling(RgieRegressor* r, int on);
int rgie_regressor_num_ops(const RgieRegressor* r);
int rgie_regressor_get_profile(RgieRegressor* r, float* h_ms, double* h_flops, double* h_bytes, int* h_info, int capacity,
                               int* n_out);
/* number of kernel launches this library has issued in this process (bench.py's gpu_launches evidence) */
long rgie_launch_count(void);

/* debugging / parity taps: copies a named activation of the last forward as fp32 NCHW into `out` (device).
 * names: "stem","pool","layer{1..4}.{i}","layer{1..4}.{i}.c1","...c2","feat".  Returns element count via *n. */
int rgie_regressor_tap(RgieRegressor* r, const char* name, float* out, long capacity, long* n, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Loss head: MeanReplicatedCrops -> Sigmoid -> [:, output_ixs] -> squared error to target
 * (utilities/MeanReplicatedCrops.py:18-27; baselines/losses/ValenceArousalLoss.py:59-73,114-129).  Each image is its own
 * problem (reference batch size 1): loss[b] = scale * sum_{k in mask} (target[b,k]-pred[b,k])^2.
 * preds [B,num_classes]; target [B,2] (valence,arousal) or NULL with (tv_default, ta_default) untargeted values
 * (:82-109); use_mask bit0 = valence, bit1 = arousal; sigmoid: 0/1; dlogits [B*reps,num_classes] (may be NULL).
 * ------------------------------------------------------------------------------------------------------------ */
int rgie_va_head(const float* logits, int B, int reps, int num_classes, int sigmoid, const float* target,
                 float tv_default, float ta_default, int use_mask, float scale, float* preds, float* loss,
                 float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Update: torch.optim.Adam single-tensor step as used by baselines/optimize_image.py:56-97 (betas (0.9,0.999),
 * eps 1e-8), one independent problem per row of x [B,n], with the reference's best-x tracking done on device:
 * if loss[b] < best_loss[b] (strict, :78) then best_loss[b]=loss[b], best_x[b]=x[b] (BEFORE the update), then the Adam
 * update with host-computed lr / bias corrections (float64 on the host as torch does, passed as step_size = lr/bc1 and
 * bc2_sqrt; one_minus_beta{1,2} = (float)(1.0 - beta) evaluated in double like Python does).  loss/best_* may be NULL.
 * ------------------------------------------------------------------------------------------------------------ */
int rgie_adam_step(float* x, const float* g, float* m, float* v, int B, int n, float step_size, float bc2_sqrt,
                   float one_minus_beta1, float beta2, float one_minus_beta2, float eps, const float* loss,
                   float* best_loss, float* best_x, int* best_step, int step, void* stream);

/* CUDA-graph-replayable form: sched is a device table [num_steps,2] of (step_size, bc2_sqrt) indexed by the device
 * counter *step_ptr (the learning-rate ramp of optimize_image.py:69-75 is evaluated on the host in float64 once). */
int rgie_adam_step_sched(float* x, const float* g, float* m, float* v, int B, int n, const float* sched,
                        const int* step_ptr, float one_minus_beta1, float beta2, float one_minus_beta2, float eps,
                        const float* loss, float* best_loss, float* best_x, int* best_step, void* stream);
/* table[(*step_ptr)*n + i] = src[i]  (per-step loss / prediction log kept on the device; the reference formats
 * float(loss) on the host every step, optimize_image.py:89-92) */
int rgie_record(const float* src, float* table, const int* step_ptr, int n, void* stream);
int rgie_counter_add(int* counter, int delta, void* stream);

/* Regressor-guidance update of pipelines/InversionResamplingStableDiffusionPipeline.py:134-142:
 * g /= (||g||_2 + 1e-10) over the WHOLE tensor of each problem (per_problem elements each), x -= scale * g.
 * ws: 2*n_problems*128 floats of scratch. */
int rgie_guidance_update(float* x, const float* g, int n_problems, long per_problem, float scale, int normalize,
                         float* ws, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * MiDU guidance head (SD variant): guidance_classifier/MiduClassifier.py:145-160
 * Conv(1280->256,3,p1) ReLU MaxPool2 Conv(256->128,3,p1) ReLU AdaptiveAvgPool(2,2) Flatten Linear(512,64) ReLU
 * Linear(64,n_out) and the score of guidance_scores.py:4-22.  feat [B,1280,8,8] fp32 NCHW -> pred [B,n_out];
 * backward returns d(score)/d(feat) so the caller's autograd continues into its UNet.
 * h_tensors: 0.w,0.b,3.w,3.b,7.w,7.b,9.w,9.b (host fp32, PyTorch layouts).
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct RgieMiduHead RgieMiduHead;
int rgie_midu_create(const float* const* h_tensors, int n_tensors, int n_out, int max_batch, int hw, int precision,
                     RgieMiduHead** out);
void rgie_midu_destroy(RgieMiduHead* h);
int rgie_midu_forward(RgieMiduHead* h, const float* feat, int B, float* pred, void* stream);
int rgie_midu_backward(RgieMiduHead* h, const float* dpred, float* dfeat, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Self-test hook for the two GEMM backends (tests only): runs the row-shifted GEMM
 *   D[m,n] = sum_t sum_c A[m+row_off[t], c] * W[n, t*Cin + c] (+bias) (+res) (relu)
 * on bf16 device buffers through backend 0 (CUDA-core) or 1 (tcgen05) into an fp32 or bf16 D.
 * ------------------------------------------------------------------------------------------------------------ */
int rgie_gemm_selftest(int backend, const void* A, long a_rows, int Cin, const void* W, int n_pad, int ntaps,
                       const long* h_row_off, long m_begin, long m_end, int Cout, const float* bias, const void* res,
                       int relu, void* D, int d_fp32, void* stream);
/* extended form: optional second operand A2 [a2_rows, Cin2] whose contraction is concatenated along K
 * (W is then [n_pad, ntaps*Cin + Cin2]), optional ReLU mask as 1 bit per element (mask_bits[m, Cout/32]) and optional
 * sign-bit output (D_bits[m, Cout/32], bit set where the stored value is > 0). */
int rgie_gemm_selftest_ex(int backend, const void* A, long a_rows, int Cin, const void* A2, long a2_rows, int Cin2,
                          const void* W, int n_pad, int ntaps, const long* h_row_off, long m_begin, long m_end, int Cout,
                          const float* bias, const void* res, const unsigned* mask_bits, int relu, void* D, int d_fp32,
                          unsigned* D_bits, void* stream);
/* fp32 form (tests only): fp32 device buffers A [a_rows, a_ld or Cin], A2 [a2_rows, Cin2], res / mask [m, Cout], D [m, Cout];
 * W is a HOST fp32 matrix [Cout, ntaps*Cin + Cin2].  backend 0 = CUDA cores (gemm_simt), 2 = tcgen05 with the exact bf16x3
 * operand split (gemm_tc32: the weights are split and uploaded inside the call).  Synchronises the stream before returning. */
int rgie_gemm_selftest_fp32(int backend, const float* A, long a_rows, int Cin, int a_ld, const float* A2, long a2_rows,
                            int Cin2, const float* h_W, int ntaps, const long* h_row_off, long m_begin, long m_end, int Cout,
                            const float* bias, const float* res, const float* mask, int relu, float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RGIE_H_ */
