/*
 * hg_api.h -- C ABI of libhgb200.so: hand-written sm_100a kernels for the stacked-hourglass
 * hot path of minhhoangbui/hourglass-pose-estimation.
 *
 * The reference has NO native interface for this path: everything is Python calling into
 * PyTorch/cuDNN.  The boundary therefore sits one level down: each entry
 * point below replaces the library op(s) a reference call site issues, takes raw DEVICE pointers
 * with explicit shapes, and is bound from the reference-facing Python modules with ctypes
 * (hourglass-pose-estimation_b200/hgb200/_lib.py; stub shown in INTEGRATION.md).
 *
 * Conventions
 *   - every function returns 0 (HG_OK) or a negative HG_ERR_* code; text via hg_last_error().
 *   - never allocates device memory, never synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - the caller (PyTorch) owns every buffer; device = the calling thread's current device.
 *   - activations are NHWC bf16 unless stated; heat maps / images at the API edge are NCHW fp32
 *     exactly as the reference's tensors are.
 *   - thread-safe: no mutable global state except a per-device property cache behind a mutex.
 */
#ifndef HG_API_H_
#define HG_API_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HG_OK 0
#define HG_ERR_INVALID (-1)  /* bad argument / unsupported shape */
#define HG_ERR_CUDA (-2)     /* a CUDA runtime/driver call failed */
#define HG_ERR_ARCH (-3)     /* device is not sm_100 */
#define HG_ERR_DEVICE (-4)   /* a kernel reported a protocol timeout through its error word */

#define HG_API_VERSION 9

int hg_api_version(void);
/* Copies the calling thread's last error text (NUL-terminated) into buf; returns its length. */
size_t hg_last_error(char* buf, size_t cap);
/* 0 if the current device can run the library (compute capability 10.x). */
int hg_check_device(void);
/* Device-side error word: kernels with bounded mbarrier waits write a non-zero code here when a
 * wait times out instead of hanging.  The word lives in caller-owned device memory (4 bytes). */

/* ------------------------------------------------------------------------------------------- *
 * Convolution as implicit GEMM on tcgen05/TMEM (replaces nn.Conv2d + folded nn.BatchNorm2d +
 * nn.ReLU + residual add + nearest-upsample add:  src/models/modules.py:27-47,80-96 and
 * src/models/hourglass.py:60-67,71-73,83-89 of the reference).
 *
 *   out[n,y,x,:] = epi( sum_taps  W[:, tap, :] . pro(in[n, y+dy, x+dx, :])
 *                       + W2 . in2[n,y,x,:] + bias )
 *   pro(v)  = relu(v * in_scale[c] + in_shift[c])          if in_scale != NULL (pre-activation bn1)
 *   epi(a)  = [relu]( a + residual[n,y,x,:] + up_low[n, y/2, x/2, :] )   (each term optional)
 *
 * weight: bf16 [cout_pad][ktot], K-major, ktot = taps*cin + cin2, k = tap*cin + c (tap = ky*3+kx),
 *         then the cin2 columns of the second (1x1) operand; cout_pad = cout rounded up to 16.
 * ------------------------------------------------------------------------------------------- */
typedef struct hg_conv_desc {
    const void* in;          /* bf16 NHWC [n][h][w][cin]                                        */
    const void* in2;         /* optional second 1x1 operand, bf16 NHWC [n][h][w][cin2], or NULL  */
    const void* weight;      /* bf16 [cout_pad][ktot]                                           */
    const float* bias;       /* fp32 [cout_pad]                                                 */
    const float* in_scale;   /* fp32 [cin] or NULL (prologue on `in`; 1x1 only)                 */
    const float* in_shift;   /* fp32 [cin] or NULL                                              */
    const void* residual;    /* bf16 NHWC [n][h][w][cout] or NULL                               */
    const void* up_low;      /* bf16 NHWC [n][h/2][w/2][cout] or NULL (nearest x2 then add)     */
    void* out;               /* bf16 NHWC [n][h][w][cout]      (NULL when out_nchw_f32 is used) */
    float* out_nchw_f32;     /* fp32 NCHW [n][cout][h][w] heat-map output (cout_pad <= 32 only) */
    unsigned int* err_word;  /* device uint32, set non-zero on a kernel protocol timeout         */
    int32_t n, h, w;
    int32_t cin, cin2, cout;
    int32_t ksize;           /* 1 or 3 (stride 1, pad ksize/2)                                  */
    int32_t relu;            /* apply ReLU in the epilogue                                      */
    int32_t out_halo;        /* 1: `out` is a HALO-PADDED buffer (see hg_conv3x3_halo_bf16); 1x1 only, w <= 253,
                                cout in {64,128} unless 128 %% w == 0 and (h*w) %% 128 == 0        */
    float* stats;            /* optional fp32 [2*cout]: per-channel sum | sum of squares over all pixels of the
                                epilogue's result, ADDED (atomics) -- the batch statistics the next train-mode
                                BatchNorm needs (nn.BatchNorm2d in training, src/models/modules.py:30-41), fused
                                into the producing GEMM so that no separate pass re-reads the tensor            */
    void* pool_out;          /* optional second output: F.max_pool2d(out, 2, stride=2) as bf16 NHWC [n][h/2][w/2][cout]
                                (the hourglass pools every level's input, src/models/modules.py:82), written by the
                                same epilogue.  1x1 only: cout 256, cin + cin2 <= 128, no prologue / stats / out_halo,
                                w a power of two <= 64 with 128 %% (2*w) == 0, even h                              */
    void* pool_in;           /* optional second output: F.max_pool2d(in, 2, stride=2) of the RAW input (before the
                                in_scale / in_shift prologue) as bf16 NHWC [n][h/2][w/2][cin].  Hourglass pools the very
                                tensor its `up1` bottleneck opens with (src/models/modules.py:81-83): the 1x1 kernel's
                                prologue warps already hold every input tile in shared memory, so the pool kernel's re-read
                                of the tensor disappears.  1x1 with prologue only: cout 128, cin2 = 0, no stats, w a power
                                of two <= 64 with 128 %% (2*w) == 0, even h (128-pixel tiles hold whole 2x2 windows)   */
} hg_conv_desc;

int hg_conv_nhwc_bf16(const hg_conv_desc* d, void* stream);

/* 3x3 convolution (stride 1, pad 1) + bias (+ReLU) reading a HALO-PADDED input: a bf16 buffer of
 * hg_halo_padded_elems(n,h,w,c) elements laid out as [1 zero row of (w+1) pixels][n][h+1][w+1][c], i.e. one zero
 * column right of every image row and one zero row under every image.  The caller zero-initialises the
 * buffer ONCE; producers (hg_conv_nhwc_bf16 with out_halo=1) only ever write the interior.  In this layout
 * every filter tap is a constant offset, so each input pixel is fetched once per tile instead of nine times.
 * weight: bf16 [cout][9*cin], k = (ky*3+kx)*cin + c; out: dense bf16 NHWC [n][h][w][cout]; cout in {64,128}. */
int64_t hg_halo_padded_elems(int32_t n, int32_t h, int32_t w, int32_t c);
int hg_conv3x3_halo_bf16(const void* in_padded, const void* weight, const float* bias, void* out, unsigned int* err_word,
                         float* stats /* optional fp32 [2*cout], as hg_conv_desc.stats */, int32_t n, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t relu, void* stream);

/* The tail of a bottleneck in ONE launch (reference src/models/modules.py:38-46: conv2 (3x3, 128->128) -> bn3 -> relu ->
 * conv3 (1x1, 128->256) -> `out += residual`, plus Hourglass's `up1 + F.interpolate(low3, scale_factor=2)`,
 * modules.py:90-95, when up_low is given).  The 3x3's bf16 result stays in shared memory as the 1x1's operand; the
 * kernel runs on CTA pairs (cta_group::2).  in_padded / w2 / b2 as hg_conv3x3_halo_bf16 with cin = cout = 128, relu = 1
 * (b2 carries the folded bn3); w3: bf16 [256][128], b3: fp32 [256]; residual: dense bf16 NHWC [n][h][w][256] or NULL;
 * up_low: dense bf16 NHWC [n][h/2][w/2][256] or NULL (h, w even); out: dense bf16 NHWC [n][h][w][256].
 * hg_conv3x3_k3_fusable returns 1 when the size suits the kernel (enough 256-position tiles to fill the CTA pairs,
 * shared-memory budget); otherwise run hg_conv3x3_halo_bf16 followed by hg_conv_nhwc_bf16. */
int hg_conv3x3_k3_fusable(int32_t n, int32_t h, int32_t w);
int hg_conv3x3_k3_fused_bf16(const void* in_padded, const void* w2, const float* b2, const void* w3, const float* b3,
                             const void* residual, const void* up_low, void* out, unsigned int* err_word, int32_t n,
                             int32_t h, int32_t w, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Input preprocessing (the step before the path).  mean3 / std3 are HOST arrays.
 * ------------------------------------------------------------------------------------------- */
/* transforms.ToTensor() + Normalize(mean, std) (src/datasets/common.py:57-64) on uint8 HWC crops [n][h][w][3]:
 * float32 (x/255 - mean[c]) / std[c] with IEEE division (bit-identical to torch).  Writes fp32 NCHW
 * [n][3][h][w] (out_nchw, may be NULL) and/or the stem's packed NHWC4 bf16 staging image [n][h][w+8][4]
 * (packed, may be NULL; interior only, flip_w mirrors) -- hg_stem_pack's layout. */
int hg_normalize_u8_nhwc(const void* in_u8, const float* mean3, const float* std3, float* out_nchw, void* packed,
                         int32_t n, int32_t h, int32_t w, int32_t flip_w, void* stream);
/* Estimator.preprocess_bbox (src/runner/estimator.py:39-54) for n equally sized uint8 HWC frames [n][fh][fw][3]:
 * x/255, (x - mean[c]) / std[c] in float64 (skipped when mean3 == std3 == NULL: the reference's datasets without
 * a branch), cv2.resize(INTER_LINEAR) to (w, h) in float64, cast to float32 NCHW [n][3][h][w]. */
int hg_preprocess_frames_u8(const void* frames_u8, const double* mean3, const double* std3, float* out_nchw, int32_t n,
                            int32_t fh, int32_t fw, int32_t h, int32_t w, void* stream);

/* Stem: conv 7x7 stride 2 pad 3 (3 -> cout) + folded BN + ReLU (src/models/hourglass.py:71-73).
 * Step 1 gathers NCHW fp32 pixels into K-major bf16 rows [n*oh*ow][192] (k = (ky*7+kx)*3 + c,
 * zero padded 147 -> 192); flip_w != 0 reads the image mirrored left-right (flip test).
 * Step 2 is hg_conv_nhwc_bf16 with ksize 1, cin 192 on that matrix. */
int hg_stem_im2col(const float* in_nchw, void* out_rows, int32_t n, int32_t h, int32_t w, int32_t flip_w,
                   void* stream);

/* Stem without an im2col matrix (preferred path):
 *   hg_stem_pack : NCHW fp32 [n][3][h][w] -> NHWC4 bf16 [n][h][w+8][4] (interior only; the caller zero-
 *                  initialises the buffer once so the 4+4 padding pixels stay zero); flip_w = 1 mirrors; flip_w = 2
 *                  writes BOTH orientations from one read: `packed` is then [2n][h][w+8][4], images first, their
 *                  left-right mirrors second (the flip test's batch).
 *   hg_stem_conv : implicit GEMM over 8-pixel windows fetched by an overlapping-stride TMA tensor map.
 *                  weight: bf16 [64][224], k = ky*32 + (1+kx)*4 + c, zero at unused slots; bias fp32 [64];
 *                  out: bf16 NHWC [n][h/2][w/2][64] = relu(conv7x7s2(x) + bias).  h, w even; w <= 256 or a multiple of 256. */
int hg_stem_pack(const float* in_nchw, void* packed, int32_t n, int32_t h, int32_t w, int32_t flip_w, void* stream);
int hg_stem_conv(const void* packed, const void* weight, const float* bias, void* out, unsigned int* err_word, int32_t n,
                 int32_t h, int32_t w, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * Bandwidth-bound NHWC bf16 ops
 * ------------------------------------------------------------------------------------------- */
/* F.max_pool2d(x, 2, stride=2)  -- src/models/modules.py:82, src/models/hourglass.py:76 */
int hg_maxpool2x2_nhwc(const void* in, void* out, int32_t n, int32_t h, int32_t w, int32_t c, void* stream);
/* out = a + nearest_upsample_x2(low)  -- src/models/modules.py:90,95 (stand-alone form) */
int hg_upsample2x_add_nhwc(const void* a, const void* low, void* out, int32_t n, int32_t h, int32_t w, int32_t c,
                           void* stream);
/* out = relu(x*scale[c] + shift[c])  -- eval BatchNorm2d + ReLU (stand-alone form of the prologue) */
int hg_bn_relu_nhwc(const void* in, const float* scale, const float* shift, void* out, int64_t pixels, int32_t c,
                    void* stream);
/* layout converters at the API edge */
int hg_nchw_f32_to_nhwc_bf16(const float* in, void* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);
int hg_nhwc_bf16_to_nchw_f32(const void* in, float* out, int32_t n, int32_t c, int32_t h, int32_t w, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * Heat-map decode  (src/utils/evaluation.py:8-27, src/utils/inference.py:48-67,
 *                   src/utils/transforms.py:32-94)
 * ------------------------------------------------------------------------------------------- */
/* get_preds: hm fp32 [b][j][h][w] -> preds fp32 [b][j][2] (quirk coords, masked by maxval>0),
 * maxval fp32 [b][j] (may be NULL), flat argmax int32 [b][j] (may be NULL; first max wins). */
int hg_decode_argmax(const float* hm, float* preds, float* maxval, int32_t* argidx, int32_t b, int32_t j, int32_t h,
                     int32_t w, void* stream);
/* get_final_preds_v1 applied to EVERY batch element: argmax + quarter-pixel sign shift + inverse
 * affine.  center/scale: fp64 [b][2]; out: fp64 [b][j][2].  out_w/out_h = `output_size`. */
int hg_decode_final_preds(const float* hm, const double* center, const double* scale, double* out, int32_t b,
                          int32_t j, int32_t h, int32_t w, int32_t out_w, int32_t out_h, void* stream);

/* get_final_preds_v2 (src/utils/inference.py:70-87, DARK-style) for EVERY image of the batch: arg-max, 11x11
 * Gaussian blur (sigma 2) of the zero-padded map in float64 rescaled to the original peak, log(max(., 1e-10)),
 * second-order Taylor step for joints [0, refine_joints) -- the reference's loop covers joints 0 and 1 only
 * (`range(coords.shape[1])`, :79): pass 2 for parity, j for every joint -- then the same inverse affine.
 * Maps up to 128x128. */
int hg_decode_final_preds_v2(const float* hm, const double* center, const double* scale, double* out, int32_t b, int32_t j,
                             int32_t h, int32_t w, int32_t out_w, int32_t out_h, int32_t refine_joints, void* stream);
/* flip-test merge (defined from the reference's flip_pairs):
 * out[b][k] = 0.5*(hm[b][k] + mirror_w(hm_flip[b][perm[k]])) ; perm: int32 [j] on the device. */
int hg_flip_average(const float* hm, const float* hm_flip, const int32_t* perm, float* out, int32_t b, int32_t j,
                    int32_t h, int32_t w, void* stream);
/* PCK on heat maps (accuracy(), src/utils/evaluation.py:30-76), device part: per (b,j) normalised
 * distance or -1 -> dists fp32 [j][b]; the tiny per-joint averaging stays on the host. */
int hg_pck_dists(const float* out_hm, const float* tgt_hm, float* dists, int32_t b, int32_t j, int32_t h, int32_t w,
                 void* stream);

/* ------------------------------------------------------------------------------------------- *
 * JointsMSE loss (src/loss/mse.py:14-44) and Gaussian targets (src/datasets/common.py:197-248)
 * ------------------------------------------------------------------------------------------- */
#define HG_MAX_STACKS 16

/* Per-joint Gaussian centres: mu int32 [b][j][2] = (int(x/stride_x + .5), int(y/stride_y + .5)) with
 * C truncation, weight fp32 [b][j] = vis[..,0], forced to 0 when the (2r+1)^2 patch is entirely off the
 * map (common.py:217-227).  joints/vis: fp64 [b][j][3] in input-pixel coordinates. */
int hg_joint_centers(const double* joints, const double* vis, int32_t* mu, float* weight, int32_t b, int32_t j,
                     int32_t h, int32_t w, int32_t in_w, int32_t in_h, int32_t radius, void* stream);
/* target fp32 [b][j][h][w]: the (2r+1)x(2r+1) `patch` (device fp32, row-major; the host fills it with
 * exp(-((x-r)^2+(y-r)^2)/(2 sigma^2)) exactly as the reference's numpy does) pasted at mu, clipped, for
 * joints with weight > 0.5; zero elsewhere (common.py:229-246). */
int hg_gaussian_target(const int32_t* mu, const float* weight, const float* patch, float* target, int32_t b,
                       int32_t j, int32_t h, int32_t w, int32_t radius, void* stream);
/* loss = sum_s 1/(2*J*B*h*w) sum_{b,j,hw} (w*p - w*g)^2  accumulated into loss_out (device fp32[1],
 * zeroed by the caller); grad[s] = w^2 (p-g)/(J*B*h*w) * grad_scale.
 * preds/grads: HOST arrays of `stacks` device pointers to fp32 [b][j][h][w] (grads may be NULL).
 * target_weight: fp32 [b][j] or NULL (use_target_weight=False).
 * The target is either `target` (fp32 [b][j][h][w]) or, when target == NULL, generated on the fly from
 * (mu, patch, radius) with the visibility rule of hg_gaussian_target applied to `target_weight`. */
int hg_jmse_loss(const float* const* preds, float* const* grads, const float* target, const float* target_weight,
                 const int32_t* mu, const float* patch, int32_t radius, float* loss_out, int32_t stacks, int32_t b,
                 int32_t j, int32_t h, int32_t w, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------- *
 * Training (src/runner/trainer.py:82-99: loss.backward() + optimizer.step()).
 * Gradients w.r.t. activations ("dgrad") reuse hg_conv_nhwc_bf16 / hg_conv3x3_halo_bf16 on transposed
 * (3x3: also tap-flipped) weight copies; the entries below are the rest of the backward pass.
 * ------------------------------------------------------------------------------------------- */
/* conv.weight.grad as a split-K tcgen05 GEMM over pixels, reading both NHWC operands as they lie
 * (MN-major UMMA descriptors):  dw[(o-co_first)*ld + tap*tap_stride + i] += sum_f dout[f][o] * z[f + off(tap)][i]
 * for co_first <= o < co_valid, i < ci_valid (a window of dout's channels: the two groups of skip_mode='concat').  dout: bf16 [rows][co], z: bf16 [rows][ci]; dw is fp32 and must be
 * zero-initialised (or hold a partial sum): CTAs add with red.global.
 * taps == 1: off = 0.  taps == 9: both tensors are halo-padded buffers of identical geometry
 * (hg_conv3x3_halo_bf16) INCLUDING the leading zero row, rows = all positions, halo_pitch = w+1,
 * off(tap) = (tap/3-1)*halo_pitch + tap%3-1.   co <= 256 (multiple of 8), ci in {64,128,192,256}.
 * The pixel range is split over at most max_ctas CTAs (0: the default of 36, or the environment variable HG_WGRAD_CTAS):
 * the kernel is meant to run beside other work of the backward pass.  max_ctas = 1 disables the split, so that every
 * element of dw is accumulated by one CTA in one fixed order (bit-reproducible; validation runs). */
int hg_wgrad_bf16(const void* dout, const void* z, float* dw, unsigned int* err_word, int64_t rows, int32_t co,
                  int32_t co_first, int32_t co_valid, int32_t ci, int32_t ci_valid, int32_t taps, int32_t halo_pitch, int32_t ld,
                  int32_t tap_stride, int32_t max_ctas, void* stream);

/* Per-channel sum (and sum of squares) over the pixels of an NHWC bf16 tensor, ADDED into fp32
 * accumulators (zeroed by the caller): batch statistics of nn.BatchNorm2d in train mode, and conv.bias.grad.
 * sumsq may be NULL; channels >= c_valid are skipped.  c in {64,128,256}.
 * shift != 0: the sums are taken about k[c] = x[pixel 0][c] -- sum (x-k) and sum (x-k)^2 -- so that the variance derived
 * from them does not cancel when |mean| >> std (hg_bn_train_fwd with shifted != 0 undoes the shift).
 * scratch == NULL: blocks add their partial sums with atomics (the fp32 additions happen in order of arrival).
 * scratch != NULL (hg_colreduce_scratch_bytes(pixels, c) bytes of caller-owned device memory, zeroed ONCE before its first
 * use, not shared between launches that may run concurrently): blocks park their partial sums there and the last one to
 * arrive adds them in a fixed order -- bit-reproducible results whatever the scheduling. */
int64_t hg_colreduce_scratch_bytes(int64_t pixels, int32_t c);
int hg_colstats_nhwc(const void* x, float* sum, float* sumsq, int64_t pixels, int32_t c, int32_t c_valid, int32_t shift,
                     float* scratch, void* stream);

/* Train-mode BatchNorm2d (+ReLU) forward from batch sums (sums = [sum | sumsq], 2c floats):
 * out = [relu]((x - mean) * invstd * gamma + beta), biased variance, eps as given (torch: 1e-5).
 * Also writes saved = [mean | invstd | scale | shift] (4c floats, read by the backward kernels) and, when
 * running_mean != NULL, updates running_mean / running_var (unbiased) with `momentum` and increments
 * *num_batches_tracked (int64, may be NULL).   out_halo != 0: `out` is a halo-padded buffer (interior only
 * is written).  Reference: the BatchNorm2d modules of src/models/modules.py:11-19 under model.train(). */
int hg_bn_train_fwd(const void* x, const float* sums, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, int64_t* num_batches_tracked, float* saved, void* out, int32_t n, int32_t h,
                    int32_t w, int32_t c, int32_t out_halo, int32_t relu, int32_t shifted, float eps, float momentum,
                    void* stream);

/* BatchNorm2d(+ReLU) backward.  Pass 1 adds s1 = sum dY and s2 = sum dY*xhat into sums (2c floats, zeroed by the
 * caller), dY = dz masked by the recomputed ReLU.  Pass 2 writes
 *   out = scale*(dY - s1/N - xhat*s2/N) [+ add1] [+ add2]       (bf16 NHWC, or halo-padded when out_halo)
 * and dgamma = s2, dbeta = s1 (fp32 [c], may both be NULL).  `out` may alias add1/add2 (element-wise in place). */
int hg_bn_bwd_reduce(const void* dz, const void* x, const float* saved, float* sums, int64_t pixels, int32_t c,
                     int32_t relu, float* scratch /* as hg_colstats_nhwc */, void* stream);
int hg_bn_bwd_apply(const void* dz, const void* x, const float* saved, const float* sums, const void* add1,
                    const void* add2, void* out, float* dgamma, float* dbeta, int32_t n, int32_t h, int32_t w, int32_t c,
                    int32_t out_halo, int32_t relu, void* stream);

/* Depthwise 3x3, stride 1, zero padding 1 (mobile=True bottlenecks: nn.Conv2d(planes, planes, 3, padding=1,
 * groups=planes), src/models/modules.py:15-17) on NHWC bf16 [n][h][w][c]; w: fp32 [c][9] (torch's [c,1,3,3]
 * memory), bias fp32 [c] or NULL; out = (relu?)(conv + bias) rounded once to bf16.  flip_taps != 0 applies the
 * taps reversed: the convolution's input gradient.  c/8 must be a power of two <= 32.  CUDA-core stencil,
 * HBM-bound (9 MAC per element). */
int hg_dwconv3x3_nhwc(const void* x, const float* w, const float* bias, void* out, int32_t n, int32_t h, int32_t w_,
                      int32_t c, int32_t relu, int32_t flip_taps, void* stream);
/* Weight gradient of the depthwise 3x3: dw[c][tap] += sum over pixels dout[p][c] * z[p + tap][c] (fp32 [c][9],
 * accumulated into what is there). */
int hg_dwconv3x3_wgrad(const void* dout, const void* z, float* dw, int32_t n, int32_t h, int32_t w_, int32_t c,
                       void* stream);

/* F.max_pool2d(x,2,2) backward: dx (+)= dpool routed to the first maximum of each window (torch's rule).
 * With accumulate == 0 every element of dx is written.  src/models/modules.py:82, hourglass.py:76. */
int hg_maxpool2x2_bwd_nhwc(const void* x, const void* dpool, void* dx, int32_t n, int32_t h, int32_t w, int32_t c,
                           int32_t accumulate, void* stream);
/* F.interpolate(scale_factor=2, nearest) backward: dlow (+)= 2x2 sums of dy [n][h][w][c].  modules.py:90. */
int hg_sumpool2x2_nhwc(const void* dy, void* dlow, int32_t n, int32_t h, int32_t w, int32_t c, int32_t accumulate,
                       void* stream);
/* dst += src (bf16, fp32 add, one rounding): gradient fan-in of the residual stream. */
int hg_add_inplace_bf16(void* dst, const void* src, int64_t count, void* stream);
/* fp32 NCHW [n][c][h][w] -> bf16 NHWC [n][h][w][c_pad] with zero channels c..c_pad (heat-map gradients as a
 * GEMM operand).  c <= c_pad <= 64, c_pad a multiple of 8, out 16-byte aligned. */
int hg_nchw_f32_to_nhwc_bf16_pad(const float* in, void* out, int32_t n, int32_t c, int32_t c_pad, int32_t h, int32_t w,
                                 void* stream);

/* Weight packing, one launch for the whole network.  Every entry converts one fp32 master tensor
 * [co][taps][ci] (GEMM-natural order = torch channels_last) (+ optional src2, added element-wise) into
 *   dst_f32   [co*taps*ci] fp32 (optional),
 *   dst_fwd   bf16, element (o, r=tap*ci+i) at o*fwd_ld + fwd_col0 + r          (forward GEMM weight)
 *   dst_dgrad bf16, element at i*dgrad_ld + (taps-1-tap)*co + o                  (transposed, tap-flipped: the
 *             weight of the data-gradient convolution).   Null destinations are skipped.
 * The table lives in DEVICE memory. */
typedef struct hg_pack_entry {
    const float* src;
    const float* src2;
    float* dst_f32;
    void* dst_fwd;
    void* dst_dgrad;
    int32_t co, taps, ci;
    int32_t fwd_ld, fwd_col0, dgrad_ld;
} hg_pack_entry;
/* which: 1 = the forward layouts (dst_fwd, dst_f32), 2 = the transposed dgrad layouts, 3 = both. */
int hg_pack_weights(const hg_pack_entry* table_dev, int32_t n_entries, int32_t which, void* stream);

/* torch.optim.RMSprop step (momentum 0, centered False, weight_decay 0; trainer.py:39-41) over flat buffers:
 * g' = grad_scale*g;  v = alpha*v + (1-alpha)*g'^2;  p -= lr*g'/(sqrt(v)+eps).  count % 4 == 0. */
int hg_rmsprop_step(float* params, const float* grads, float* square_avg, int64_t count, float lr, float alpha, float eps,
                    float grad_scale, void* stream);

/* Strided fp32 GEMM for PARAMETER-space algebra only (merged remap weights and their chain rule):
 * C[i*sci + j*scj] = beta*C + (D ? D[same index] : 0) + sum_k A[i*sai + k*sak] * B[k*sbk + j*sbj]. */
int hg_small_gemm_f32(float* c, const float* a, const float* b, const float* d, int32_t m, int32_t n, int32_t k,
                      int32_t sai, int32_t sak, int32_t sbk, int32_t sbj, int32_t sci, int32_t scj, float beta,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HG_API_H_ */
