/*
 * imm_b200.h -- C ABI of libimm_b200.so: the sm_100a kernels of the IMM training hot path.
 *
 * The reference (tomasjakab/imm) has NO FFI/plugin interface: its hot path is a Python class
 * (imm/models/imm_model.py:95 IMMModel) whose arithmetic is executed by TensorFlow-1.10 ops.
 * Each entry point below therefore replaces one TF op call site of the reference; the call site is
 * cited as file:line (relative to /root/reference) next to the function.  INTEGRATION.md shows the
 * ctypes binding.
 *
 * Conventions
 *  - plain C, no torch types: raw device pointers + sizes + a cudaStream_t passed as void*.
 *  - tensors are NHWC fp32 with an explicit channel stride ("cstride", in elements) so that several
 *    producers can write into one concat buffer.
 *  - "split" tensors are a pair of fp32 planes (hi, lo) with hi = rna_tf32(v), lo = rna_tf32(v - hi).
 *    They are the operands of the 3xTF32 tensor-core convolutions.  lo may be NULL (single-pass TF32 /
 *    plain fp32 consumers use hi only, or hi+lo when lo is given).
 *  - every function only ENQUEUES work on the given stream; no allocation, no synchronisation, no
 *    host<->device copies, no global state except the thread-local last-error string.
 *  - return value: 0 on success, negative immb_status otherwise; never throws, never exits.
 *    Asynchronous CUDA errors surface at the caller's next synchronisation.
 */
#ifndef IMM_B200_H_
#define IMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMMB_VERSION 200

typedef enum {
  IMMB_OK = 0,
  IMMB_ERR_INVALID = -1,      /* bad argument / unsupported shape */
  IMMB_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed at enqueue time */
  IMMB_ERR_WORKSPACE = -3,    /* caller-provided workspace too small */
  IMMB_ERR_UNSUPPORTED = -4   /* engine cannot run this shape (e.g. forced tensor-core engine) */
} immb_status;

typedef enum {
  IMMB_ENGINE_AUTO = 0,       /* tcgen05 engine when the shape is eligible, else the SIMT engine */
  IMMB_ENGINE_SIMT = 1,       /* fp32 CUDA-core implicit GEMM (odd shapes: Cin=3/1, Cout=9/K; cross-check) */
  IMMB_ENGINE_TC = 2          /* tcgen05 + TMA implicit GEMM; fails with IMMB_ERR_UNSUPPORTED if not eligible */
} immb_engine;

typedef enum {
  IMMB_PREC_TF32X3 = 0,       /* error-compensated: hi*hi + hi*lo + lo*hi, fp32 accumulate (parity grade) */
  IMMB_PREC_TF32 = 1,         /* single pass on the hi planes (does NOT meet the 1e-3 parity bar) */
  IMMB_PREC_TF32X2 = 2,       /* hi*w + lo*w: exact 3xTF32 result when the WEIGHT operand is exactly representable in
                                 TF32 (its lo plane is all zeros), e.g. the frozen VGG16 tower whose weights are rounded
                                 to TF32 once at load; saves one of the three tensor-core passes */
  IMMB_PREC_F16X3 = 3,        /* scaled fp16 split planes ("H16": hi, lo*2^11, one power-of-two scale per tensor), products
                                 hi*hi + (hi*lo + lo*hi)*2^-11 at kind::f16 (twice the TF32 MMA rate, half the operand
                                 bytes), fp32 accumulate; same 22 significant bits per operand as IMMB_PREC_TF32X3 */
  IMMB_PREC_F16X2 = 4         /* the same with a weight operand that is exactly representable in scaled fp16 (frozen
                                 VGG16 tower, rounded once at load): hi*w + (lo*w)*2^-11 */
} immb_precision;

typedef enum {
  IMMB_XLAYOUT_NHWC = 0,      /* x is [N,H,W,x_cstride] */
  IMMB_XLAYOUT_ROWWIN4 = 1    /* 7x7 / Cin=3 / stride-1 first layer only: x is the staged image [N,H,W+8,4]
                                 (immb_stage_image_rowwin): 3 zero columns left, 5 right, 4th channel zero, so that
                                 the 7 taps of one filter row are ONE contiguous 128-byte TMA row per output pixel */
} immb_xlayout;

typedef enum {
  IMMB_EPI_BIAS = 0,          /* y = conv + b              (trainable stack: nn_utils.py:100,108) */
  IMMB_EPI_BIAS_RELU = 1      /* y = relu(conv + b)        (VGG: selfsup/vgg16.py:182-230) */
} immb_epilogue;

/* One convolution.  tf.nn.conv2d(x, w, [1,s,s,1], 'SAME') semantics: NHWC x HWIO cross-correlation,
 * TF SAME padding (pad_t/pad_l = the "before" pads; the remainder goes after). */
typedef struct {
  int32_t N, H, W, Cin;       /* input tensor (logical channels) */
  int32_t Cout, kh, kw, stride;
  int32_t Ho, Wo;             /* ceil(H/stride), ceil(W/stride) */
  int32_t pad_t, pad_l;
  int32_t x_cstride;          /* channel stride of the input tensor (>= Cin) */
  int32_t y_cstride;          /* channel stride of the output tensor (>= Cout) */
  int32_t cin_pad;            /* Cin rounded up to 32: K-extent of the packed weights (zero filled) */
  int32_t epilogue;           /* enum immb_epilogue; fwd only */
  int32_t precision;          /* immb_precision */
  int32_t engine;             /* immb_engine */
  int32_t x_layout;           /* immb_xlayout */
  int32_t reserved_;          /* keeps the pointers below 8-byte aligned */
  /* IMMB_PREC_F16*: scale records (device memory, int32 {e, amax bits}; see "H16 planes" below) of the tensors the
   * call touches as fp16 planes: x (input activations), y (the output planes of a forward call; dy in dgrad / wgrad)
   * and the packed weights.  NULL for the TF32 precisions. */
  int32_t* x_scale;
  int32_t* y_scale;
  int32_t* w_scale;
} immb_conv_desc;

int immb_version(void);
const char* immb_last_error(void);
/* which engine AUTO resolves to for this problem: returns IMMB_ENGINE_SIMT or IMMB_ENGINE_TC.
 * op: 0 fwd, 1 dgrad, 2 wgrad */
int immb_conv_engine_for(const immb_conv_desc* d, int op);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t immb_launch_count(void);

/* ---- convolution: nn_utils.py:100 (tf.nn.conv2d) + :108 (bias_add); vgg16.py:182-230 -------------- */
/* w      : master weights HWIO [kh,kw,Cin,Cout] (checkpoint layout, base_model.py:110)
 * wp_*   : packed copy  [kh*kw][Cout][cin_pad]   (K-major B operand of the forward GEMM)
 * wh_*   : split copy   [kh*kw][cin_pad][cout_pad] (K-major B operand of the dgrad GEMM; cout_pad = y_cstride)
 * y_lo   : NULL -> y_hi receives the full fp32 result; else the result is written as split planes.
 * Split planes (x_*, y_*, dy_*, wp_*, wh_*) are `void*`: fp32 TF32 pairs for the IMMB_PREC_TF32* precisions, scaled
 * fp16 pairs ("H16 planes", section below) with the scale records of the descriptor for IMMB_PREC_F16*. */
int immb_conv2d_fwd(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const float* w,
                    const void* wp_hi, const void* wp_lo, const float* bias, void* y_hi, void* y_lo,
                    void* stream);
/* dx[N,H,W,x_cstride] = conv2d_backprop_input(dy[N,Ho,Wo,y_cstride], w)  (autodiff of nn_utils.py:100) */
int immb_conv2d_dgrad(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const float* w,
                      const void* wh_hi, const void* wh_lo, float* dx, void* stream);
/* dw[kh,kw,Cin,Cout] = conv2d_backprop_filter(x, dy).  workspace: split-K partials (query size first). */
/* Forward conv with the batch-normalisation statistics of its output (tf.layers.batch_normalization(fused=True),
 * nn_utils.py:201: per-channel sum and sum of squares over N*Ho*Wo) accumulated in the epilogue, so the raw output is
 * not re-read by immb_bn_stats.  y[N,Ho,Wo,y_cstride] = conv + bias (single fp32 plane);  partials = rows x [2][Cout]
 * doubles, rows = immb_conv2d_fwd_stats_rows(d) (4 per CTA; 0 when the shape is not served by the 3-pass pair kernel),
 * one row per (CTA, epilogue warp) accumulated in a fixed order;  immb_bn_stats_from_partials adds them, in a fixed
 * order, to sums[2*Cout] (zeroed by the caller): deterministic, no atomics. */
int immb_conv2d_fwd_stats_rows(const immb_conv_desc* d);
int immb_conv2d_fwd_bnstats(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* wp_hi,
                            const void* wp_lo, const float* bias, float* y, double* partials, size_t partial_elems,
                            void* stream);
int immb_bn_stats_from_partials(const double* partials, int rows, int C, double* sums, void* stream);
/* immb_bn_stats_from_partials + immb_bn_finalize(training = 1) in ONE launch: sums the partial rows of the conv epilogue
 * in the same fixed order and finalises every channel (batch mean / biased variance, moving statistics updated in
 * place, scale / shift / mean / invstd written). */
int immb_bn_finalize_partials(const double* partials, int rows, int64_t count, int C, const float* gamma,
                              const float* beta, float* moving_mean, float* moving_var, float* scale, float* shift,
                              float* mean, float* invstd, void* stream);
/* dgrad that also accumulates, in its epilogue, the two per-channel sums of the BN backward of the layer that produced
 * the conv's input (tf.layers.batch_normalization gradient, nn_utils.py:201-209): sum(dz) and sum(dz * xhat) with
 * dz = dx * [relu ? y*scale+shift > 0 : 1], xhat = (y - mean) * invstd, y = that layer's raw conv output
 * [N,H,W,y_prev_cstride].  dx is written as by immb_conv2d_dgrad; partials = rows x [2][Cin] doubles
 * (rows = immb_conv2d_dgrad_stats_rows(d), 0 when not served), reduced by immb_bn_stats_from_partials into the
 * `sums` immb_bn_bwd_apply reads -- replaces the immb_bn_bwd_reduce pass over dx and y. */
int immb_conv2d_dgrad_stats_rows(const immb_conv_desc* d);
int immb_conv2d_dgrad_bnreduce(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                               const void* wh_lo, float* dx, const float* y_prev, int y_prev_cstride,
                               const float* scale, const float* shift, const float* mean, const float* invstd,
                               int relu, double* partials, size_t partial_elems, void* stream);
/* dgrad fused with the backward of the ReLU that produced the conv's input (vgg16.py:229-236: activations are stored
 * post-ReLU) and with the TF32 split of the result: out_{hi,lo}[N,H,W,x_cstride] = split(dgrad(dy) * [act_hi > 0]),
 * act_hi = hi plane of the conv's (post-ReLU) input, channel stride act_cstride.  Only shapes served by the halo pair
 * kernel (immb_conv2d_dgrad_relu_supported); saves one write + one read of the fp32 gradient and one launch per layer. */
int immb_conv2d_dgrad_relu_supported(const immb_conv_desc* d);
/* out_scale: scale record of the output planes (IMMB_PREC_F16*; NULL for the TF32 precisions) */
int immb_conv2d_dgrad_relu(const immb_conv_desc* d, const void* dy_hi, const void* dy_lo, const void* wh_hi,
                           const void* wh_lo, const void* act_hi, int act_cstride, void* out_hi, void* out_lo,
                           int32_t* out_scale, void* stream);
size_t immb_conv2d_wgrad_workspace(const immb_conv_desc* d);
int immb_conv2d_wgrad(const immb_conv_desc* d, const void* x_hi, const void* x_lo, const void* dy_hi,
                      const void* dy_lo, float* dw, void* workspace, size_t ws_bytes, void* stream);
/* master HWIO -> wp_{hi,lo} [taps][Cout][cin_pad], wh_{hi,lo} [taps][cin_pad][cout_pad] (either pair may be NULL;
 * padding is zero filled; cout_pad >= Cout is the channel stride of the layer's output / dy tensors) */
/* H16: amax = device pointer to the tensor's largest |w| (immb_multi_amax / immb_adam_apply), w_scale = its scale
 * record, written here (weights are rescaled in the step that changes them: no delayed scaling); both NULL for TF32. */
int immb_pack_weights(const float* w, int kh, int kw, int Cin, int Cout, int cin_pad, int cout_pad, void* wp_hi,
                      void* wp_lo, void* wh_hi, void* wh_lo, const float* amax, int32_t* w_scale, void* stream);
/* first-layer staging for IMMB_XLAYOUT_ROWWIN4: image [N,H,W,3] -> x4 split planes [N,H,W+8,4];
 * weights [7,7,3,Cout] -> wp_{hi,lo} [7][Cout][32] with k = s*4 + c (zero for c == 3 and s == 7) */
/* Every weight tensor in ONE launch: `items` = device array of n_items records, block ranges assigned by the caller
 * (item i owns blocks [block0, block0 + nblocks), nblocks = ceil(elements / 2048), elements = taps*cin_pad*cout_pad or
 * 7*Cout*32 for kind 2); total_blocks = sum of nblocks.  Same results as the per-tensor calls. */
typedef struct {
  const float* w;             /* master HWIO weights */
  void *wp_hi, *wp_lo;        /* packed [taps][Cout][cin_pad] planes (kind 2: [7][Cout][32]) */
  void *wh_hi, *wh_lo;        /* split  [taps][cin_pad][cout_pad] planes (may be NULL) */
  const float* amax;          /* kind 1: the tensor's largest |w| */
  int32_t* rec;               /* kind 1: its scale record */
  int32_t taps, Cin, Cout, cin_pad, cout_pad;
  int32_t kind;               /* 0 fp32 TF32 planes, 1 scaled fp16 planes, 2 row-window first layer (TF32 planes) */
  int32_t block0, nblocks;
} immb_pack_item;
int immb_pack_weights_multi(const immb_pack_item* items, int n_items, int total_blocks, void* stream);
int immb_stage_image_rowwin(const float* image, int N, int H, int W, float* x4_hi, float* x4_lo, void* stream);
int immb_pack_weights_rowwin(const float* w, int Cout, float* wp_hi, float* wp_lo, void* stream);
/* v -> (hi, lo) planes, contiguous n elements */
int immb_split_planes(const float* v, void* hi, void* lo, int64_t n, int32_t* scale, void* stream);

/* ---- H16 planes: scaled fp16 split operands (IMMB_PREC_F16X3 / F16X2) -------------------------------------------
 * A tensor v is held as two fp16 planes and one power-of-two scale 2^e:
 *     t = v * 2^e;   hi = rn_f16(t);   lo = rn_f16((t - hi) * 2^11);        v ~= (hi + lo * 2^-11) * 2^-e
 * (22 significant bits, like the TF32 pair; the scale is exact, so results do not depend on e while hi stays inside
 * fp16's normal range).  Every H16 tensor has a SCALE RECORD in device memory, int32 {e, amax_bits}: a kernel that
 * writes the planes reads e and atomically maxes the bit pattern of the largest |v| it wrote into amax_bits; a kernel
 * that reads the planes reads e.  immb_scale_update turns the observed maxima into the exponents of the NEXT step
 * (delayed scaling: largest magnitude near 2^8, i.e. 256x headroom before fp16 saturates -- gradient planes were seen to
 * grow 20-40x between consecutive steps early in training -- and 22 binades of full hi precision below), counts tensors whose maximum did not fit (overflow[0] += 1; immb_total_loss then poisons the loss
 * with NaN so that the caller's NaN guard fires) and clears the maxima.  It must run while no H16 activation / gradient
 * tensor is live, i.e. between training steps.  Weight planes are rescaled by immb_pack_weights from the exact maximum
 * of the updated weights instead.
 * Convention for the functions below: a plane pair is `void*`; its trailing `*_scale` argument is the scale record for
 * fp16 planes and NULL for fp32 TF32 planes. */
int immb_scale_update(int32_t* recs, int n, int32_t* overflow, void* stream);
/* amax[t] (float, zeroed by the caller) = max |p| over tensor t of the flat parameter buffer (chunk table as below) */
int immb_multi_amax(const float* p, const int32_t* chunk_tensor, const int64_t* chunk_off, const int32_t* chunk_len,
                    int n_chunks, float* amax, void* stream);

/* ---- batch norm: nn_utils.py:201 tf.layers.batch_normalization(training=..., fused=True) ---------- */
/* sums[2*C] (double, zeroed by the caller): sum(y), sum(y^2) per channel over npix pixels.
 * scratch (optional, caller-owned, scratch_elems doubles): when large enough (immb_bn_scratch_elems) the per-block
 * partial sums are written there and combined by a second tiny kernel in a fixed order -- deterministic and free of
 * same-address atomics; otherwise (NULL / too small) the blocks atomically add into sums. */
size_t immb_bn_scratch_elems(int64_t npix, int C);
int immb_bn_stats(const float* y, int64_t npix, int C, int y_cstride, double* sums, double* scratch,
                  size_t scratch_elems, void* stream);
/* training!=0: batch mean / biased var from sums; moving_mean/var updated in place (momentum .99, Bessel);
 * training==0: moving stats.  Outputs: scale = gamma*rsqrt(var+1e-3), shift = beta-mean*scale, mean, invstd. */
int immb_bn_finalize(const double* sums, int64_t count, int C, const float* gamma, const float* beta,
                     float* moving_mean, float* moving_var, int training, float* scale, float* shift,
                     float* mean, float* invstd, void* stream);
/* out = [relu](y*scale+shift), written as split planes; up2x!=0 additionally applies
 * tf.image.resize_images(x, 2x) (imm_model.py:175; TF1 legacy bilinear) to the activated tensor, so the
 * output is [N,2H,2W,C]. */
int immb_bn_apply(const float* y, int N, int H, int W, int C, int y_cstride, const float* scale,
                  const float* shift, int relu, int up2x, void* out_hi, void* out_lo, int out_cstride,
                  int32_t* out_scale, void* stream);
/* adjoint of the legacy x2 bilinear resize: g_up[N,2H,2W,C] -> g[N,H,W,C] */
int immb_upsample2x_bwd(const float* g_up, int N, int H, int W, int C, int gup_cstride, float* g, void* stream);
/* sums[2*C] (double, zeroed): sum(dz), sum(dz*xhat) with dz = g * (relu ? (y*scale+shift > 0) : 1) */
int immb_bn_bwd_reduce(const float* g, int g_cstride, const float* y, int y_cstride, int64_t npix, int C,
                       const float* scale, const float* shift, const float* mean, const float* invstd,
                       int relu, double* sums, double* scratch, size_t scratch_elems, void* stream);
/* dy = scale*(dz - mean(dz) - xhat*mean(dz*xhat)) as split planes; dgamma = sum(dz*xhat), dbeta = sum(dz),
 * dbias = sum(dy) (double accumulators dbias_acc[C], zeroed by the caller; finalised by immb_cast_d2f).
 * dbias_out (optional, needs the scratch path): the bias gradient is written there as float by the second-level
 * reduction and dbias_acc is left untouched -- no accumulator clearing, no cast launch. */
int immb_bn_bwd_apply(const float* g, int g_cstride, const float* y, int y_cstride, int64_t npix, int C,
                      const float* scale, const float* shift, const float* mean, const float* invstd,
                      int relu, const double* sums, void* dy_hi, void* dy_lo, float* dgamma, float* dbeta,
                      double* dbias_acc, double* scratch, size_t scratch_elems, int32_t* dy_scale, float* dbias_out,
                      int defer_dbias, void* stream);
/* defer_dbias != 0 (scratch path): the bias-gradient partials stay in scratch as [immb_bn_bwd_apply_blocks(npix, C)][C]
 * doubles; the caller sums the partials of ALL layers with one immb_reduce_partials_multi launch (items: device array;
 * item i owns blocks [block0, block0 + ceil(nvals / 32)); out_f[v] = float(sum over nblocks rows), fixed order). */
typedef struct {
  const double* partials;
  float* out_f;
  int32_t nblocks, nvals, block0, reserved_;
} immb_reduce_item;
int immb_bn_bwd_apply_blocks(int64_t npix, int C);
int immb_reduce_partials_multi(const immb_reduce_item* items, int n_items, int total_blocks, void* stream);
/* column sums: acc[C] (double, zeroed) += sum over pixels of g[:, c] */
int immb_bias_grad(const void* g_hi, const void* g_lo, int g_cstride, int64_t npix, int C, double* acc,
                   const int32_t* g_scale, void* stream);
int immb_cast_d2f(const double* src, float* dst, int64_t n, void* stream);

/* ---- landmark bottleneck: imm_model.py:252-263 (get_coord) + :34-78 (get_gaussian_maps, 'rot') ---- */
/* heat[B,S,S,K] -> mu[B,K,2] (y,x), py[B,S,K], px[B,S,K]; and the Sg x Sg Gaussian maps written as split
 * planes into channels [c_off, c_off+K) of a [B,Sg,Sg,out_cstride] buffer (the renderer's concat input,
 * imm_model.py:341-344).  maps_hi may be NULL (coordinates only). */
int immb_softargmax_gauss_fwd(const float* heat, int B, int S, int K, int heat_cstride, float inv_std,
                              float* mu, float* py, float* px, int Sg, void* maps_hi, void* maps_lo,
                              int out_cstride, int c_off, int32_t* maps_scale, void* stream);
/* g_maps[B,Sg,Sg,g_cstride] (channels c_off..c_off+K) -> g_heat[B,S,S,K] */
int immb_softargmax_gauss_bwd(const float* g_maps, int g_cstride, int c_off, const float* mu,
                              const float* py, const float* px, int B, int S, int K, int Sg, float inv_std,
                              float* g_heat, int gheat_cstride, void* stream);
/* free function get_gaussian_maps(mu, [S,S], inv_std, mode='rot') -> maps[B,S,S,K] fp32 */
int immb_gaussian_maps(const float* mu, int B, int K, int S, float inv_std, float* maps, void* stream);

/* ---- perceptual tower glue: build_vgg16.py:22-26, ops.py:16-26, imm_model.py:111-151,408-410 ------ */
/* gray = mean_c(rgb)/255 - 114.451/255 for [gt ; pred[..., :3]] as split planes.
 * patches == 0: out is vgg_in [2B,R,R,1].
 * patches != 0: out is the 3x3 SAME-padded patch tensor [2B,R,R,12] (channel r*3+s = gray[h+r-1, w+s-1], channels
 *   9..11 zero), which turns conv1_1 (Cin = 1) into a tensor-core friendly 1x1 convolution with Cin = 9. */
int immb_vgg_prologue(const float* gt, const float* pred, int pred_cstride, int B, int R, int patches,
                      float* out_hi, float* out_lo, void* stream);
/* VGG conv1_1 fused with the prologue (vgg16.py:182-230 on build_vgg16.py:22-26): out[2B,R,R,Cout] split planes =
 * relu(conv3x3_SAME(gray([gt ; pred])) + b), w [3,3,1,Cout] HWIO.  Cin = 1 makes this layer HBM-bound (it writes
 * 2*Cout*4 bytes per pixel), so it runs on the CUDA cores in exact fp32 straight from the RGB inputs.
 * which: 0 = both halves, 1 = only the gt half (images [0,B) of out; pred may be NULL), 2 = only the pred half
 * (images [B,2B) of out; gt may be NULL) -- the gt half depends on the input batch alone and can run on its own stream. */
int immb_vgg_conv1_1_fused(const float* gt, const float* pred, int pred_cstride, int B, int R, const float* w,
                           const float* bias, int Cout, void* out_hi, void* out_lo, int which, int32_t* out_scale,
                           void* stream);
/* 2x2/2 max pool on split planes [N,H,W,C] -> [N,H/2,W/2,C] split planes */
int immb_maxpool2x2_fwd(const void* x_hi, const void* x_lo, int N, int H, int W, int C, void* o_hi,
                        void* o_lo, const int32_t* x_scale, int32_t* o_scale, void* stream);
/* The same pool on a perceptual level x = [gt ; pred] ([2B,H,W,C] split planes, B images each) fused with that level's
 * masked squared-difference sum (imm_model.py:143-147, _loss_mask :408-410): acc[0] += sum mask[b, y*R/H, x*R/H] *
 * (x[b] - x[B+b])^2 (mask [B,R,R,1] or NULL).  Replaces immb_perceptual_level_sum for levels that feed a pool. */
int immb_maxpool2x2_fwd_levelsum(const void* x_hi, const void* x_lo, int B, int H, int W, int C, void* o_hi,
                                 void* o_lo, const float* mask, int R, double* acc, const int32_t* x_scale,
                                 int32_t* o_scale, void* stream);
/* g_in[N,H,W,C] from g_out[N,H/2,W/2,C]; first-max-wins tie rule of TF's CPU MaxPoolGrad */
int immb_maxpool2x2_bwd(const float* g_out, const float* x_hi, const float* x_lo, int N, int H, int W,
                        int C, float* g_in, void* stream);
/* acc[0] (double, zeroed) += sum_{b,h,w,c} m[b, h*s, w*s] * (fg[b] - fp[b])^2 ; fg / fp are the gt / pred
 * halves [B,h,w,C] of one feature level (separate pointers + channel strides; for the VGG levels fp is the
 * second half of the same [2B,...] activation).  *_lo may be NULL.  mask[B,R,R,1] may be NULL.  s = R/h. */
int immb_perceptual_level_sum(const void* fg_hi, const void* fg_lo, int fg_cstride, const void* fp_hi,
                              const void* fp_lo, int fp_cstride, int B, int h, int w, int C,
                              const float* mask, int R, double* acc, const int32_t* f_scale, void* stream);
/* Given the n_levels sums: s_k = acc_k / count_k; wl_k = a_k + 0.01 (s_k - a_k); L_k = s_k / wl_k;
 * rec = 1000 sum L_k; coef_k = 1000 * 0.99 a_k / wl_k^2 * (-2 / count_k)  (gradient wrt f_pred is coef*m*d);
 * training: a_k <- wl_k (base_model.py:39-50).  out: levels[n_levels], rec_loss[1], coef[n_levels]. */
int immb_perceptual_finalize(const double* acc, const double* counts, int n_levels, float* agg, int training,
                             float* levels, float* rec_loss, float* coef, void* stream);
/* VGG backward glue for one activation a_k = relu(.) of the pred half:
 *   g = (g_next ? g_next : 0) + (coef ? coef[0]*m*(fg - fp) : 0);  dy = g * (fp > 0) -> split planes
 * fg / fp: gt / pred halves [B,h,w,C] (contiguous, split planes; fg may be NULL when coef is NULL);
 * g_next [B,h,w,C] or NULL. */
int immb_vgg_bwd_combine(const float* g_next, const void* fg_hi, const void* fg_lo, const void* fp_hi,
                         const void* fp_lo, int B, int h, int w, int C, const float* mask, int R,
                         const float* coef, void* dy_hi, void* dy_lo, const int32_t* f_scale, int32_t* dy_scale,
                         void* stream);
/* immb_maxpool2x2_bwd followed by immb_vgg_bwd_combine in one pass, for a VGG activation that feeds a pool (ops.py:16-26
 * backward + imm_model.py:143-147 backward + ReLU backward): dy = split([fp > 0] * ([fp is the first max of its 2x2
 * window] * g_out + coef * mask * (fg - fp))), g_out [B,H/2,W/2,C], fg / fp the gt / pred halves [B,H,W,C] of the
 * activation; coef NULL = no loss term at this layer.  No fp32 gradient round trip, fp read once. */
int immb_maxpool2x2_bwd_combine(const float* g_out, const void* fg_hi, const void* fg_lo, const void* fp_hi,
                                const void* fp_lo, int B, int H, int W, int C, const float* mask, int R,
                                const float* coef, void* dy_hi, void* dy_lo, const int32_t* f_scale, int32_t* dy_scale,
                                void* stream);
/* gradient wrt the renderer output [B,R,R,pred_cstride] (channels >=3 get 0) as split planes:
 *   g_pred_c = coef_input * m * (gt_c - pred_c) + g_gray / (3*255)
 * g_vggin (may be NULL): gradient wrt the VGG input; [B,R,R,1] when g_is_patch == 0, else the gradient wrt the
 * [B,R,R,12] patch tensor (g_gray[h,w] = sum_{r,s} g[h-r+1, w-s+1, r*3+s], the adjoint of the patch extraction). */
int immb_pred_grad(const float* gt, const float* pred, int pred_cstride, const float* mask,
                   const float* coef_input, const float* g_vggin, int g_is_patch, int B, int R, float* g_hi,
                   float* g_lo, void* stream);
/* Backward of the VGG input stage in one kernel (vgg16.py:182-230 backward for conv1_1, build_vgg16.py:22-26 adjoint,
 * imm_model.py:143-147 'input' level): the gradient wrt the renderer output from the split gradient dy[B,R,R,64] wrt
 * conv1_1's raw output (ReLU backward already applied), w [3,3,1,64] HWIO:
 *   g_pred_c = coef_input * m * (gt_c - pred_c) + (1/(3*255)) * sum_{r,s,co} dy[h-r+1, w-s+1, co] * w[r,s,0,co]  (c < 3)
 * Equals immb_conv2d_dgrad on the patch form + immb_pred_grad(g_is_patch = 1), in exact fp32. */
int immb_vgg_conv1_1_bwd_fused(const void* dy_hi, const void* dy_lo, const float* w, int Cout, const float* gt,
                               const float* pred, int pred_cstride, const float* mask, const float* coef_input,
                               int B, int R, void* g_hi, void* g_lo, const int32_t* dy_scale, int32_t* g_scale,
                               void* stream);
/* tf.image.resize_bilinear(align_corners=True) (imm_model.py:334) on split planes, and its adjoint */
int immb_resize_ac_fwd(const void* x_hi, const void* x_lo, int x_cstride, int N, int H, int W, int C,
                       int Ho, int Wo, void* o_hi, void* o_lo, int o_cstride, const int32_t* x_scale,
                       int32_t* o_scale, void* stream);
int immb_resize_ac_bwd(const float* g_out, int g_cstride, int N, int H, int W, int C, int Ho, int Wo,
                       float* g_in, void* stream);

/* ---- input pipeline: imm/utils/tps_sampler.py:77-165 + tps_dataset.py:70-96 ----------------
 * dst[B,H,W,C] = grid_sample(src[B,H,W,C], TPSGridGen(H,W,Hc,Wc)(w_tps[B,Hc*Wc+3,2])): thin-plate-spline grid
 * (U(d2) = d2 log d2 on a regular Hc x Wc control lattice in [-1,1]^2, affine rows last) evaluated per output pixel
 * and sampled bilinearly with zero padding and corner-aligned coordinates (PyTorch-0.4 grid_sample), C <= 8. */
int immb_tps_warp(const float* src, int B, int H, int W, int C, const float* w_tps, int Hc, int Wc, float* dst,
                  void* stream);

/* ---- optimiser: cnn_train_multi.py:93-98,232-241 + scripts/train.py:92-98 + nn_utils.py:44-46 ---- */
/* Flat buffers of n floats hold every trainable tensor back to back.  The chunk table splits them into
 * chunks: chunk_tensor[i] = tensor id, chunk_off[i] = start offset, chunk_len[i] <= 1024*? elements.
 * Step 1 (norms): sq[t] (double, zeroed) += sum (g*gscale + wd_t*p)^2; wsq[t] += sum p^2.
 * Step 2 (apply): g' = (g*gscale + wd_t*p) * clip/max(||.||, clip); TF Adam with lr_t (device scalar pair
 * hyper[0]=lr_t, hyper[1]=clip); params updated in place.  amax (optional, float[n_tensors], zeroed by the caller):
 * per-tensor max |p| of the UPDATED parameters, the input of immb_pack_weights' fp16 scaling. */
int immb_adam_norms(const float* p, const float* g, int64_t n, const int32_t* chunk_tensor,
                    const int64_t* chunk_off, const int32_t* chunk_len, int n_chunks, const float* tensor_wd,
                    float gscale, double* sq, double* wsq, void* stream);
int immb_adam_apply(float* p, const float* g, float* m, float* v, int64_t n, const int32_t* chunk_tensor,
                    const int64_t* chunk_off, const int32_t* chunk_len, int n_chunks, const float* tensor_wd,
                    float gscale, const double* sq, float clip, float lr_t, float beta1, float beta2,
                    float eps, float* amax, void* stream);
/* Same update with the step-dependent scalar lr_t = lr*sqrt(1-beta2^t)/(1-beta1^t) read from device memory, so that a
 * captured CUDA graph of the whole training step can be replayed every iteration (the host refreshes lr_t_dev[0] with
 * one 4-byte async copy before each replay). */
int immb_adam_apply_dev(float* p, const float* g, float* m, float* v, int64_t n, const int32_t* chunk_tensor,
                        const int64_t* chunk_off, const int32_t* chunk_len, int n_chunks, const float* tensor_wd,
                        float gscale, const double* sq, float clip, const float* lr_t_dev, float beta1, float beta2,
                        float eps, float* amax, void* stream);
/* total = rec_loss[0] + sum_t 0.5*wd_t*wsq[t]  (imm_model.py:395-400, base_model.py:33-37).
 * overflow (optional): the saturation counter of immb_scale_update; when positive the total is NaN. */
int immb_total_loss(const float* rec_loss, const double* wsq, const float* tensor_wd, int n_tensors,
                    float* weights_loss, float* total, const int32_t* overflow, void* stream);

/* ---- host utility: CRC-32C (Castagnoli) of n bytes continuing from `crc` (0 to start), as
 * TensorFlow's lib/hash/crc32c used by the TensorBundle files behind tf.train.Saver (cnn_train_multi.py:432-439,
 * 511-513).  Pure host code, no CUDA call. */
uint32_t immb_crc32c(const void* data, size_t n, uint32_t crc);

#ifdef __cplusplus
}
#endif
#endif /* IMM_B200_H_ */
