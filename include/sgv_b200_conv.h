/*
 * sgv_b200 — C ABI, part 2: the dense contraction of the hot path on tcgen05 tensor cores.
 *
 * Replaces the library calls the reference makes for the contraction — torch.nn.functional.conv2d /
 * conv_transpose2d -> cuDNN (src/torch_utils/ops/conv2d_gradfix.py:35-43,108-116) and
 * aten::cudnn_convolution_backward_weight (conv2d_gradfix.py:140-148) — and fuses into it what the
 * reference runs as separate elementwise passes around the contraction in modulated_conv2d
 * (src/training/networks.py:64-74: x*styles before, *dcoefs after) and bias_act (bias_act.cu:23-147).
 *
 * Layout: activations are NHWC fp32 (torch channels_last); weights are pre-arranged once per call by
 * sgv_conv_prep_weights into [tap][out_channel][in_channel] fp32 rounded to TF32 (round-to-nearest).
 * Arithmetic, two modes:
 *   tf32x1 (default)  TF32 x TF32 products, fp32 accumulation in TMEM (tcgen05.mma.kind::tf32): <= 1e-3 of fp32 per contraction.
 *   tf32x3            fp32-grade: every operand is split into hi = tf32(v) and lo = tf32(v - hi); the kernels accumulate
 *                     hi*hi + lo*hi + hi*lo (three tensor-core products per term, fp32 accumulation): ~1e-6 of fp32, the mode that
 *                     corresponds to the reference's torch.backends.cudnn.allow_tf32 = False (training_loop.py:141-142).
 *                     Selected per call: sgv_conv_params.wp_lo != NULL, sgv_wgrad_params.precision = 1.
 * Same contract as include/sgv_b200.h: caller-owned buffers, no allocation, no synchronisation, work is
 * enqueued on `stream`, 0 on success.
 */
#ifndef SGV_B200_CONV_H
#define SGV_B200_CONV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGV_CONV_MAX_TAPS 16

/* Gathers + rounds weights:  wp[t][r][k] = tf32_rn( w[r*stride_row + k*stride_col + tap_ky[t]*stride_ky + tap_kx[t]*stride_kx] )
 * for t < ntaps, r < rows, k < cols.  (rows = GEMM N = output channels of the contraction, cols = GEMM K.) */
int sgv_conv_prep_weights(const float* w, int64_t stride_row, int64_t stride_col, int64_t stride_ky, int64_t stride_kx,
                          int32_t rows, int32_t cols, int32_t ntaps, const int32_t* tap_ky, const int32_t* tap_kx,
                          float* wp, void* stream);
/* The same with a scale folded in (equalised-lr weight gain, layers.py:131,361) and, for the tf32x3 mode, the hi/lo split:
 *   v = w[...] * w_scale;   wp[t][r][k] = tf32_rn(v);   wp_lo[t][r][k] = tf32_rn(v - wp[t][r][k])   (wp_lo may be NULL) */
int sgv_conv_prep_weights_ex(const float* w, int64_t stride_row, int64_t stride_col, int64_t stride_ky, int64_t stride_kx,
                             int32_t rows, int32_t cols, int32_t ntaps, const int32_t* tap_ky, const int32_t* tap_kx,
                             float w_scale, float* wp, float* wp_lo, void* stream);

/* Both slab sets of a layer from one read of a DENSE weight w[out_ch][in_ch][kh][kw] (1x1 or 3x3, channel counts % 32 == 0):
 *   wp_a[t][o][i] = tf32_rn(w[o][i][a_ky[t]][a_kx[t]])   (forward contraction: rows = out_ch)          — skipped when wp_a is NULL
 *   wp_b[t][i][o] = tf32_rn(w[o][i][b_ky[t]][b_kx[t]])   (data-gradient contraction: rows = in_ch)     — skipped when wp_b is NULL */
int sgv_conv_prep_weights_pair(const float* w, int32_t out_ch, int32_t in_ch, int32_t kh, int32_t kw,
                               int32_t ntaps_a, const int32_t* a_ky, const int32_t* a_kx, float* wp_a,
                               int32_t ntaps_b, const int32_t* b_ky, const int32_t* b_kx, float* wp_b, void* stream);
/* tf32x3 variant: additionally writes the residual slabs wp_a_lo / wp_b_lo (same shapes; lo = tf32_rn(w - tf32_rn(w))); each may be NULL
 * together with its hi set. */
int sgv_conv_prep_weights_pair_x3(const float* w, int32_t out_ch, int32_t in_ch, int32_t kh, int32_t kw,
                                  int32_t ntaps_a, const int32_t* a_ky, const int32_t* a_kx, float* wp_a, float* wp_a_lo,
                                  int32_t ntaps_b, const int32_t* b_ky, const int32_t* b_kx, float* wp_b, float* wp_b_lo, void* stream);

/* y[n, oy, ox, o] = epilogue( sum_{t, i}  x[n, oy*in_stride + tap_dy[t], ox*in_stride + tap_dx[t], i] * a_scale[n, i] * wp[t][o][i] )
 *   epilogue(v) = clamp( act( v * o_scale[n, o] + noise[n, oy, ox] + bias[o] ) * gain )        (each piece optional)
 * Out-of-range input pixels read as zero.  Output element (n, oy, ox, o) lives at
 *   y + n*out_stride_n + oy*out_stride_y + ox*out_stride_x + o        (element strides; lets one call write a
 *   polyphase sub-lattice of a larger tensor, which is how the stride-2 transposed convolution is issued).
 * Requirements: cin % 32 == 0, cout % 64 == 0 (or cout == 32), x and wp 16-byte aligned, y and strides multiples of 4 elements.
 */
typedef struct sgv_conv_params {
    const float* x;            /* [n, h, w, cin] NHWC */
    const float* wp;           /* [ntaps, cout, cin] from sgv_conv_prep_weights */
    float*       y;
    int32_t n, h, w, cin, cout;
    int32_t out_h, out_w;
    int64_t out_stride_n, out_stride_y, out_stride_x;
    int32_t in_stride;         /* 1 or 2 */
    int32_t ntaps;
    int32_t tap_dy[SGV_CONV_MAX_TAPS], tap_dx[SGV_CONV_MAX_TAPS];
    const float* a_scale;      /* [n, cin]  or NULL : StyleGAN modulation (styles), dcoefs for the data gradient */
    const float* o_scale;      /* [n, cout] or NULL : demodulation coefficients, styles for the data gradient */
    const float* bias;         /* [cout] or NULL */
    int32_t act;               /* 1 = linear, 3 = lrelu (bias_act cuda_idx numbering) */
    float   alpha, gain, clamp;/* clamp < 0 disables */
    /* optional: x is a strided VIEW [n, h, w, cin] of a larger NHWC tensor (element strides, channel stride 1); all zero = dense.
     * Used to address one polyphase sub-lattice (pixel stride 2) of the transposed-conv gradient without a stride-2 gather. */
    int64_t in_stride_n, in_stride_y, in_stride_x;
    int32_t accumulate;        /* 1: y += result (no o_scale/bias/act allowed) — sums the four polyphase data-gradient launches */
    /* optional fused reduction (gradient of the style modulation, networks.py:66): red_out[n, o] += sum_{oy,ox} raw[n,oy,ox,o] * red_x[n,oy,ox,o]
     * where raw is the accumulator BEFORE o_scale/bias/act and red_x is a tensor addressed exactly like y (same strides).
     * With o_scale = styles this turns the data-gradient launch into  dx = dxs * s  and  dstyles = sum_hw dxs * x  in one pass. */
    const float* red_x;
    float*       red_out;      /* [n, cout] float32, caller-zeroed */
    /* optional: x already holds TF32-representable values and needs no scaling (a_scale must be NULL): the persistent kernel skips its
     * operand-staging pass over the activation patches (they go TMA -> MMA directly) */
    int32_t      a_ready;
    /* optional, tf32x3 mode: residual slabs [ntaps, cout, cin] from sgv_conv_prep_weights_ex / _pair_x3 (same allocation as wp, at a
     * non-negative offset that is a multiple of cin elements).  a_ready must be 0 (the activations are split inside the kernel). */
    const float* wp_lo;
    /* optional noise-add of the modulated convolution (networks.py:68-69,130-134; fma.py:15): per-pixel plane(s) already multiplied by
     * the noise strength, added after o_scale and before the bias; element (n, oy, ox) at noise[n*noise_stride_n + oy*noise_stride_y +
     * ox*noise_stride_x]; noise_stride_n = 0 broadcasts one plane over the batch.  Not with accumulate = 1. */
    const float* noise;
    int64_t      noise_stride_n, noise_stride_y, noise_stride_x;
} sgv_conv_params;

int sgv_conv2d_tf32(const sgv_conv_params* p, void* stream);

/* Which kernel variant sgv_conv2d_tf32 would launch for *p (no launch, no device work): lets tests assert that a shape really
 * exercises the variant a benchmark runs.  kernel = 1 per-tap kernel (conv_tf32.cu), 3 persistent halo-patch
 * kernel (conv_tf32_v3.cu); bn = N tile (output channels per CTA or CTA pair), mh = 128-pixel M halves per tile, cluster =
 * CTAs sharing weight slabs by TMA multicast, cta_pair = 1 when the MMAs are issued as tcgen05 cta_group::2, x3 = 1 in tf32x3 mode. */
typedef struct sgv_conv_variant { int32_t kernel, bn, mh, cluster, cta_pair, x3; } sgv_conv_variant;
int sgv_conv2d_tf32_variant(const sgv_conv_params* p, sgv_conv_variant* out);

/* Weight gradient of the same contraction (replaces aten::cudnn_convolution_backward_weight /
 * cudnn_convolution_transpose_backward_weight, conv2d_gradfix.py:140-148), with both per-sample scalings fused:
 *
 *   dw[t][o][i] += sum_{n, p}  g[n, p*g_stride + g_dy[t], ..., o] * g_scale[n, o]  *  x[n, p*x_stride + x_dy[t], ..., i] * x_scale[n, i]
 *
 * p runs over an out_h x out_w lattice per sample.  stride-1 3x3 correlation (padding 1): g unshifted, x shifted by
 * (ky-1, kx-1).  stride-2 transposed 3x3: g = gradient of the (2h+1)x(2w+1) map sampled with g_stride = 2 at offset
 * (ky, kx), x unshifted.  Out-of-range pixels read as zero.  `dw` is a float32 [ntaps, cout, cin] accumulation buffer the
 * caller zeroes (split-K partial sums are added with atomics).  Requirements: cin % 32 == 0, cout % 32 == 0.
 */
typedef struct sgv_wgrad_params {
    const float* g;            /* [n, gh, gw, cout] NHWC gradient w.r.t. the contraction output */
    const float* x;            /* [n, xh, xw, cin]  NHWC contraction input */
    float*       dw;           /* [ntaps, cout, cin] */
    int32_t n, gh, gw, xh, xw, cin, cout;
    int32_t out_h, out_w;      /* lattice summed over, per sample */
    int32_t g_stride, x_stride;
    int32_t ntaps;
    int32_t g_dy[SGV_CONV_MAX_TAPS], g_dx[SGV_CONV_MAX_TAPS];
    int32_t x_dy[SGV_CONV_MAX_TAPS], x_dx[SGV_CONV_MAX_TAPS];
    const float* g_scale;      /* [n, cout] or NULL */
    const float* x_scale;      /* [n, cin]  or NULL */
    /* optional: x is a strided VIEW [n, xh, xw, cin] of a larger NHWC tensor (element strides); all zero = dense */
    int64_t x_stride_n, x_stride_y, x_stride_x;
    /* optional: tap t accumulates into dw + dw_slot[t] * cout * cin instead of dw + t * cout * cin (use_dw_slot != 0) — lets the four
     * polyphase calls of a stride-2 transposed conv fill ONE [9][..][..] gradient buffer */
    int32_t use_dw_slot;
    int32_t dw_slot[SGV_CONV_MAX_TAPS];
    /* optional: the operand already holds TF32-representable values with its scale applied (g_scale / x_scale must then be NULL): the
     * kernel skips that operand's staging pass (scale + round in shared memory), which is its bottleneck (profiles/wgrad_ablation_r1.txt) */
    int32_t g_ready, x_ready;
    int32_t precision;         /* 0 = tf32x1, 1 = tf32x3 (three passes over the hi / lo parts of both operands; g_ready / x_ready must be 0) */
} sgv_wgrad_params;

int sgv_conv2d_wgrad_tf32(const sgv_wgrad_params* p, void* stream);

/* Variant query for the weight gradient (see sgv_conv2d_tf32_variant): kernel = 1 per-tap kernel, 2 grouped-tap kernel, 3 stacked-M kernel for 64 output
 * channels (wgrad_tf32_s64.cu); nt = N tile
 * (input channels per CTA), stages = pipeline depth, ksplit = split-K factor, passes = launches issued (3 in tf32x3 mode). */
typedef struct sgv_wgrad_variant { int32_t kernel, nt, stages, ksplit, passes; } sgv_wgrad_variant;
int sgv_conv2d_wgrad_tf32_variant(const sgv_wgrad_params* p, sgv_wgrad_variant* out);

/* ---- one-pass NHWC companions of the fused layer (csrc/layer_elementwise.cu) -------------------------------------------
 * All tensors float32; activations [n, hw, c] (NHWC, c % 4 == 0, 256 % (c/4) == 0); accumulation outputs (db, dd, ds, dwmod)
 * are ADDED to (caller zeroes them).
 *
 * sgv_modconv_act_bwd: gradient of  y = act(v + bias) * gain  w.r.t. v, from the saved OUTPUT y (bias_act.cu:69-73,133 for
 *   act = linear(1) / lrelu(3)):  dz = act'(.) * dy * gain;  db[c] += sum_{n,hw} dz;  dd[n,c] += sum_hw dz * (v)   where
 *   v = act^-1(y / gain) - bias is recovered from y (so the un-activated conv output never needs to be stored).
 *   Replaces BiasActCudaGrad + dx.sum() (bias_act.py:161-186) and the reduction in the autograd of x*dcoefs (networks.py:68-71).
 * sgv_modconv_act_bwd_rgb: the same with the ToRGB branch that reads the same activation folded in (networks.py:262-265: the block
 *   output x feeds both the next block and torgb): the incoming gradient is dy (may be NULL: last block) + sum_j dyimg[n,j,hw] *
 *   wmod[n,j,c], and dwmod[n,j,c] += sum_hw dyimg[n,j,hw] * y[n,hw,c]  (dyimg: [n,3,hw] NCHW, wmod / dwmod: [n,3,c]).  One pass
 *   instead of sgv_torgb_bwd + the autograd sum of the two gradients + sgv_modconv_act_bwd.  dyimg == NULL: plain sgv_modconv_act_bwd.
 * sgv_modconv_scale_reduce:  dx = dxs * s[n,c] (dx may be NULL or alias dxs);  ds[n,c] += sum_hw dxs * x    (networks.py:66 autograd)
 * sgv_torgb_fwd:  y[n,j,hw] = sum_c x[n,hw,c] * wmod[n,j,c] + bias[j],  j < 3, y in NCHW            (networks.py:159-163)
 * sgv_torgb_bwd:  dx[n,hw,c] = sum_j dy[n,j,hw] * wmod[n,j,c];  dwmod[n,j,c] += sum_hw dy[n,j,hw] * x[n,hw,c]
 */
int sgv_modconv_act_bwd(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                        int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream);
int sgv_modconv_act_bwd_rgb(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                            const float* dyimg, const float* wmod, float* dwmod,
                            int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream);
/* sgv_modconv_act_bwd_ex: sgv_modconv_act_bwd_rgb plus an optional output scale: when oscale [n,c] is given dz is STORED as
 * tf32_rn(dz * oscale[n,c]) (db / dd are still reduced from the unscaled dz) — the demodulation factor of the layer folded into the
 * gradient once, so that the data-gradient and weight-gradient contractions that read dz need no operand scaling or rounding. */
int sgv_modconv_act_bwd_ex(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                           const float* dyimg, const float* wmod, float* dwmod, const float* oscale,
                           int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream);
int sgv_modconv_scale_reduce(const float* dxs, const float* x, const float* s, float* dx, float* ds,
                             int32_t n, int32_t hw, int32_t c, void* stream);
int sgv_torgb_fwd(const float* x, const float* wmod, const float* bias, float* y, int32_t n, int32_t hw, int32_t c, void* stream);
/* ToRGB modulated weights (networks.py:159-160):  wmod[n, j, c] = w[j, c] * styles[n, c] * gain   (w [img_channels <= 4, c]; styles [n, c] with row
 * stride styles_stride), and both gradients in one launch:  d_styles[n, c] = gain * sum_j dwmod[n, j, c] * w[j, c]  (dense [n, c]; may be NULL),
 * dw[j, c] = gain * sum_n dwmod[n, j, c] * styles[n, c]  (may be NULL). */
int sgv_torgb_wmod_fwd(const float* w, const float* styles, int64_t styles_stride, float* wmod, int32_t n, int32_t c, int32_t img_channels,
                       float gain, void* stream);
int sgv_torgb_wmod_bwd(const float* dwmod, const float* w, const float* styles, int64_t styles_stride, float* d_styles, float* dw,
                       int32_t n, int32_t c, int32_t img_channels, float gain, void* stream);
int sgv_torgb_bwd(const float* dy, const float* x, const float* wmod, float* dx, float* dwmod, int32_t n, int32_t hw, int32_t c, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGV_B200_CONV_H */
