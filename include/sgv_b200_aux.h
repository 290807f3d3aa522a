/*
 * sgv_b200_aux.h — C ABI of the two small pieces either side of the contraction stack (same contract as sgv_b200.h:
 * caller-owned buffers, explicit stream, no allocation / synchronisation, int status + sgv_last_error()).
 *
 *   1. the elementwise tail of the continuous Fourier time-encoder (AlignedTimeEncoder.forward,
 *      src/training/motion.py:185-214) and its gradient;
 *   2. the parameter update that follows every backward pass: nan_to_num on the gradients, Adam, and the
 *      G_ema lerp (src/training/training_loop.py:381-386,392-400) as ONE pass over flat fp32 buffers.
 */
#ifndef SGV_B200_AUX_H
#define SGV_B200_AUX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Fourier time-encoder tail --------------------------------------------------------------------
 * Replaces ~25 PyTorch kernels of AlignedTimeEncoder.forward (motion.py:198-212).  With F = num_freqs and, per row m,
 *   heads_left[m]     = [P u_L | Phi u_L | A u_L]  (F + F + 2F columns: the three bias-free predictors of motion.py:174-180
 *                       applied to the left neighbour code u_L as one stacked GEMM by the caller)
 *   aligners_right[m] = A u_R                       (2F columns)
 *   r = t mod d (python remainder), t_L = t - r, t_R = t_L + d, a = r / d          (motion.py:105-115)
 *   raw(tau)[f] = freqs[f] * (tanh(P u_L)[f] + 1) * tau + (Phi u_L)[f] * phase_scales[f]          (motion.py:201-203)
 *   out[m, f]     = sin raw(t) - lerp(sin raw(t_L), sin raw(t_R), a) + lerp(A u_L, A u_R, a)[f]
 *   out[m, F + f] = cos ...                                                    + lerp(...)[F + f]  (motion.py:205-212)
 * Arithmetic: every product / sum is rounded separately (no FMA contraction) in the reference's operation order, so
 * `raw` is bit-identical to the PyTorch expression; sin/cos/tanh are the accurate libdevice functions (arguments reach
 * hundreds of radians — no fast-math intrinsics).
 * The gradient entry point recomputes sin/cos and returns d(heads_left) [m, 4F] and d(aligners_right) [m, 2F]; t is not
 * differentiated (the reference never asks for it).
 */
int sgv_time_encoder_fwd(const float* heads_left, const float* aligners_right, const float* t,
                         const float* freqs, const float* phase_scales, float* out,
                         int32_t m, int32_t num_freqs, float motion_z_distance, void* stream);
int sgv_time_encoder_bwd(const float* dout, const float* heads_left, const float* t,
                         const float* freqs, const float* phase_scales, float* d_heads_left, float* d_aligners_right,
                         int32_t m, int32_t num_freqs, float motion_z_distance, void* stream);

/* ---- fused optimiser step ---------------------------------------------------------------------------
 * One pass over flat fp32 buffers of `numel` elements (all parameters of a module laid out back to back, gradients
 * in the matching flat buffer that the NCCL all-reduce ran on):
 *   g      = clamp(nan->0 (grad * grad_scale), -grad_clamp, +grad_clamp)         misc.nan_to_num(nan=0, posinf=1e5, neginf=-1e5)
 *                                                                                 = clamp(nansum) (torch_utils/misc.py:49-56)
 *   m      = m + (g - m) * (1 - beta1)                                           torch.optim.Adam (training_loop.py:386,
 *   v      = v * beta2 + (1 - beta2) * g * g                                      opt_kwargs train.py:192-193: betas (0, 0.99), eps 1e-8)
 *   p      = p - step_size * m / (sqrt(v) / bias_correction2_sqrt + eps)
 *   p_ema  = lerp(p, p_ema, ema_beta)   (optional; torch.lerp's two-sided formula)  training_loop.py:392-400
 *   grad   = 0                          (optional; replaces zero_grad)
 * with step_size = lr / (1 - beta1^t) and bias_correction2_sqrt = sqrt(1 - beta2^t) evaluated in double like torch does.
 * The step number t comes either by value (`step`, host-side counter) or — for launches that sit inside a replayed CUDA
 * graph — from the DEVICE counter `step_count`: with `advance_step` != 0 a one-thread kernel increments it first, so each
 * replay is the next optimiser step without any host involvement.
 * Buffers must be 16-byte aligned.  HBM traffic per parameter: 5 loads + 4 stores of 4 B with EMA and gradient zeroing.
 */
typedef struct sgv_adam_params {
    float*       param;
    float*       grad;
    float*       exp_avg;
    float*       exp_avg_sq;
    float*       param_ema;              /* NULL = no EMA update */
    int64_t      numel;
    float        lr, beta1, beta2, eps;
    float        ema_beta;
    float        grad_scale;             /* 1 = none; e.g. 1/world_size after a sum all-reduce */
    float        grad_clamp;             /* <= 0 disables the nan_to_num + clamp */
    int32_t      step;                   /* t >= 1, used when step_count == NULL */
    int32_t*     step_count;             /* optional device counter holding t (see above) */
    int32_t      advance_step;           /* != 0: ++*step_count before the update */
    int32_t      zero_grad;              /* != 0: store zeros to grad */
} sgv_adam_params;

int sgv_adam_ema_step(const sgv_adam_params* p, void* stream);

/* ---- discriminator-side companions (csrc/disc_ops.cu) -------------------------------------------------------------------------
 * sgv_fromrgb_fwd: the 1x1 `fromrgb` convolution of the first discriminator block with its bias / activation fused
 *   (src/training/networks.py:447-449,467-470; layers.py:184-197 with <= 4 input channels — a streaming pass, not a tensor-core
 *   contraction):   y[n, hw, c] = act(sum_j img[n, j, hw] * w[c, j] * weight_gain + bias[c]) * gain
 *   img [n, img_channels, hw] NCHW (as the frames arrive), w [c, img_channels], y [n, hw, c] NHWC (the layout of the conv that follows);
 *   c = 4 * (a power of two <= 32); act 1 = linear, 3 = lrelu.
 * sgv_fromrgb_bwd: from dz = d(loss)/d(pre-activation) [n, hw, c] (sgv_modconv_act_bwd on the saved output gives it, and d bias):
 *   dimg[n, j, hw] = weight_gain * sum_c dz * w[c, j]   (dimg may be NULL);   dw[c, j] += weight_gain * sum_{n,hw} dz * img   (caller zeroes dw).
 */
int sgv_fromrgb_fwd(const float* img, const float* w, const float* bias, float* y, int32_t n, int32_t hw, int32_t c, int32_t img_channels,
                    float weight_gain, int32_t act, float alpha, float gain, void* stream);
int sgv_fromrgb_bwd(const float* dz, const float* img, const float* w, float* dimg, float* dw, int32_t n, int32_t hw, int32_t c,
                    int32_t img_channels, float weight_gain, void* stream);

/* sgv_mbstd_fwd: MinibatchStdLayer (networks.py:492-516) + the channel concat + zero padding of the channel count to `cpad`:
 *   samples are grouped as n = g * M + m (G = group, M = n / G); with c1 = c / num_channels
 *     s[m, f]            = mean_{c' < c1, p} sqrt( var_g x[g*M+m, f*c1+c', p] + 1e-8 )
 *     y[n, p, 0:c]       = x[n, 0:c, p];   y[n, p, c+f] = s[n % M, f];   y[n, p, c+F:cpad] = 0          (y is NHWC [n, hw, cpad])
 *   x is addressed through element strides (sample, channel, pixel), so NCHW and NHWC inputs both work; sd_mean [M, F] receives s.
 * sgv_mbstd_bwd: dx[n, p, c] (NHWC [n, hw, c]) = dy[n, p, c] + (sum_{g,p} dy[g*M+m, p, C+f]) / (c1 * hw) * (x - mean_g x) / (G * sd).
 */
int sgv_mbstd_fwd(const float* x, int64_t x_stride_n, int64_t x_stride_c, int64_t x_stride_p, float* y, float* sd_mean,
                  int32_t n, int32_t c, int32_t hw, int32_t cpad, int32_t group, int32_t num_channels, void* stream);
int sgv_mbstd_bwd(const float* dy, const float* x, int64_t x_stride_n, int64_t x_stride_c, int64_t x_stride_p, float* dx,
                  int32_t n, int32_t c, int32_t hw, int32_t cpad, int32_t group, int32_t num_channels, void* stream);

/* ---- exact-fp32 dense layers (csrc/dense_f32.cu) --------------------------------------------------------------------------------
 * Replace torch.addmm / matmul (+ bias_act) of FullyConnectedLayer / EqualizedLinear (src/training/layers.py:108-138: style affines,
 * mapping networks, discriminator dense layers, time-encoder heads) and F.conv1d of EqualizedConv1d (layers.py:331-373; as a GEMM over
 * windows).  Plain fp32 FMAs with round-to-nearest accumulation (CUDA cores): these GEMMs have M = batch rows and feed sin / cos.
 *
 *   fwd     y[m, n]   = act(w_gain * sum_k A[m, k] * w[n, k] + b_gain * bias[n]) * gain            act: 1 linear, 3 lrelu(alpha)
 *   dgrad   dA[m, k] += w_gain * sum_n dz[m, n] * w[n, k]        (atomic adds: the caller zero-fills dA or passes a running gradient)
 *   wgrad   dW[n, k]  = w_gain * sum_m dz[m, n] * A[m, k];   db[n] = b_gain * sum_m dz[m, n]      (accumulate != 0: added to dW / db)
 *           with dz[m, n] = dy[m, n] * gain * (act == lrelu && y[m, n] <= 0 ? alpha : 1)   — the gradients read dy and the saved y.
 *
 * Addressing of A: row m starts at  a + (a_row_off ? a_row_off[m] : m * lda) + group offset.  a_row_off (device array, elements) makes rows
 * overlapping WINDOWS: the valid conv1d of z [B, L, C] with kt taps is this GEMM with row (b, q) -> offset (b * L + q) * C, k = kt * C
 * and the weight re-ordered to [n, kt * C].  Column groups (device arrays group_col [groups + 1], ascending multiples of 8, and group_off
 * [groups], elements): columns [group_col[g], group_col[g+1]) read A (fwd, wgrad) / write dA (dgrad) displaced by group_off[g] — one
 * launch for all style affines of the synthesis network, group g = the layers fed by ws[:, g, :] (networks.py:350-357).
 * Every row / group offset, lda, ldda and k are multiples of 4 elements and a, w, dA, dW 16-byte aligned (16-byte vector loads).
 */
typedef struct sgv_dense_params {
    const float*   a;                /* A operand (fwd, wgrad) */
    const int64_t* a_row_off;        /* NULL, or [m] element offsets of the rows (device memory) */
    int64_t        lda;
    const float*   w;                /* [n, k] row-major */
    const float*   bias;             /* [n] or NULL */
    float*         y;                /* [m, n] with row stride ldy: output of fwd, saved output for the gradients (lrelu) */
    int64_t        ldy;
    int32_t        m, n, k;
    float          w_gain, b_gain;
    int32_t        act;
    float          alpha, gain;
    int32_t        groups;           /* 0 = none */
    const int32_t* group_col;
    const int64_t* group_off;
    const float*   dy;               /* gradients: d(loss)/dy, row stride lddy */
    int64_t        lddy;
    float*         da;               /* dgrad output, row stride ldda */
    int64_t        ldda;
    float*         dw;               /* wgrad output [n, k] */
    float*         db;               /* wgrad: [n] or NULL */
    int32_t        accumulate;
} sgv_dense_params;

int sgv_dense_f32_fwd(const sgv_dense_params* p, void* stream);
int sgv_dense_f32_dgrad(const sgv_dense_params* p, void* stream);
int sgv_dense_f32_wgrad(const sgv_dense_params* p, void* stream);

/* ---- demodulation coefficients (csrc/demod.cu) -----------------------------------------------------------------------------------
 * dcoefs[n, o] = rsqrt(sum_{i, k} (w[o, i, k] * styles[n, i])^2 + eps)     (src/training/networks.py:57-59), evaluated as
 * rsqrt(sum_i styles[n, i]^2 * wsq[o, i] + eps) with wsq[o, i] = sum_k w[o, i, k]^2 formed on the fly; w [o, i, taps] (taps = 9 only),
 * styles [n, i] with row stride styles_stride (elements, multiple of 4), dcoefs [n, o] dense; exact fp32.
 * Gradients from d_dcoefs = d(loss)/d(dcoefs), with g = -0.5 * dcoefs^3 * d_dcoefs:
 *   sgv_demod_bwd_styles   d_styles[n, i] += 2 * styles[n, i] * sum_o g[n, o] * wsq[o, i]        (atomic adds into the caller's buffer)
 *   sgv_demod_bwd_weight   dw[o, i, k]     = 2 * w[o, i, k] * sum_n g[n, o] * styles[n, i]^2     (dw dense, same shape as w)
 */
int sgv_demod_fwd(const float* w, const float* styles, int64_t styles_stride, float* dcoefs, int32_t n, int32_t o, int32_t i, int32_t taps,
                  float eps, void* stream);
int sgv_demod_bwd_styles(const float* w, const float* styles, int64_t styles_stride, const float* dcoefs, const float* d_dcoefs,
                         float* d_styles, int64_t d_styles_stride, int32_t n, int32_t o, int32_t i, int32_t taps, void* stream);
int sgv_demod_bwd_weight(const float* w, const float* styles, int64_t styles_stride, const float* dcoefs, const float* d_dcoefs,
                         float* dw, int32_t n, int32_t o, int32_t i, int32_t taps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGV_B200_AUX_H */
