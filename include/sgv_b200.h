/*
 * sgv_b200 — C ABI of the B200-native (sm_100a) StyleGAN-V synthesis/discriminator hot path.
 *
 * This is the drop-in boundary: the entry points below are what the reference's two
 * pybind plugins (`_plugin.upfirdn2d`, `_plugin.bias_act`) and its library calls for the dense
 * contraction (cuDNN through torch.nn.functional.conv2d / conv_transpose2d) bind to.
 *
 * Contract (all entry points):
 *   - plain pointers and sizes only; every device buffer (inputs, outputs, workspaces) is allocated by
 *     the caller; the library never allocates, frees or synchronises, and enqueues work only on the
 *     `stream` argument (a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - the caller selects the device (cudaSetDevice) before calling — reference: OptionalCUDAGuard,
 *     upfirdn2d.cpp:31, bias_act.cpp:54;
 *   - re-entrant; global state is limited to a thread-local error string, a launch counter and per-device caches of
 *     one-time set-up (kernel attributes, occupancy, SM count) kept in atomics;
 *   - returns 0 on success, non-zero on error (sgv_last_error() describes it) — the reference raises
 *     through TORCH_CHECK / AT_CUDA_CHECK (upfirdn2d.cpp:19-28,92; bias_act.cpp:35-51,88);
 *   - sizes are int32 like the reference (numel <= INT_MAX, upfirdn2d.cpp:22-23,36; bias_act.cpp:40);
 *     strides are in ELEMENTS (int64).
 *   - there is NO CPU path: calling without a CUDA device of compute capability 10.x is an error.
 */
#ifndef SGV_B200_H
#define SGV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGV_ABI_VERSION 2

enum sgv_dtype { SGV_F32 = 0, SGV_F16 = 1, SGV_F64 = 2 };

enum sgv_status {
    SGV_OK = 0,
    SGV_ERR_INVALID = 1,      /* bad argument                                  */
    SGV_ERR_UNSUPPORTED = 2,  /* valid but not implemented by this build       */
    SGV_ERR_CUDA = 3,         /* CUDA runtime / launch failure                 */
    SGV_ERR_NO_DEVICE = 4     /* no sm_100 device (there is no CPU fallback)   */
};

/* ---- housekeeping ------------------------------------------------------------------------------ */
int         sgv_abi_version(void);
const char* sgv_last_error(void);          /* thread-local, valid until the next failing call */
int         sgv_device_check(void);        /* SGV_OK iff current device is compute capability 10.x */
int64_t     sgv_kernel_launch_count(void); /* kernels launched by this library in this process (bench "gpu_launches") */

/* ---- upfirdn2d ---------------------------------------------------------------------------------
 * Replaces `_plugin.upfirdn2d` = upfirdn2d() in src/torch_utils/ops/upfirdn2d.cpp:16-94 and the kernels of
 * upfirdn2d.cu:29-200.  Field meaning follows upfirdn2d_kernel_params (upfirdn2d.h:14-40); sizes/strides
 * are [width, height, channel, batch] like the reference's int4s.  The caller allocates `y` with
 * out = (in*up + pad0 + pad1 - fsize + down) / down (upfirdn2d.cpp:32-33; sgv_upfirdn2d_out_size).
 * Any dense/strided NCHW or channels_last layout is accepted through the strides.
 *
 * Optional fused epilogue (not in the reference plugin; used by the native synthesis layers):
 *   y = act((fir(x) * scale[n,c] + noise[n,oy,ox]) + bias[c]) * act_gain, clamp   with act in {linear(1), lrelu(3)}
 * (the order of modulated_conv2d's fma(x, dcoefs, noise) followed by bias_act: networks.py:68-69,141-143)
 * disabled when epi_scale == epi_bias == NULL and epi_act == 0.
 */
typedef struct sgv_upfirdn2d_params {
    const void*  x;
    const float* f;            /* FIR taps, float32, [f_h, f_w] with element strides f_stride_{y,x} */
    void*        y;
    int32_t dtype;             /* enum sgv_dtype of x and y */
    int32_t up_x, up_y, down_x, down_y;
    int32_t pad_x0, pad_x1, pad_y0, pad_y1;
    int32_t flip;              /* 0 = convolution, 1 = correlation (upfirdn2d.py:138) */
    float   gain;
    int32_t in_w, in_h, in_c, in_n;
    int64_t in_stride_x, in_stride_y, in_stride_c, in_stride_n;
    int32_t f_w, f_h;
    int64_t f_stride_x, f_stride_y;
    int32_t out_w, out_h;
    int64_t out_stride_x, out_stride_y, out_stride_c, out_stride_n;
    /* fused epilogue (all optional) */
    const float* epi_scale;    /* [n, c] row-major or NULL */
    const float* epi_bias;     /* [c] or NULL */
    int32_t epi_act;           /* 0 = none, 1 = linear, 3 = lrelu (bias_act cuda_idx numbering) */
    float   epi_alpha, epi_gain, epi_clamp;   /* clamp < 0 disables */
    int32_t epi_round_tf32;                   /* != 0 (needs epi_act != 0): round the result to TF32 (nearest, ties away) — lets a tensor-core
                                                 consumer skip its operand-rounding pass; channels_last 4x4 up=down=1 kernels only */
    const float* epi_noise;                   /* per-pixel noise plane(s), already multiplied by the noise strength, or NULL (needs epi_act != 0):
                                                 element (n, oy, ox) at epi_noise[n*epi_noise_stride_n + oy*epi_noise_stride_y + ox*epi_noise_stride_x];
                                                 stride_n = 0 broadcasts one plane over the batch (noise_mode='const', networks.py:133-134) */
    int64_t epi_noise_stride_n, epi_noise_stride_y, epi_noise_stride_x;
} sgv_upfirdn2d_params;

int sgv_upfirdn2d_out_size(int in_size, int up, int pad0, int pad1, int fsize, int down);
int sgv_upfirdn2d(const sgv_upfirdn2d_params* p, void* stream);

/* ---- bias_act ----------------------------------------------------------------------------------
 * Replaces `_plugin.bias_act` = bias_act() in src/torch_utils/ops/bias_act.cpp:32-90 and bias_act_kernel
 * (bias_act.cu:23-147).  Same fields as bias_act_kernel_params (bias_act.h:12-31).  NULL pointer = absent
 * (the reference passes an empty tensor, bias_act.py:39,153).  `act` is the reference's cuda_idx 1..9
 * (bias_act.py:23-33), `grad` 0/1/2, clamp < 0 disables clamping.  x, xref, yref, dy, y share one dense
 * layout; b is contiguous; bias index = (xi / step_b) % size_b (bias_act.cu:44).
 *
 * Optional fused reduction (the reference does `dx.sum(...)` in PyTorch, bias_act.py:173):
 * when `db_accum` != NULL the kernel also atomically accumulates sum over non-bias dims of y into
 * db_accum[size_b] (float32 accumulation buffer that the caller zeroed).
 */
typedef struct sgv_bias_act_params {
    const void* x;
    const void* b;
    const void* xref;
    const void* yref;
    const void* dy;
    void*       y;
    int32_t dtype;
    int32_t grad, act;
    float   alpha, gain, clamp;
    int32_t size_x, size_b, step_b;
    float*  db_accum;
} sgv_bias_act_params;

int sgv_bias_act(const sgv_bias_act_params* p, void* stream);

/* ---- dense contraction (implicit-GEMM convolution on tcgen05 tensor cores) -----------------------
 * Replaces the cuDNN calls behind conv2d_gradfix.conv2d / conv_transpose2d (conv2d_gradfix.py:35-43,
 * 108-116,140-148) for the hot-path shapes, with StyleGAN modulation (networks.py:57-74), bias,
 * activation, gain and clamp (bias_act.cu:23-147) fused into the operand load / epilogue.
 * Declared in include/sgv_b200_conv.h.
 */

#ifdef __cplusplus
}
#endif
#endif /* SGV_B200_H */
