// Elementwise tail of the continuous Fourier time-encoder (AlignedTimeEncoder.forward, src/training/motion.py:185-214)
// and its gradient — one launch each instead of ~25 (forward) / ~60 (autograd) PyTorch kernels on [M, F]-sized tensors.
//
// Thread = one (row m, frequency f) pair: it produces out[m, f] (sin half) and out[m, F + f] (cos half), so the three
// phase arguments raw(t), raw(t_L), raw(t_R) and their sin/cos are evaluated once.  Columns are the fast thread index:
// every global access of a warp is one contiguous 128-byte line.  The problem is a few thousand elements (M = frames of
// the batch, F = 256): latency-, not bandwidth-bound; what matters is that it is ONE node of the step's CUDA graph.
//
// Arithmetic follows the reference expression order with separately rounded products / sums (__fmul_rn / __fadd_rn stop
// nvcc from contracting into FMAs), so `raw` is bit-identical to the PyTorch result; sin / cos / tanh are the accurate
// libdevice versions (|raw| reaches ~800 rad; this file must NOT be built with --use_fast_math).
#include "common.cuh"
#include "aux_math.cuh"
#include "../../include/sgv_b200_aux.h"

namespace sgv {

__global__ void __launch_bounds__(256) time_encoder_fwd_kernel(const float* __restrict__ hl, const float* __restrict__ ar,
                                                               const float* __restrict__ t, const float* __restrict__ freqs,
                                                               const float* __restrict__ pscale, float* __restrict__ out,
                                                               int m, int nf, float d)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * nf) return;
    const int row = idx / nf, f = idx - row * nf;
    const TimeGeom g = time_geom(t[row], d);
    float* o = out + (size_t)row * 2 * nf;
    time_encoder_fwd_elem(hl + (size_t)row * 4 * nf, ar + (size_t)row * 2 * nf, nf, f, freqs[f], pscale[f], g, o + f, o + nf + f);
}

__global__ void __launch_bounds__(256) time_encoder_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ hl,
                                                               const float* __restrict__ t, const float* __restrict__ freqs,
                                                               const float* __restrict__ pscale, float* __restrict__ dhl,
                                                               float* __restrict__ dar, int m, int nf, float d)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= m * nf) return;
    const int row = idx / nf, f = idx - row * nf;
    const TimeGeom g = time_geom(t[row], d);
    const float gs = dout[(size_t)row * 2 * nf + f], gc = dout[(size_t)row * 2 * nf + nf + f];
    time_encoder_bwd_elem(hl + (size_t)row * 4 * nf, nf, f, freqs[f], pscale[f], g, gs, gc, dhl + (size_t)row * 4 * nf, dar + (size_t)row * 2 * nf);
}

} // namespace sgv

extern "C" int sgv_time_encoder_fwd(const float* heads_left, const float* aligners_right, const float* t,
                                    const float* freqs, const float* phase_scales, float* out,
                                    int32_t m, int32_t num_freqs, float motion_z_distance, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(heads_left && aligners_right && t && freqs && phase_scales && out, "sgv_time_encoder_fwd: NULL buffer");
    SGV_CHECK_ARG(m >= 0 && num_freqs > 0 && (int64_t)m * num_freqs * 4 <= INT32_MAX, "sgv_time_encoder_fwd: bad sizes m=%d num_freqs=%d", m, num_freqs);
    SGV_CHECK_ARG(motion_z_distance > 0.f, "sgv_time_encoder_fwd: motion_z_distance must be > 0");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    if (m == 0) return SGV_OK;
    const int total = m * num_freqs;
    time_encoder_fwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream_>>>(heads_left, aligners_right, t, freqs, phase_scales, out,
                                                                                       m, num_freqs, motion_z_distance);
    SGV_LAUNCH_OK("time_encoder_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_time_encoder_bwd(const float* dout, const float* heads_left, const float* t,
                                    const float* freqs, const float* phase_scales, float* d_heads_left, float* d_aligners_right,
                                    int32_t m, int32_t num_freqs, float motion_z_distance, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dout && heads_left && t && freqs && phase_scales && d_heads_left && d_aligners_right, "sgv_time_encoder_bwd: NULL buffer");
    SGV_CHECK_ARG(m >= 0 && num_freqs > 0 && (int64_t)m * num_freqs * 4 <= INT32_MAX, "sgv_time_encoder_bwd: bad sizes m=%d num_freqs=%d", m, num_freqs);
    SGV_CHECK_ARG(motion_z_distance > 0.f, "sgv_time_encoder_bwd: motion_z_distance must be > 0");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    if (m == 0) return SGV_OK;
    const int total = m * num_freqs;
    time_encoder_bwd_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream_>>>(dout, heads_left, t, freqs, phase_scales, d_heads_left,
                                                                                       d_aligners_right, m, num_freqs, motion_z_distance);
    SGV_LAUNCH_OK("time_encoder_bwd_kernel");
    return SGV_OK;
}
