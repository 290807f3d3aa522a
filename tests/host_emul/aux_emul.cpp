// TEST INFRASTRUCTURE: compiles stylegan_v_b200/csrc/aux_math.cuh — the per-element arithmetic the CUDA kernels of
// time_encoder.cu / optim_step.cu execute per thread — with g++ so that the CPU-only build box can check it against the
// oracle (tests/test_aux_cpu.py).  Loops mirror the kernels' (row, f) / flat indexing.  Never shipped, never called by the product.
#include <stdint.h>
#include <stddef.h>
#include "../../stylegan_v_b200/csrc/aux_math.cuh"

using namespace sgv;

extern "C" void emul_time_encoder_fwd(const float* hl, const float* ar, const float* t, const float* freqs, const float* pscale,
                                      float* out, int m, int nf, float d)
{
    for (int idx = 0; idx < m * nf; idx++)
    {
        const int row = idx / nf, f = idx - row * nf;
        const TimeGeom g = time_geom(t[row], d);
        float* o = out + (size_t)row * 2 * nf;
        time_encoder_fwd_elem(hl + (size_t)row * 4 * nf, ar + (size_t)row * 2 * nf, nf, f, freqs[f], pscale[f], g, o + f, o + nf + f);
    }
}

extern "C" void emul_time_encoder_bwd(const float* dout, const float* hl, const float* t, const float* freqs, const float* pscale,
                                      float* dhl, float* dar, int m, int nf, float d)
{
    for (int idx = 0; idx < m * nf; idx++)
    {
        const int row = idx / nf, f = idx - row * nf;
        const TimeGeom g = time_geom(t[row], d);
        const float gs = dout[(size_t)row * 2 * nf + f], gc = dout[(size_t)row * 2 * nf + nf + f];
        time_encoder_bwd_elem(hl + (size_t)row * 4 * nf, nf, f, freqs[f], pscale[f], g, gs, gc, dhl + (size_t)row * 4 * nf, dar + (size_t)row * 2 * nf);
    }
}

extern "C" void emul_adam_ema(float* p, float* g, float* m, float* v, float* pe, int64_t numel, float lr, float beta1, float beta2, float eps,
                              float ema_beta, float grad_scale, float grad_clamp, int step, int zero_grad)
{
    float step_size, bc2_sqrt;
    adam_bias_corrections(lr, beta1, beta2, (double)step, &step_size, &bc2_sqrt);
    const AdamScalars s = make_adam_scalars(beta1, beta2, eps, step_size, bc2_sqrt, ema_beta, grad_scale, grad_clamp);
    for (int64_t k = 0; k < numel; k++)
    {
        adam_one(p[k], g[k], m[k], v[k], s);
        if (pe) pe[k] = lerp_torch(p[k], pe[k], s.ema_beta);
        if (zero_grad) g[k] = 0.f;
    }
}
