// Per-element arithmetic of the time-encoder tail and of the fused optimiser step, written once as host+device inline
// functions: the CUDA kernels (time_encoder.cu, optim_step.cu) call them per thread, and tests/host_emul compiles the very
// same header with g++ (-ffp-contract=off) to check the arithmetic against the oracle on the CPU-only build box.
// This header holds no loops and no memory traffic policy — it is not a CPU implementation of the ops.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define SGV_HD __host__ __device__ __forceinline__
#else
#define SGV_HD static inline
#endif

namespace sgv {

// separately rounded fp32 operations (the compiler must not contract them into FMAs)
#if defined(__CUDA_ARCH__)
SGV_HD float mul_rn_f(float a, float b) { return __fmul_rn(a, b); }
SGV_HD float add_rn_f(float a, float b) { return __fadd_rn(a, b); }
SGV_HD float sub_rn_f(float a, float b) { return __fsub_rn(a, b); }
SGV_HD float div_rn_f(float a, float b) { return __fdiv_rn(a, b); }
#else
SGV_HD float mul_rn_f(float a, float b) { volatile float r = a * b; return r; }
SGV_HD float add_rn_f(float a, float b) { volatile float r = a + b; return r; }
SGV_HD float sub_rn_f(float a, float b) { volatile float r = a - b; return r; }
SGV_HD float div_rn_f(float a, float b) { volatile float r = a / b; return r; }
#endif

// ---------------------------------------------------------------------------------------------------------------
// time encoder (motion.py:105-115, 198-212)

struct TimeGeom { float t, t_left, t_right, a, one_minus_a; };

// t % d with Python semantics (what torch's `%` computes), then the neighbour positions and the interpolation weight
// exactly as motion.py:111-115 evaluates them in fp32.
SGV_HD TimeGeom time_geom(float t, float d)
{
    TimeGeom g;
    float r = fmodf(t, d);
    if (r != 0.f && ((r < 0.f) != (d < 0.f))) r = add_rn_f(r, d);
    g.t = t;
    g.t_left = sub_rn_f(t, r);
    g.t_right = add_rn_f(g.t_left, d);
    g.a = div_rn_f(r, d);
    g.one_minus_a = sub_rn_f(1.f, g.a);
    return g;
}

struct TimePhase { float th, s0, c0, s1, c1, s2, c2; };

// h = heads_left row [4F]: P u_L | Phi u_L | A u_L (2F)
SGV_HD TimePhase time_phase(const float* h, int nf, int f, float freq, float pscale, const TimeGeom& g)
{
    TimePhase p;
    p.th = tanhf(h[f]);
    const float base = mul_rn_f(freq, add_rn_f(p.th, 1.f));
    const float shift = mul_rn_f(h[nf + f], pscale);
    const float r0 = add_rn_f(mul_rn_f(base, g.t), shift);
    const float r1 = add_rn_f(mul_rn_f(base, g.t_left), shift);
    const float r2 = add_rn_f(mul_rn_f(base, g.t_right), shift);
    p.s0 = sinf(r0); p.c0 = cosf(r0);
    p.s1 = sinf(r1); p.c1 = cosf(r1);
    p.s2 = sinf(r2); p.c2 = cosf(r2);
    return p;
}

// out_s = out[m, f], out_c = out[m, F + f]:  (emb - remove) + add, each lerp as left * (1 - a) + right * a  (motion.py:208-210)
SGV_HD void time_encoder_fwd_elem(const float* h, const float* a_r, int nf, int f, float freq, float pscale, const TimeGeom& g,
                                  float* out_s, float* out_c)
{
    const TimePhase p = time_phase(h, nf, f, freq, pscale, g);
    const float rem_s = add_rn_f(mul_rn_f(p.s1, g.one_minus_a), mul_rn_f(p.s2, g.a));
    const float rem_c = add_rn_f(mul_rn_f(p.c1, g.one_minus_a), mul_rn_f(p.c2, g.a));
    const float add_s = add_rn_f(mul_rn_f(h[2 * nf + f], g.one_minus_a), mul_rn_f(a_r[f], g.a));
    const float add_c = add_rn_f(mul_rn_f(h[3 * nf + f], g.one_minus_a), mul_rn_f(a_r[nf + f], g.a));
    *out_s = add_rn_f(sub_rn_f(p.s0, rem_s), add_s);
    *out_c = add_rn_f(sub_rn_f(p.c0, rem_c), add_c);
}

// gradient wrt the row of heads_left (dh[f], dh[F+f], dh[2F+f], dh[3F+f]) and aligners_right (da[f], da[F+f])
SGV_HD void time_encoder_bwd_elem(const float* h, int nf, int f, float freq, float pscale, const TimeGeom& g, float gs, float gc,
                                  float* dh, float* da)
{
    const TimePhase p = time_phase(h, nf, f, freq, pscale, g);
    // d out / d raw(tau): sin' = cos, cos' = -sin; the left / right embeddings enter with weights -(1 - a) and -a
    const float dr0 = gs * p.c0 - gc * p.s0;
    const float dr1 = -g.one_minus_a * (gs * p.c1 - gc * p.s1);
    const float dr2 = -g.a * (gs * p.c2 - gc * p.s2);
    const float dbase = dr0 * g.t + dr1 * g.t_left + dr2 * g.t_right;
    const float dshift = dr0 + dr1 + dr2;
    dh[f] = dbase * freq * (1.f - p.th * p.th);          // through periods = tanh(.) + 1
    dh[nf + f] = dshift * pscale;                         // through phases * phase_scales
    dh[2 * nf + f] = gs * g.one_minus_a;                  // aligners_left
    dh[3 * nf + f] = gc * g.one_minus_a;
    da[f] = gs * g.a;                                     // aligners_right
    da[nf + f] = gc * g.a;
}

// ---------------------------------------------------------------------------------------------------------------
// optimiser step (training_loop.py:381-386,392-400; torch.optim.Adam single-tensor arithmetic)

struct AdamScalars { float one_minus_b1, b2, one_minus_b2, eps, step_size, bc2_sqrt, ema_beta, grad_scale, grad_clamp; };

SGV_HD AdamScalars make_adam_scalars(float beta1, float beta2, float eps, float step_size, float bc2_sqrt, float ema_beta, float grad_scale, float grad_clamp)
{
    AdamScalars s;
    s.one_minus_b1 = 1.f - beta1; s.b2 = beta2; s.one_minus_b2 = 1.f - beta2; s.eps = eps;
    s.step_size = step_size; s.bc2_sqrt = bc2_sqrt; s.ema_beta = ema_beta; s.grad_scale = grad_scale; s.grad_clamp = grad_clamp;
    return s;
}

SGV_HD void adam_bias_corrections(float lr, float beta1, float beta2, double t, float* step_size, float* bc2_sqrt)
{
    // in double like torch.optim.Adam: step_size = lr / (1 - beta1^t), bias_correction2_sqrt = (1 - beta2^t)^0.5
    *step_size = (float)((double)lr / (1.0 - pow((double)beta1, t)));
    *bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
}

SGV_HD float sanitize_grad(float g, float scale, float clampv)
{
    g *= scale;
    if (clampv > 0.f)
    {
        if (g != g) g = 0.f;                              // nansum over a singleton dimension: nan -> 0
        g = fminf(fmaxf(g, -clampv), clampv);             // clamp(min=neginf, max=posinf) — clamps finite values too, like the reference
    }
    return g;
}

// torch.lerp(start = p, end = p_ema, weight): weight < 0.5 ? start + weight * (end - start) : end - (end - start) * (1 - weight)
SGV_HD float lerp_torch(float start, float end, float w)
{
    const float diff = end - start;
    return (w < 0.5f) ? start + w * diff : end - diff * (1.f - w);
}

SGV_HD void adam_one(float& p, float g, float& m, float& v, const AdamScalars& s)
{
    g = sanitize_grad(g, s.grad_scale, s.grad_clamp);
    m = m + (g - m) * s.one_minus_b1;                     // exp_avg.lerp_(grad, 1 - beta1)
    v = v * s.b2 + s.one_minus_b2 * g * g;                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
    const float denom = sqrtf(v) / s.bc2_sqrt + s.eps;    // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - s.step_size * (m / denom);                    // param.addcdiv_(exp_avg, denom, value = -step_size)
}

} // namespace sgv
