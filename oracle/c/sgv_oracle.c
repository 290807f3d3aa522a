/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path
 * (stylegan_v_b200/...).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.
 *
 * Plain-C, scalar, CPU restatement of the two native plugins of universome/stylegan-v:
 *
 *   oracle_upfirdn2d_{f32,f64}  follows src/torch_utils/ops/upfirdn2d.cu:29-92 (upfirdn2d_kernel_large:
 *                               integer geometry lines 43-48 / 60-71, floor_div lines 20-24, inner loop
 *                               77-87, gain 90) and the output-size formula of upfirdn2d.cpp:32-33.
 *   oracle_bias_act_{f32,f64}   follows src/torch_utils/ops/bias_act.cu:23-147 (all nine activations,
 *                               grad orders 0/1/2, gain*dy at line 133, clamp at 136-142) with the
 *                               bias index arithmetic of line 44 ((xi / stepB) % sizeB).
 *
 * Float accumulation mirrors the device code: the reference is compiled by nvcc with FMA contraction
 * (and --use_fast_math), so `v += a * b` is a fused multiply-add; fmaf() reproduces that exactly.
 * Parity pinned by tests/golden/*.npz generated from the reference's own Python `impl='ref'`
 * functions by oracle/make_goldens.py (the reference's test-suite holds no vectors for this path).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

static int floor_div(int a, int b)
{
    int t = 1 - a / b;
    return (a + t * b) / b - t;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
    int up_x, up_y, down_x, down_y, pad_x0, pad_y0, flip;
    int in_w, in_h, in_c, in_n;
    int64_t in_sx, in_sy, in_sc, in_sn;     /* element strides */
    int f_w, f_h;
    int64_t f_sx, f_sy;
    int out_w, out_h;
    int64_t out_sx, out_sy, out_sc, out_sn;
} oracle_upfirdn2d_geom;

/* Output extent: upfirdn2d.cpp:32-33. */
int oracle_upfirdn2d_out_size(int in_size, int up, int pad0, int pad1, int fsize, int down)
{
    return (in_size * up + pad0 + pad1 - fsize + down) / down;
}

#define DEFINE_UPFIRDN2D(NAME, T, ACC, FMA)                                                          \
int NAME(const T* x, const float* f, T* y, const oracle_upfirdn2d_geom* g, double gain_d)            \
{                                                                                                    \
    const ACC gain = (ACC)(float)gain_d;                                                             \
    for (int n = 0; n < g->in_n; n++)                                                                \
    for (int c = 0; c < g->in_c; c++)                                                                \
    for (int outY = 0; outY < g->out_h; outY++)                                                      \
    {                                                                                                \
        int midY = outY * g->down_y + g->up_y - 1 - g->pad_y0;                                       \
        int inY = imin(imax(floor_div(midY, g->up_y), 0), g->in_h);                                  \
        int h = imin(imax(floor_div(midY + g->f_h, g->up_y), 0), g->in_h) - inY;                     \
        int filterY = midY + g->f_h - (inY + 1) * g->up_y;                                           \
        if (g->flip) filterY = g->f_h - 1 - filterY;                                                 \
        for (int outX = 0; outX < g->out_w; outX++)                                                  \
        {                                                                                            \
            int midX = outX * g->down_x + g->up_x - 1 - g->pad_x0;                                   \
            int inX = imin(imax(floor_div(midX, g->up_x), 0), g->in_w);                              \
            int w = imin(imax(floor_div(midX + g->f_w, g->up_x), 0), g->in_w) - inX;                 \
            int filterX = midX + g->f_w - (inX + 1) * g->up_x;                                       \
            if (g->flip) filterX = g->f_w - 1 - filterX;                                             \
            const T* xp = x + inX * g->in_sx + inY * g->in_sy + c * g->in_sc + n * g->in_sn;         \
            const float* fp = f + filterX * g->f_sx + filterY * g->f_sy;                             \
            int64_t stepX = (g->flip ? g->up_x : -g->up_x) * g->f_sx;                                \
            int64_t stepY = (g->flip ? g->up_y : -g->up_y) * g->f_sy;                                \
            ACC v = 0;                                                                               \
            for (int yy = 0; yy < h; yy++)                                                           \
            {                                                                                        \
                for (int xx = 0; xx < w; xx++)                                                       \
                {                                                                                    \
                    v = FMA((ACC)(*xp), (ACC)(*fp), v);                                              \
                    xp += g->in_sx;                                                                  \
                    fp += stepX;                                                                     \
                }                                                                                    \
                xp += g->in_sy - (int64_t)w * g->in_sx;                                              \
                fp += stepY - (int64_t)w * stepX;                                                    \
            }                                                                                        \
            v *= gain;                                                                               \
            y[outX * g->out_sx + outY * g->out_sy + c * g->out_sc + n * g->out_sn] = (T)v;           \
        }                                                                                            \
    }                                                                                                \
    return 0;                                                                                        \
}

DEFINE_UPFIRDN2D(oracle_upfirdn2d_f32, float, float, fmaf)
DEFINE_UPFIRDN2D(oracle_upfirdn2d_f64, double, double, fma)

/* ------------------------------------------------------------------------------------------------ */

#define DEFINE_BIAS_ACT(NAME, T, EXP, LOG)                                                           \
int NAME(const T* xin, const T* bin, const T* xrefin, const T* yrefin, const T* dyin, T* yout,       \
         int grad, int act, double alpha_d, double gain_d, double clamp_d,                           \
         int64_t sizeX, int sizeB, int64_t stepB)                                                    \
{                                                                                                    \
    const int G = grad, A = act;                                                                     \
    const T alpha = (T)(float)alpha_d, gain = (T)(float)gain_d, clamp = (T)(float)clamp_d;           \
    const T one = 1, two = 2, expRange = 80, halfExpRange = 40;                                      \
    const T seluScale = (T)1.0507009873554804934193349852946;                                        \
    const T seluAlpha = (T)1.6732632423543772848170429916717;                                        \
    if (A < 1 || A > 9) return 1;                                                                    \
    for (int64_t xi = 0; xi < sizeX; xi++)                                                           \
    {                                                                                                \
        T x = xin[xi];                                                                               \
        T b = bin ? bin[(xi / stepB) % sizeB] : 0;                                                   \
        T xref = xrefin ? xrefin[xi] : 0;                                                            \
        T yref = yrefin ? yrefin[xi] : 0;                                                            \
        T dy = dyin ? dyin[xi] : one;                                                                \
        T yy = (gain != 0) ? yref / gain : 0;                                                        \
        T y = 0;                                                                                     \
        if (G == 0) x += b; else xref += b;                                                          \
        if (A == 1) { if (G == 0) y = x; if (G == 1) y = x; }                                        \
        if (A == 2) { if (G == 0) y = (x > 0) ? x : 0; if (G == 1) y = (yy > 0) ? x : 0; }           \
        if (A == 3) { if (G == 0) y = (x > 0) ? x : x * alpha; if (G == 1) y = (yy > 0) ? x : x * alpha; } \
        if (A == 4) {                                                                                \
            if (G == 0) { T c = EXP(x); T d = one / c; y = (x < -expRange) ? -one : (x > expRange) ? one : (c - d) / (c + d); } \
            if (G == 1) y = x * (one - yy * yy);                                                     \
            if (G == 2) y = x * (one - yy * yy) * (-two * yy);                                       \
        }                                                                                            \
        if (A == 5) {                                                                                \
            if (G == 0) y = (x < -expRange) ? 0 : one / (EXP(-x) + one);                             \
            if (G == 1) y = x * yy * (one - yy);                                                     \
            if (G == 2) y = x * yy * (one - yy) * (one - two * yy);                                  \
        }                                                                                            \
        if (A == 6) {                                                                                \
            if (G == 0) y = (x >= 0) ? x : EXP(x) - one;                                             \
            if (G == 1) y = (yy >= 0) ? x : x * (yy + one);                                          \
            if (G == 2) y = (yy >= 0) ? 0 : x * (yy + one);                                          \
        }                                                                                            \
        if (A == 7) {                                                                                \
            if (G == 0) y = (x >= 0) ? seluScale * x : (seluScale * seluAlpha) * (EXP(x) - one);     \
            if (G == 1) y = (yy >= 0) ? x * seluScale : x * (yy + seluScale * seluAlpha);            \
            if (G == 2) y = (yy >= 0) ? 0 : x * (yy + seluScale * seluAlpha);                        \
        }                                                                                            \
        if (A == 8) {                                                                                \
            if (G == 0) y = (x > expRange) ? x : LOG(EXP(x) + one);                                  \
            if (G == 1) y = x * (one - EXP(-yy));                                                    \
            if (G == 2) { T c = EXP(-yy); y = x * c * (one - c); }                                   \
        }                                                                                            \
        if (A == 9) {                                                                                \
            if (G == 0) y = (x < -expRange) ? 0 : x / (EXP(-x) + one);                               \
            else {                                                                                   \
                T c = EXP(xref); T d = c + one;                                                      \
                if (G == 1) y = (xref > halfExpRange) ? x : x * c * (xref + d) / (d * d);            \
                else y = (xref > halfExpRange) ? 0 : x * c * (xref * (two - d) + two * d) / (d * d * d); \
                yref = (xref < -expRange) ? 0 : xref / (EXP(-xref) + one) * gain;                    \
            }                                                                                        \
        }                                                                                            \
        y *= gain * dy;                                                                              \
        if (clamp >= 0) {                                                                            \
            if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0) ? clamp : -clamp;               \
            else y = (yref > -clamp && yref < clamp) ? y : 0;                                        \
        }                                                                                            \
        yout[xi] = y;                                                                                \
    }                                                                                                \
    return 0;                                                                                        \
}

DEFINE_BIAS_ACT(oracle_bias_act_f32, float, expf, logf)
DEFINE_BIAS_ACT(oracle_bias_act_f64, double, exp, log)

/* ---------------------------------------------------------------------------------------------------------------
 * oracle_time_encoder_tail_f32 — plain-C restatement of the elementwise tail of the continuous Fourier time-encoder:
 * /root/reference/src/training/motion.py:111-115 (t_left = t - t % d, t_right = t_left + d, interp = (t % d) / d; Python / torch
 * remainder semantics) and :198-212 (periods = tanh(.) + 1; raw = freqs * periods * tau + phases * phase_scales for tau in
 * {t, t_left, t_right}; [sin, cos]; pos - lerp(left, right) + lerp(aligners_left, aligners_right)), every product / sum rounded to
 * fp32 on its own like the chain of PyTorch kernels (the file is built with -ffp-contract=off; volatile stores stop x87-style excess
 * precision).  heads_left [m, 4F] = [P u_L | Phi u_L | A u_L], aligners_right [m, 2F], out [m, 2F].
 */
static float r32(float v) { volatile float r = v; return r; }

void oracle_time_encoder_tail_f32(const float* heads_left, const float* aligners_right, const float* t, const float* freqs,
                                  const float* phase_scales, float* out, int m, int nf, float d)
{
    for (int row = 0; row < m; row++)
    {
        float rem = fmodf(t[row], d);                               /* torch remainder: fmod, then shifted to the sign of d */
        if (rem != 0.f && ((rem < 0.f) != (d < 0.f))) rem = r32(rem + d);
        const float t_left = r32(t[row] - rem);                     /* motion.py:111 */
        const float t_right = r32(t_left + d);                      /* :112 */
        const float a = r32(rem / d);                               /* :114 */
        const float na = r32(1.f - a);
        const float* h = heads_left + (size_t)row * 4 * nf;
        const float* ar = aligners_right + (size_t)row * 2 * nf;
        float* o = out + (size_t)row * 2 * nf;
        for (int f = 0; f < nf; f++)
        {
            const float periods = r32(tanhf(h[f]) + 1.f);           /* :198 */
            const float base = r32(freqs[f] * periods);             /* :203 freqs * periods ... */
            const float shift = r32(h[nf + f] * phase_scales[f]);   /*      ... phases * phase_scales */
            const float taus[3] = {t[row], t_left, t_right};
            float s[3], c[3];
            for (int k = 0; k < 3; k++)
            {
                const float raw = r32(r32(base * taus[k]) + shift); /* :203-205 */
                s[k] = sinf(raw); c[k] = cosf(raw);                 /* :207-209 */
            }
            const float rem_s = r32(r32(s[1] * na) + r32(s[2] * a));                        /* :212 */
            const float rem_c = r32(r32(c[1] * na) + r32(c[2] * a));
            const float add_s = r32(r32(h[2 * nf + f] * na) + r32(ar[f] * a));              /* :213 */
            const float add_c = r32(r32(h[3 * nf + f] * na) + r32(ar[nf + f] * a));
            o[f] = r32(r32(s[0] - rem_s) + add_s);                                          /* :214 */
            o[nf + f] = r32(r32(c[0] - rem_c) + add_c);
        }
    }
}
