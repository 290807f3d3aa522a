// Demodulation coefficients of the modulated convolution and their gradients (sm_100a, CUDA cores, exact fp32):
//
//     dcoefs[n, o] = rsqrt( sum_{i, k} (w[o, i, k] * s[n, i])^2 + eps )  =  rsqrt( sum_i s[n, i]^2 * wsq[o, i] + eps ),   wsq[o, i] = sum_k w[o, i, k]^2
//
// (src/training/networks.py:57-59: `w = weight * styles; dcoefs = (w.square().sum(dim=[2,3,4]) + 1e-8).rsqrt()` — the reference materialises
// the [N, O, I, 3, 3] tensor; written with torch ops on [O, I]-sized intermediates this was still ~18 launches per layer and step, forward +
// autograd, five of them passes over a weight-sized tensor.)  Three launches per layer instead:
//
//     sgv_demod_fwd          dcoefs from (w, s); wsq is formed on the fly from the 9 taps (w is read once, 16-byte vectors)
//     sgv_demod_bwd_styles   ds[n, i] += 2 s[n, i] * sum_o g[n, o] wsq[o, i],          g = -0.5 * dcoefs^3 * d(dcoefs)
//     sgv_demod_bwd_weight   dw[o, i, k] = 2 w[o, i, k] * sum_n g[n, o] s[n, i]^2
//
// Thread mappings are those of csrc/dense_f32.cu (lanes along the reduction dimension forward, along the output's contiguous dimension in
// the gradients).  3x3 kernels only (taps = 9): the only demodulated layers of the path (ToRGB has demodulate = False, networks.py:160).
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/sgv_b200_aux.h"

namespace sgv {

constexpr int kDemodThreads = 128;
constexpr int kDemodTaps = 9;

struct DemodArgs
{
    const float* w; const float* s; long long lds;      // w [O, I, 9]; s [N, I] with row stride lds
    float* dc; const float* ddc;                         // dcoefs [N, O] (out forward, in backward); d(loss)/d(dcoefs) [N, O]
    float* ds; long long ldds; float* dw;
    int n, o, i; float eps; int osplit;
};

// wsq of 4 consecutive input channels of one output channel: 36 contiguous weights = 9 x 16-byte loads
__device__ __forceinline__ float4 wsq4(const float* wp)
{
    float v[36];
#pragma unroll
    for (int j = 0; j < 9; j++)
    {
        const float4 t = __ldg(reinterpret_cast<const float4*>(wp) + j);
        v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; c++)
    {
        float a = 0.f;
#pragma unroll
        for (int k = 0; k < kDemodTaps; k++) a = fmaf(v[c * 9 + k], v[c * 9 + k], a);
        r[c] = a;
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

// ---- forward: CTA = 8 output channels x 32 samples; warp = 8 samples; lane = 4 consecutive input channels per 128-wide trip ----
// wsq of the CTA's 8 output channels is formed ONCE, cooperatively (every thread: I / 64 items of 9 independent 16-byte loads), into shared
// memory; the four warps then read it as LDS.128.  (First version: every warp recomputed wsq from global memory inside the reduction loop —
// 4x redundant, dependent loads: 20 us per 512 x 512 layer, 28 us for the style gradient; the kernels sit in the serial head / tail of the step.)
__global__ void __launch_bounds__(kDemodThreads) demod_fwd_kernel(const DemodArgs p)
{
    extern __shared__ __align__(16) float wsq_s[];          // [8][i]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = blockIdx.x * 8;
    const int ivecs = p.i >> 2;
    for (int item = threadIdx.x; item < 8 * ivecs; item += kDemodThreads)
    {
        const int c = item / ivecs, iv = item - c * ivecs;
        const float4 q = wsq4(p.w + ((long long)min(o0 + c, p.o - 1) * p.i + iv * 4) * kDemodTaps);
        *reinterpret_cast<float4*>(wsq_s + c * p.i + iv * 4) = q;
    }
    __syncthreads();
    const int n0 = blockIdx.y * 32 + warp * 8;
    if (n0 >= p.n) return;
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;
    for (int i0 = lane * 4; i0 < p.i; i0 += 128)
    {
        float4 s2[8], wq[8];
#pragma unroll
        for (int r = 0; r < 8; r++)
        {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.s + (long long)min(n0 + r, p.n - 1) * p.lds + i0));
            s2[r] = make_float4(t.x * t.x, t.y * t.y, t.z * t.z, t.w * t.w);
        }
#pragma unroll
        for (int c = 0; c < 8; c++) wq[c] = *reinterpret_cast<const float4*>(wsq_s + c * p.i + i0);
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++)
            {
                float a = acc[r][c];
                a = fmaf(s2[r].x, wq[c].x, a); a = fmaf(s2[r].y, wq[c].y, a); a = fmaf(s2[r].z, wq[c].z, a); a = fmaf(s2[r].w, wq[c].w, a);
                acc[r][c] = a;
            }
    }
    float v0[32], v1[32];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { v0[r * 4 + c] = acc[r][c]; v1[r * 4 + c] = acc[r][4 + c]; }
    const float t0 = ptx::warp_reduce_32x32(v0, lane);
    const float t1 = ptx::warp_reduce_32x32(v1, lane);
    const int n = n0 + (lane >> 2);
    if (n >= p.n) return;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const int o = o0 + (lane & 3) + 4 * h;
        if (o < p.o) p.dc[(long long)n * p.o + o] = rsqrtf((h ? t1 : t0) + p.eps);
    }
}

__device__ __forceinline__ float demod_g(const DemodArgs& p, int n, int o)
{
    const float d = __ldg(p.dc + (long long)n * p.o + o);
    return -0.5f * d * d * d * __ldg(p.ddc + (long long)n * p.o + o);
}

// ---- d styles: CTA = 32 samples x 128 input channels over one slice of the output channels; lane = 4 consecutive input channels ----
// per chunk of 32 output channels the CTA stages g [32 samples][32] and wsq [32][128 input channels] in shared memory cooperatively
__global__ void __launch_bounds__(kDemodThreads) demod_bwd_styles_kernel(const DemodArgs p)
{
    __shared__ __align__(16) float gs[32][36];
    __shared__ __align__(16) float wq_s[32][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ib = blockIdx.x * 128;
    const int i0 = ib + lane * 4;
    const int n0 = blockIdx.y * 32;
    const int chunk = ((p.o + p.osplit - 1) / p.osplit + 31) & ~31;
    const int ob = (int)blockIdx.z * chunk, oe = min(p.o, ob + chunk);
    if (ob >= oe) return;
    const bool iok = i0 < p.i;
    float4 acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int srow = threadIdx.x >> 2, sseg = (threadIdx.x & 3) * 8;
    for (int oc = ob; oc < oe; oc += 32)
    {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const int n = n0 + srow, o = oc + sseg + j;
            gs[srow][sseg + j] = (n < p.n && o < oe) ? demod_g(p, n, o) : 0.f;
        }
        for (int item = threadIdx.x; item < 32 * 32; item += kDemodThreads)
        {
            const int ol = item >> 5, iv = item & 31;
            const int o = oc + ol, ii = ib + iv * 4;
            const float4 q = (o < oe && ii < p.i) ? wsq4(p.w + ((long long)o * p.i + ii) * kDemodTaps) : make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(&wq_s[ol][iv * 4]) = q;
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < 32; j++)
        {
            const float4 wq = *reinterpret_cast<const float4*>(&wq_s[j][lane * 4]);
#pragma unroll
            for (int r = 0; r < 8; r++)
            {
                const float g = gs[warp * 8 + r][j];
                acc[r].x = fmaf(g, wq.x, acc[r].x); acc[r].y = fmaf(g, wq.y, acc[r].y); acc[r].z = fmaf(g, wq.z, acc[r].z); acc[r].w = fmaf(g, wq.w, acc[r].w);
            }
        }
    }
    if (!iok) return;
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        const int n = n0 + warp * 8 + r;
        if (n >= p.n) break;
        const float4 sv = __ldg(reinterpret_cast<const float4*>(p.s + (long long)n * p.lds + i0));
        float* dst = p.ds + (long long)n * p.ldds + i0;
        atomicAdd(dst + 0, 2.f * sv.x * acc[r].x); atomicAdd(dst + 1, 2.f * sv.y * acc[r].y);
        atomicAdd(dst + 2, 2.f * sv.z * acc[r].z); atomicAdd(dst + 3, 2.f * sv.w * acc[r].w);
    }
}

// ---- d weight: CTA = 8 output channels x 512 input channels; warp = 128 input channels; lane = 4 consecutive; loop over all samples;
//      the epilogue (read w, scale, write dw: 9 + 9 vector accesses per output channel) is rolled two channels at a time ----
__global__ void __launch_bounds__(kDemodThreads) demod_bwd_weight_kernel(const DemodArgs p)
{
    __shared__ __align__(16) float gs[32][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o0 = blockIdx.y * 8;
    const int i0 = blockIdx.x * 512 + warp * 128 + lane * 4;
    const bool iok = i0 < p.i;
    float4 acc[8];
#pragma unroll
    for (int c = 0; c < 8; c++) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int nc = 0; nc < p.n; nc += 32)
    {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 2; e++)
        {
            const int idx = threadIdx.x * 2 + e, r = idx >> 3, c = idx & 7;
            const int n = nc + r, o = o0 + c;
            gs[r][c] = (n < p.n && o < p.o) ? demod_g(p, n, o) : 0.f;
        }
        __syncthreads();
        const int rows = min(32, p.n - nc);
        if (iok)
        {
#pragma unroll 4
            for (int r = 0; r < rows; r++)
            {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p.s + (long long)(nc + r) * p.lds + i0));
                const float4 s2 = make_float4(t.x * t.x, t.y * t.y, t.z * t.z, t.w * t.w);
                const float4 g0 = *reinterpret_cast<const float4*>(&gs[r][0]);
                const float4 g1 = *reinterpret_cast<const float4*>(&gs[r][4]);
                const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int c = 0; c < 8; c++)
                {
                    acc[c].x = fmaf(gv[c], s2.x, acc[c].x); acc[c].y = fmaf(gv[c], s2.y, acc[c].y);
                    acc[c].z = fmaf(gv[c], s2.z, acc[c].z); acc[c].w = fmaf(gv[c], s2.w, acc[c].w);
                }
            }
        }
    }
    if (!iok) return;
#pragma unroll 2
    for (int c = 0; c < 8; c++)
    {
        if (o0 + c >= p.o) break;
        const long long base = ((long long)(o0 + c) * p.i + i0) * kDemodTaps;
        const float f[4] = {2.f * acc[c].x, 2.f * acc[c].y, 2.f * acc[c].z, 2.f * acc[c].w};
        float v[36];
#pragma unroll
        for (int j = 0; j < 9; j++)
        {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.w + base) + j);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
        }
#pragma unroll
        for (int e = 0; e < 36; e++) v[e] *= f[e / 9];
#pragma unroll
        for (int j = 0; j < 9; j++)
            reinterpret_cast<float4*>(p.dw + base)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
}

static int demod_args(DemodArgs* a, const float* w, const float* s, int64_t lds, int n, int o, int i, int taps, const char* who)
{
    SGV_CHECK_ARG(w && s, "%s: NULL argument", who);
    SGV_CHECK_ARG(taps == kDemodTaps, "%s: 3x3 kernels (9 taps) only, got %d", who, taps);
    SGV_CHECK_ARG(n >= 1 && o >= 1 && i >= 4 && i % 4 == 0 && lds % 4 == 0, "%s: bad extents n=%d o=%d i=%d (i and the row stride of s must be multiples of 4)", who, n, o, i);
    SGV_CHECK_ARG(((uintptr_t)w & 15) == 0 && ((uintptr_t)s & 15) == 0, "%s: w and s must be 16-byte aligned", who);
    memset(a, 0, sizeof(*a));
    a->w = w; a->s = s; a->lds = lds; a->n = n; a->o = o; a->i = i; a->osplit = 1;
    return SGV_OK;
}

} // namespace sgv

extern "C" int sgv_demod_fwd(const float* w, const float* styles, int64_t styles_stride, float* dcoefs, int32_t n, int32_t o, int32_t i,
                             int32_t taps, float eps, void* stream_)
{
    using namespace sgv;
    DemodArgs a;
    int rc = demod_args(&a, w, styles, styles_stride, n, o, i, taps, "sgv_demod_fwd");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(dcoefs != nullptr, "sgv_demod_fwd: NULL output");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    a.dc = dcoefs; a.eps = eps;
    dim3 grid((unsigned)ceil_div(o, 8), (unsigned)ceil_div(n, 32));
    const size_t smem = (size_t)8 * i * sizeof(float);
    SGV_CHECK_ARG(smem <= 48 * 1024, "sgv_demod_fwd: at most 1536 input channels (got %d)", i);
    demod_fwd_kernel<<<grid, kDemodThreads, smem, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("demod_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_demod_bwd_styles(const float* w, const float* styles, int64_t styles_stride, const float* dcoefs, const float* d_dcoefs,
                                    float* d_styles, int64_t d_styles_stride, int32_t n, int32_t o, int32_t i, int32_t taps, void* stream_)
{
    using namespace sgv;
    DemodArgs a;
    int rc = demod_args(&a, w, styles, styles_stride, n, o, i, taps, "sgv_demod_bwd_styles");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(dcoefs && d_dcoefs && d_styles && d_styles_stride % 4 == 0, "sgv_demod_bwd_styles: NULL argument or unaligned stride");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    a.dc = const_cast<float*>(dcoefs); a.ddc = d_dcoefs; a.ds = d_styles; a.ldds = d_styles_stride;
    const int base = ceil_div(i, 128) * ceil_div(n, 32);
    int osplit = 1;
    while (base * osplit < num_sms() && o / (osplit * 2) >= 32) osplit *= 2;
    a.osplit = osplit;
    dim3 grid((unsigned)ceil_div(i, 128), (unsigned)ceil_div(n, 32), (unsigned)osplit);
    demod_bwd_styles_kernel<<<grid, kDemodThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("demod_bwd_styles_kernel");
    return SGV_OK;
}

extern "C" int sgv_demod_bwd_weight(const float* w, const float* styles, int64_t styles_stride, const float* dcoefs, const float* d_dcoefs,
                                    float* dw, int32_t n, int32_t o, int32_t i, int32_t taps, void* stream_)
{
    using namespace sgv;
    DemodArgs a;
    int rc = demod_args(&a, w, styles, styles_stride, n, o, i, taps, "sgv_demod_bwd_weight");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(dcoefs && d_dcoefs && dw && ((uintptr_t)dw & 15) == 0, "sgv_demod_bwd_weight: NULL argument or unaligned dw");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    a.dc = const_cast<float*>(dcoefs); a.ddc = d_dcoefs; a.dw = dw;
    dim3 grid((unsigned)ceil_div(i, 512), (unsigned)ceil_div(o, 8));
    demod_bwd_weight_kernel<<<grid, kDemodThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("demod_bwd_weight_kernel");
    return SGV_OK;
}
