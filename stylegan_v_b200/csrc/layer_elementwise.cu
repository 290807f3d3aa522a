// HBM-bound companion kernels of the fused synthesis layer (NHWC fp32), each one pass over the activation tensors:
//
//   sgv_modconv_act_bwd       dz = act'(y) * dy * gain;  db[c] += sum dz;  dd[n,c] += sum_hw dz * (pre - b)
//                             (pre = pre-activation recovered from the saved output y: lrelu is invertible) — replaces the
//                             reference's BiasActCudaGrad + dx.sum() (bias_act.py:161-186) + the autograd of `x * dcoefs`
//                             (networks.py:68-71) = 6 PyTorch kernels per layer.
//   sgv_modconv_scale_reduce  dx = dxs * s[n,c];  ds[n,c] += sum_hw dxs * x   — the autograd of `x * styles` (networks.py:66).
//   sgv_torgb_fwd / _bwd      ToRGB (1x1 modulated conv to 3 channels, no demodulation, networks.py:159-163): reads the
//                             NHWC activation once; fwd writes NCHW rgb, bwd writes dx and reduces d(wmod).
//
// Threading (all four): a thread owns 4 consecutive channels (one 128-bit lane) and walks pixels of ONE sample with a
// stride, keeping per-channel partial sums in registers; partials are merged through shared-memory atomics, then one
// global atomicAdd per channel per CTA.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/sgv_b200_conv.h"

namespace sgv {

constexpr int kEwThreads = 256;
constexpr int kEwUnroll = 4;

struct EwGeom
{
    int n, hw, c;          // samples, pixels per sample, channels (c % 4 == 0)
    int cvecs;             // c / 4
    int lanes;             // pixel lanes per CTA = kEwThreads / cvecs (>= 1) ; when cvecs > kEwThreads a thread loops over channel vectors
    int chunks;            // CTAs per sample
};

__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ void f4_acc(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

__device__ __forceinline__ void smem_acc4(float* s, int c0, float4 v)
{
    atomicAdd(s + c0 + 0, v.x); atomicAdd(s + c0 + 1, v.y); atomicAdd(s + c0 + 2, v.z); atomicAdd(s + c0 + 3, v.w);
}

// ---------------------------------------------------------------------------------------------------------------
struct ActBwdArgs
{
    const float* dy; const float* y; const float* bias; float* dz; float* db; float* dd;
    int act; float alpha, gain;
    // optional ToRGB branch hanging off the same activation (RGB = true): dy_total = dy (may be NULL) + sum_j dyimg[n,j,hw] * wmod[n,j,c];
    // dwmod[n,j,c] += sum_hw dyimg[n,j,hw] * y[n,hw,c]
    const float* dyimg; const float* wmod; float* dwmod;
    const float* oscale;      // optional [n, c]: dz is stored as tf32_rn(dz * oscale) (reductions use the unscaled dz)
    EwGeom g;
};

template <bool RGB>
__global__ void __launch_bounds__(kEwThreads) modconv_act_bwd_kernel(ActBwdArgs p)
{
    extern __shared__ float sacc[];            // [2 (+3)][c]: db, dd (, dwmod[0..2]) partials
    const EwGeom g = p.g;
    for (int i = threadIdx.x; i < (RGB ? 5 : 2) * g.c; i += kEwThreads) sacc[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int cv = threadIdx.x % g.cvecs, lane = threadIdx.x / g.cvecs;
    const bool active = lane < g.lanes;
    const long long base = (long long)n * g.hw * g.c;
    const float inv_gain = 1.f / p.gain, inv_alpha = 1.f / p.alpha;
    if (active)
    {
        const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 os4 = p.oscale ? __ldg(reinterpret_cast<const float4*>(p.oscale + (long long)n * g.c) + cv) : make_float4(1.f, 1.f, 1.f, 1.f);
        float4 sdb = make_float4(0.f, 0.f, 0.f, 0.f), sdd = sdb;
        float4 wm[3], sdw[3];
#pragma unroll
        for (int j = 0; j < 3; j++)
        {
            sdw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            wm[j] = RGB ? __ldg(reinterpret_cast<const float4*>(p.wmod + ((long long)n * 3 + j) * g.c) + cv) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float* gimg = RGB ? p.dyimg + (long long)n * 3 * g.hw : nullptr;
        const int stride = g.lanes * g.chunks;
        for (int p0 = chunk * g.lanes + lane; p0 < g.hw; p0 += stride * kEwUnroll)
        {
            float4 vdy[kEwUnroll], vy[kEwUnroll];
            float gi[kEwUnroll][3];
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px < g.hw)
                {
                    const long long off = base + (long long)px * g.c + cv * 4;
                    vdy[u] = (!RGB || p.dy) ? __ldcs(reinterpret_cast<const float4*>(p.dy + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    vy[u] = __ldcs(reinterpret_cast<const float4*>(p.y + off));
                    if (RGB)
                    {
#pragma unroll
                        for (int j = 0; j < 3; j++) gi[u][j] = __ldg(gimg + (long long)j * g.hw + px);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px >= g.hw) continue;
                if (RGB)
                {
#pragma unroll
                    for (int j = 0; j < 3; j++)
                    {
                        vdy[u].x = fmaf(gi[u][j], wm[j].x, vdy[u].x); vdy[u].y = fmaf(gi[u][j], wm[j].y, vdy[u].y);
                        vdy[u].z = fmaf(gi[u][j], wm[j].z, vdy[u].z); vdy[u].w = fmaf(gi[u][j], wm[j].w, vdy[u].w);
                        sdw[j].x = fmaf(gi[u][j], vy[u].x, sdw[j].x); sdw[j].y = fmaf(gi[u][j], vy[u].y, sdw[j].y);
                        sdw[j].z = fmaf(gi[u][j], vy[u].z, sdw[j].z); sdw[j].w = fmaf(gi[u][j], vy[u].w, sdw[j].w);
                    }
                }
                float d[4] = {vdy[u].x, vdy[u].y, vdy[u].z, vdy[u].w};
                float yv[4] = {vy[u].x, vy[u].y, vy[u].z, vy[u].w};
                float bb[4] = {b4.x, b4.y, b4.z, b4.w};
                float z[4], pr[4];
#pragma unroll
                for (int e = 0; e < 4; e++)
                {
                    // same arithmetic as bias_act grad=1 for lrelu/linear (bias_act.cu:69-73,133): slope chosen by sign(y / gain)
                    float pre = yv[e] * inv_gain;
                    float gsl = d[e];
                    if (p.act == 3) { const bool pos = (p.gain > 0.f) ? (yv[e] > 0.f) : (yv[e] < 0.f); gsl = pos ? d[e] : d[e] * p.alpha; pre = pos ? pre : pre * inv_alpha; }
                    z[e] = gsl * p.gain;
                    pr[e] = pre - bb[e];
                }
                const float4 z4 = make_float4(z[0], z[1], z[2], z[3]);
                float4 zo = z4;
                if (p.oscale) zo = make_float4(ptx::tf32_rn(z[0] * os4.x), ptx::tf32_rn(z[1] * os4.y), ptx::tf32_rn(z[2] * os4.z), ptx::tf32_rn(z[3] * os4.w));
                __stcs(reinterpret_cast<float4*>(p.dz + base + (long long)px * g.c + cv * 4), zo);
                f4_acc(sdb, z4);
                f4_acc(sdd, make_float4(z[0] * pr[0], z[1] * pr[1], z[2] * pr[2], z[3] * pr[3]));
            }
        }
        if (p.db) smem_acc4(sacc, cv * 4, sdb);
        if (p.dd) smem_acc4(sacc + g.c, cv * 4, sdd);
        if (RGB)
        {
#pragma unroll
            for (int j = 0; j < 3; j++) smem_acc4(sacc + (2 + j) * g.c, cv * 4, sdw[j]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < g.c; i += kEwThreads)
    {
        if (p.db) atomicAdd(p.db + i, sacc[i]);
        if (p.dd) atomicAdd(p.dd + (long long)n * g.c + i, sacc[g.c + i]);
        if (RGB)
        {
#pragma unroll
            for (int j = 0; j < 3; j++) atomicAdd(p.dwmod + ((long long)n * 3 + j) * g.c + i, sacc[(2 + j) * g.c + i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct ScaleReduceArgs
{
    const float* dxs; const float* x; const float* s; float* dx; float* ds;
    EwGeom g;
};

__global__ void __launch_bounds__(kEwThreads) modconv_scale_reduce_kernel(ScaleReduceArgs p)
{
    extern __shared__ float sacc[];            // [c]
    const EwGeom g = p.g;
    for (int i = threadIdx.x; i < g.c; i += kEwThreads) sacc[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int cv = threadIdx.x % g.cvecs, lane = threadIdx.x / g.cvecs;
    const long long base = (long long)n * g.hw * g.c;
    if (lane < g.lanes)
    {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.s + (long long)n * g.c) + cv);
        float4 sds = make_float4(0.f, 0.f, 0.f, 0.f);
        const int stride = g.lanes * g.chunks;
        for (int p0 = chunk * g.lanes + lane; p0 < g.hw; p0 += stride * kEwUnroll)
        {
            float4 vg[kEwUnroll], vx[kEwUnroll];
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px < g.hw)
                {
                    const long long off = base + (long long)px * g.c + cv * 4;
                    vg[u] = __ldcs(reinterpret_cast<const float4*>(p.dxs + off));
                    vx[u] = __ldcs(reinterpret_cast<const float4*>(p.x + off));
                }
            }
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px >= g.hw) continue;
                if (p.dx) __stcs(reinterpret_cast<float4*>(p.dx + base + (long long)px * g.c + cv * 4), f4_mul(vg[u], s4));
                f4_acc(sds, f4_mul(vg[u], vx[u]));
            }
        }
        if (p.ds) smem_acc4(sacc, cv * 4, sds);
    }
    __syncthreads();
    if (p.ds)
        for (int i = threadIdx.x; i < g.c; i += kEwThreads) atomicAdd(p.ds + (long long)n * g.c + i, sacc[i]);
}

// ---------------------------------------------------------------------------------------------------------------
struct ToRgbArgs
{
    const float* x;        // [n, hw, c] NHWC
    const float* wmod;     // [n, 3, c]
    const float* bias;     // [3] or NULL
    float* y;              // fwd: [n, 3, hw] NCHW out
    const float* dy;       // bwd: [n, 3, hw]
    float* dx;             // bwd: [n, hw, c]
    float* dwmod;          // bwd: [n, 3, c] accumulated
    EwGeom g;
};

__global__ void __launch_bounds__(kEwThreads) torgb_fwd_kernel(ToRgbArgs p)
{
    // a group of `cvecs` consecutive threads (a power of two <= 32... or a multiple of 32) owns one pixel at a time
    __shared__ float spart[kEwThreads / 32][3];
    const EwGeom g = p.g;
    const int n = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int cv = threadIdx.x % g.cvecs, lane = threadIdx.x / g.cvecs;
    const long long base = (long long)n * g.hw * g.c;
    const float4* wm = reinterpret_cast<const float4*>(p.wmod + (long long)n * 3 * g.c);
    const float4 w0 = __ldg(wm + cv), w1 = __ldg(wm + g.cvecs + cv), w2 = __ldg(wm + 2 * g.cvecs + cv);
    const float b0 = p.bias ? p.bias[0] : 0.f, b1 = p.bias ? p.bias[1] : 0.f, b2 = p.bias ? p.bias[2] : 0.f;
    const int stride = g.lanes * g.chunks;
    const int iters = (g.hw + stride - 1) / stride;               // uniform trip count: shuffles / barriers below need every thread
    int it0 = 0;
    if (g.cvecs <= 32)
    {
        // fast path (C <= 128, the high-resolution blocks): 4 pixels per trip so that four 128-bit loads are in flight per thread
        for (; it0 + 4 <= iters; it0 += 4)
        {
            float r[4][3];
            float4 v[4];
            bool ok[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const int px = chunk * g.lanes + lane + (it0 + u) * stride;
                ok[u] = px < g.hw && lane < g.lanes;
                v[u] = ok[u] ? __ldcs(reinterpret_cast<const float4*>(p.x + base + (long long)px * g.c + cv * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                r[u][0] = v[u].x * w0.x + v[u].y * w0.y + v[u].z * w0.z + v[u].w * w0.w;
                r[u][1] = v[u].x * w1.x + v[u].y * w1.y + v[u].z * w1.z + v[u].w * w1.w;
                r[u][2] = v[u].x * w2.x + v[u].y * w2.y + v[u].z * w2.z + v[u].w * w2.w;
            }
            for (int o = g.cvecs >> 1; o > 0; o >>= 1)
            {
#pragma unroll
                for (int u = 0; u < 4; u++)
                {
                    r[u][0] += __shfl_down_sync(0xffffffffu, r[u][0], o, 32);
                    r[u][1] += __shfl_down_sync(0xffffffffu, r[u][1], o, 32);
                    r[u][2] += __shfl_down_sync(0xffffffffu, r[u][2], o, 32);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (cv == 0 && ok[u])
                {
                    const int px = chunk * g.lanes + lane + (it0 + u) * stride;
                    float* yo = p.y + (long long)n * 3 * g.hw + px;
                    yo[0] = r[u][0] + b0; yo[g.hw] = r[u][1] + b1; yo[2 * g.hw] = r[u][2] + b2;
                }
        }
    }
    for (int it = it0; it < iters; it++)
    {
        const int px = chunk * g.lanes + lane + it * stride;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (px < g.hw && lane < g.lanes)
        {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(p.x + base + (long long)px * g.c + cv * 4));
            r0 = v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
            r1 = v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
            r2 = v.x * w2.x + v.y * w2.y + v.z * w2.z + v.w * w2.w;
        }
        if (g.cvecs <= 32)
        {
            // reduce across the cvecs lanes of this pixel (cvecs is a power of two here)
            for (int o = g.cvecs >> 1; o > 0; o >>= 1)
            {
                r0 += __shfl_down_sync(0xffffffffu, r0, o, 32);
                r1 += __shfl_down_sync(0xffffffffu, r1, o, 32);
                r2 += __shfl_down_sync(0xffffffffu, r2, o, 32);
            }
            if (cv == 0 && px < g.hw && lane < g.lanes)
            {
                float* yo = p.y + (long long)n * 3 * g.hw + px;
                yo[0] = r0 + b0; yo[g.hw] = r1 + b1; yo[2 * g.hw] = r2 + b2;
            }
        }
        else
        {
            // cvecs in {64, 128, 256}: one pixel spans several warps
            for (int o = 16; o > 0; o >>= 1)
            {
                r0 += __shfl_down_sync(0xffffffffu, r0, o);
                r1 += __shfl_down_sync(0xffffffffu, r1, o);
                r2 += __shfl_down_sync(0xffffffffu, r2, o);
            }
            const int w = threadIdx.x >> 5;
            if ((threadIdx.x & 31) == 0) { spart[w][0] = r0; spart[w][1] = r1; spart[w][2] = r2; }
            __syncthreads();
            if (cv == 0 && px < g.hw && lane < g.lanes)
            {
                const int wpp = g.cvecs >> 5;          // warps per pixel
                float a0 = b0, a1 = b1, a2 = b2;
                for (int k = 0; k < wpp; k++) { a0 += spart[w + k][0]; a1 += spart[w + k][1]; a2 += spart[w + k][2]; }
                float* yo = p.y + (long long)n * 3 * g.hw + px;
                yo[0] = a0; yo[g.hw] = a1; yo[2 * g.hw] = a2;
            }
            __syncthreads();
        }
    }
}

// ToRGB forward for the wide layers (C = 256 / 512 at 4x4 ... 64x64): a WARP owns a pixel, lane l covers channel vectors l, l + 32, ... (K per
// lane), four pixels in flight per warp, one shuffle reduction per pixel and no block barrier.  (The generic kernel above spreads a pixel over
// 2-4 warps and meets at two __syncthreads per pixel: 0.2 of the step's 0.35 ms of ToRGB forward went into 20 % of its bytes.)
template <int K>
__global__ void __launch_bounds__(kEwThreads) torgb_fwd_wide_kernel(ToRgbArgs p, int chunks)
{
    const EwGeom g = p.g;
    const int n = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kEwThreads / 32;
    const float4* wm = reinterpret_cast<const float4*>(p.wmod + (long long)n * 3 * g.c);
    float4 w0[K], w1[K], w2[K];
#pragma unroll
    for (int k = 0; k < K; k++) { w0[k] = __ldg(wm + lane + 32 * k); w1[k] = __ldg(wm + g.cvecs + lane + 32 * k); w2[k] = __ldg(wm + 2 * g.cvecs + lane + 32 * k); }
    const float b0 = p.bias ? p.bias[0] : 0.f, b1 = p.bias ? p.bias[1] : 0.f, b2 = p.bias ? p.bias[2] : 0.f;
    const long long base = (long long)n * g.hw * g.c;
    const int stride = kWarps * chunks;
    constexpr int U = K >= 4 ? 2 : 4;                       // pixels in flight per warp: U * K 16-byte loads per lane
    for (int p0 = chunk * kWarps + warp; p0 < g.hw; p0 += U * stride)
    {
        float4 v[U][K];
#pragma unroll
        for (int u = 0; u < U; u++)
        {
            const int px = p0 + u * stride;
#pragma unroll
            for (int k = 0; k < K; k++)
                v[u][k] = px < g.hw ? __ldcs(reinterpret_cast<const float4*>(p.x + base + (long long)px * g.c) + lane + 32 * k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++)
        {
            float r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int k = 0; k < K; k++)
            {
                r0 += v[u][k].x * w0[k].x + v[u][k].y * w0[k].y + v[u][k].z * w0[k].z + v[u][k].w * w0[k].w;
                r1 += v[u][k].x * w1[k].x + v[u][k].y * w1[k].y + v[u][k].z * w1[k].z + v[u][k].w * w1[k].w;
                r2 += v[u][k].x * w2[k].x + v[u][k].y * w2[k].y + v[u][k].z * w2[k].z + v[u][k].w * w2[k].w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                r0 += __shfl_down_sync(0xffffffffu, r0, o);
                r1 += __shfl_down_sync(0xffffffffu, r1, o);
                r2 += __shfl_down_sync(0xffffffffu, r2, o);
            }
            const int px = p0 + u * stride;
            if (lane == 0 && px < g.hw)
            {
                float* yo = p.y + (long long)n * 3 * g.hw + px;
                yo[0] = r0 + b0; yo[g.hw] = r1 + b1; yo[2 * g.hw] = r2 + b2;
            }
        }
    }
}

__global__ void __launch_bounds__(kEwThreads) torgb_bwd_kernel(ToRgbArgs p)
{
    extern __shared__ float sacc[];            // [3][c]
    const EwGeom g = p.g;
    for (int i = threadIdx.x; i < 3 * g.c; i += kEwThreads) sacc[i] = 0.f;
    __syncthreads();
    const int n = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int cv = threadIdx.x % g.cvecs, lane = threadIdx.x / g.cvecs;
    const long long base = (long long)n * g.hw * g.c;
    if (lane < g.lanes)
    {
        const float4* wm = reinterpret_cast<const float4*>(p.wmod + (long long)n * 3 * g.c);
        const float4 w0 = __ldg(wm + cv), w1 = __ldg(wm + g.cvecs + cv), w2 = __ldg(wm + 2 * g.cvecs + cv);
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
        const float* dyn = p.dy + (long long)n * 3 * g.hw;
        const int stride = g.lanes * g.chunks;
        for (int p0 = chunk * g.lanes + lane; p0 < g.hw; p0 += stride * kEwUnroll)
        {
            float4 vx[kEwUnroll]; float d0[kEwUnroll], d1[kEwUnroll], d2[kEwUnroll];
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px < g.hw)
                {
                    vx[u] = __ldcs(reinterpret_cast<const float4*>(p.x + base + (long long)px * g.c + cv * 4));
                    d0[u] = __ldg(dyn + px); d1[u] = __ldg(dyn + g.hw + px); d2[u] = __ldg(dyn + 2 * g.hw + px);
                }
            }
#pragma unroll
            for (int u = 0; u < kEwUnroll; u++)
            {
                const int px = p0 + u * stride;
                if (px >= g.hw) continue;
                float4 o;
                o.x = d0[u] * w0.x + d1[u] * w1.x + d2[u] * w2.x;
                o.y = d0[u] * w0.y + d1[u] * w1.y + d2[u] * w2.y;
                o.z = d0[u] * w0.z + d1[u] * w1.z + d2[u] * w2.z;
                o.w = d0[u] * w0.w + d1[u] * w1.w + d2[u] * w2.w;
                __stcs(reinterpret_cast<float4*>(p.dx + base + (long long)px * g.c + cv * 4), o);
                a0.x += d0[u] * vx[u].x; a0.y += d0[u] * vx[u].y; a0.z += d0[u] * vx[u].z; a0.w += d0[u] * vx[u].w;
                a1.x += d1[u] * vx[u].x; a1.y += d1[u] * vx[u].y; a1.z += d1[u] * vx[u].z; a1.w += d1[u] * vx[u].w;
                a2.x += d2[u] * vx[u].x; a2.y += d2[u] * vx[u].y; a2.z += d2[u] * vx[u].z; a2.w += d2[u] * vx[u].w;
            }
        }
        smem_acc4(sacc, cv * 4, a0); smem_acc4(sacc + g.c, cv * 4, a1); smem_acc4(sacc + 2 * g.c, cv * 4, a2);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * g.c; i += kEwThreads) atomicAdd(p.dwmod + (long long)n * 3 * g.c + i, sacc[i]);
}

static int make_geom(EwGeom* g, int n, int hw, int c)
{
    if (c % 4 != 0 || c < 4) return fail(SGV_ERR_INVALID, "channels must be a multiple of 4 (got %d)", c);
    g->n = n; g->hw = hw; g->c = c; g->cvecs = c / 4;
    if (g->cvecs > kEwThreads || (kEwThreads % g->cvecs) != 0) return fail(SGV_ERR_UNSUPPORTED, "channel count %d not supported by the layer element-wise kernels", c);
    g->lanes = kEwThreads / g->cvecs;
    // enough CTAs to fill the machine, at least ~16 pixels per lane
    int want = ceil_div(4 * num_sms(), n);
    int maxc = ceil_div(hw, g->lanes * 4);
    if (want > maxc) want = maxc;
    if (want < 1) want = 1;
    g->chunks = want;
    return SGV_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// ToRGB modulated weights (networks.py:159-160: styles = affine(w) * weight_gain; the 1x1 modulated conv without demodulation multiplies the
// [3, C] weight by them): wmod[n, j, c] = w[j, c] * s[n, c] * gain, and both gradients in ONE launch.  Per layer and step this replaces 2
// (forward) + 6 (autograd) element-wise / reduction launches on [N, 3, C]-sized tensors that sat in the serial tail of the step.
__global__ void __launch_bounds__(128) torgb_wmod_fwd_kernel(const float* __restrict__ w, const float* __restrict__ s, long long lds, float* __restrict__ wmod,
                                                             int n, int c, int j, float gain)
{
    const int ci = blockIdx.x * 128 + threadIdx.x, ni = blockIdx.y;
    if (ci >= c) return;
    const float sv = __ldg(s + (long long)ni * lds + ci) * gain;
    for (int jj = 0; jj < j; jj++) wmod[((long long)ni * j + jj) * c + ci] = __ldg(w + (long long)jj * c + ci) * sv;
}

// ds[n, c] = gain * sum_j dwmod[n, j, c] * w[j, c];  dw[j, c] = gain * sum_n dwmod[n, j, c] * s[n, c]      (j <= 4)
// CTA = 32 channels x 8 sample lanes: a thread walks samples lane, lane + 8, ...; the dw partials of the 8 lanes meet in shared memory.
// (One thread per channel walking all samples — 4 CTAs, 32 dependent trips — took 33 us per layer in the serial tail of the step.)
__global__ void __launch_bounds__(256) torgb_wmod_bwd_kernel(const float* __restrict__ dwmod, const float* __restrict__ w, const float* __restrict__ s, long long lds,
                                                             float* __restrict__ ds, float* __restrict__ dw, int n, int c, int j, float gain)
{
    __shared__ float part[8][4][32];
    const int cl = threadIdx.x & 31, nl = threadIdx.x >> 5;
    const int ci = blockIdx.x * 32 + cl;
    const bool ok = ci < c;
    float wv[4], acc[4];
#pragma unroll
    for (int jj = 0; jj < 4; jj++) { wv[jj] = (ok && jj < j) ? __ldg(w + (long long)jj * c + ci) : 0.f; acc[jj] = 0.f; }
    if (ok)
        for (int ni = nl; ni < n; ni += 8)
        {
            const float sv = __ldg(s + (long long)ni * lds + ci);
            float d = 0.f;
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
                if (jj < j)
                {
                    const float g = __ldg(dwmod + ((long long)ni * j + jj) * c + ci);
                    d = fmaf(g, wv[jj], d);
                    acc[jj] = fmaf(g, sv, acc[jj]);
                }
            if (ds) ds[(long long)ni * c + ci] = d * gain;
        }
#pragma unroll
    for (int jj = 0; jj < 4; jj++) part[nl][jj][cl] = acc[jj];
    __syncthreads();
    if (dw && ok && nl < j)
    {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; k++) t += part[k][nl][cl];
        dw[(long long)nl * c + ci] = t * gain;
    }
}

} // namespace sgv

extern "C" int sgv_modconv_act_bwd_ex(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                                      const float* dyimg, const float* wmod, float* dwmod, const float* oscale,
                                      int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream_)
{
    using namespace sgv;
    const bool rgb = dyimg != nullptr;
    SGV_CHECK_ARG(y && dz && (dy || rgb), "sgv_modconv_act_bwd: y, dz and at least one of dy / dyimg must be non-NULL");
    SGV_CHECK_ARG(!rgb || (wmod && dwmod), "sgv_modconv_act_bwd_rgb: dyimg needs wmod and dwmod");
    SGV_CHECK_ARG(act == 1 || act == 3, "act must be 1 (linear) or 3 (lrelu)");
    SGV_CHECK_ARG(gain != 0.f && (act != 3 || alpha != 0.f), "gain (and alpha for lrelu) must be non-zero");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    ActBwdArgs a;
    a.dy = dy; a.y = y; a.bias = bias; a.dz = dz; a.db = db; a.dd = dd; a.act = act; a.alpha = alpha; a.gain = gain;
    a.dyimg = dyimg; a.wmod = wmod; a.dwmod = dwmod; a.oscale = oscale;
    rc = make_geom(&a.g, n, hw, c);
    if (rc != SGV_OK) return rc;
    const unsigned grid = (unsigned)(n * a.g.chunks);
    if (rgb) modconv_act_bwd_kernel<true><<<grid, kEwThreads, 5 * c * sizeof(float), (cudaStream_t)stream_>>>(a);
    else modconv_act_bwd_kernel<false><<<grid, kEwThreads, 2 * c * sizeof(float), (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("modconv_act_bwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_modconv_act_bwd_rgb(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                                       const float* dyimg, const float* wmod, float* dwmod,
                                       int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream_)
{
    return sgv_modconv_act_bwd_ex(dy, y, bias, dz, db, dd, dyimg, wmod, dwmod, nullptr, n, hw, c, act, alpha, gain, stream_);
}

extern "C" int sgv_modconv_act_bwd(const float* dy, const float* y, const float* bias, float* dz, float* db, float* dd,
                                   int32_t n, int32_t hw, int32_t c, int32_t act, float alpha, float gain, void* stream_)
{
    SGV_CHECK_ARG(dy != nullptr, "sgv_modconv_act_bwd: dy must be non-NULL");
    return sgv_modconv_act_bwd_ex(dy, y, bias, dz, db, dd, nullptr, nullptr, nullptr, nullptr, n, hw, c, act, alpha, gain, stream_);
}

extern "C" int sgv_modconv_scale_reduce(const float* dxs, const float* x, const float* s, float* dx, float* ds,
                                        int32_t n, int32_t hw, int32_t c, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dxs && x && s, "sgv_modconv_scale_reduce: dxs, x, s must be non-NULL");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    ScaleReduceArgs a;
    a.dxs = dxs; a.x = x; a.s = s; a.dx = dx; a.ds = ds;
    rc = make_geom(&a.g, n, hw, c);
    if (rc != SGV_OK) return rc;
    modconv_scale_reduce_kernel<<<(unsigned)(n * a.g.chunks), kEwThreads, c * sizeof(float), (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("modconv_scale_reduce_kernel");
    return SGV_OK;
}

extern "C" int sgv_torgb_fwd(const float* x, const float* wmod, const float* bias, float* y, int32_t n, int32_t hw, int32_t c, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(x && wmod && y, "sgv_torgb_fwd: x, wmod, y must be non-NULL");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    ToRgbArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.wmod = wmod; a.bias = bias; a.y = y;
    rc = make_geom(&a.g, n, hw, c);
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(a.g.cvecs <= 32 ? ((a.g.cvecs & (a.g.cvecs - 1)) == 0) : (a.g.cvecs % 32 == 0), "channel count %d not supported by torgb", c);
    if (a.g.cvecs == 64 || a.g.cvecs == 128)
    {
        // wide layers: a warp per pixel; CTAs per sample so that the grid covers the SMs about four times (at least one pixel per warp)
        int chunks = ceil_div(4 * num_sms(), n);
        const int maxc = ceil_div(hw, kEwThreads / 32);
        if (chunks > maxc) chunks = maxc;
        if (chunks < 1) chunks = 1;
        if (a.g.cvecs == 64) torgb_fwd_wide_kernel<2><<<(unsigned)(n * chunks), kEwThreads, 0, (cudaStream_t)stream_>>>(a, chunks);
        else torgb_fwd_wide_kernel<4><<<(unsigned)(n * chunks), kEwThreads, 0, (cudaStream_t)stream_>>>(a, chunks);
        SGV_LAUNCH_OK("torgb_fwd_wide_kernel");
        return SGV_OK;
    }
    torgb_fwd_kernel<<<(unsigned)(n * a.g.chunks), kEwThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("torgb_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_torgb_bwd(const float* dy, const float* x, const float* wmod, float* dx, float* dwmod, int32_t n, int32_t hw, int32_t c, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dy && x && wmod && dx && dwmod, "sgv_torgb_bwd: NULL argument");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    ToRgbArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.wmod = wmod; a.dy = dy; a.dx = dx; a.dwmod = dwmod;
    rc = make_geom(&a.g, n, hw, c);
    if (rc != SGV_OK) return rc;
    torgb_bwd_kernel<<<(unsigned)(n * a.g.chunks), kEwThreads, 3 * c * sizeof(float), (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("torgb_bwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_torgb_wmod_fwd(const float* w, const float* styles, int64_t styles_stride, float* wmod, int32_t n, int32_t c, int32_t img_channels,
                                  float gain, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(w && styles && wmod && n >= 1 && c >= 1 && img_channels >= 1 && img_channels <= 4, "sgv_torgb_wmod_fwd: bad argument");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    dim3 grid((unsigned)ceil_div(c, 128), (unsigned)n);
    torgb_wmod_fwd_kernel<<<grid, 128, 0, (cudaStream_t)stream_>>>(w, styles, styles_stride, wmod, n, c, img_channels, gain);
    SGV_LAUNCH_OK("torgb_wmod_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_torgb_wmod_bwd(const float* dwmod, const float* w, const float* styles, int64_t styles_stride, float* d_styles, float* dw,
                                  int32_t n, int32_t c, int32_t img_channels, float gain, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dwmod && w && styles && (d_styles || dw) && n >= 1 && c >= 1 && img_channels >= 1 && img_channels <= 4, "sgv_torgb_wmod_bwd: bad argument");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    torgb_wmod_bwd_kernel<<<(unsigned)ceil_div(c, 32), 256, 0, (cudaStream_t)stream_>>>(dwmod, w, styles, styles_stride, d_styles, dw, n, c, img_channels, gain);
    SGV_LAUNCH_OK("torgb_wmod_bwd_kernel");
    return SGV_OK;
}
