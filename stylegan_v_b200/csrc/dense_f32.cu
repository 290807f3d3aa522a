// Exact-fp32 dense layers for the small GEMMs of the path (sm_100a, CUDA cores): FullyConnectedLayer / EqualizedLinear
// (src/training/layers.py:108-138: style affines of every synthesis layer, the mapping networks, the discriminator's dense layers, the
// time encoder's predictor heads) and the motion trajectory's EqualizedConv1d (layers.py:331-373) expressed as a GEMM over windows.
//
// Why not the tcgen05 kernels: these products have M = batch (32 ... a few hundred rows) — a quarter of one 128-row MMA tile — and their
// results feed sin / cos after a multiplication by phase scales up to 64 (motion.py:198-212), so they need round-to-nearest fp32
// accumulation; the tensor core's accumulator truncates (measured: error grows ~K, 4e-5 at K = 5632, profiles/dense_precision_r2.txt).
// The reference runs them as cuBLAS SIMT sgemm / cuDNN FFT conv1d plus separate bias_act passes.  Here:
//
//   sgv_dense_f32_fwd     y[m, n]  = act(w_gain * sum_k A[m, k] W[n, k] + b_gain * b[n]) * gain         one launch, bias / lrelu fused
//   sgv_dense_f32_dgrad   dA[m, k] += w_gain * sum_n dz[m, n] W[n, k]                                    dz = dy * gain * act'(y) formed on the fly
//   sgv_dense_f32_wgrad   dW[n, k] (+)= w_gain * sum_m dz[m, n] A[m, k];  db[n] (+)= b_gain * sum_m dz[m, n]
//
// A rows may be WINDOWS (a_row_off[m] = element offset of row m): the valid conv1d of a [B, L, C] sequence with k taps is the GEMM whose
// row (b, q) is the contiguous slice z[b, q : q + k, :] (k * C floats) against the weight re-ordered to [O, k * C] — so only the output
// positions the caller needs are computed (the motion encoder reads 2 of the 66 trajectory positions per frame).
// Column GROUPS (group_col / group_off) let one launch serve stacked layers that read different A rows: all style affines of the
// synthesis network are one forward launch, where group g = the layers reading ws[:, g, :] (networks.py:350-357).
//
// Thread mapping: lanes run along the REDUCTION dimension in the forward (coalesced 16-byte loads of both operands, 8 x 8 accumulators
// per lane, transposing warp reduction at the end) and along the OUTPUT's contiguous dimension in the two gradients (dz staged in
// shared memory, read as broadcast LDS.128).  Same contract as include/sgv_b200.h (caller-owned buffers, explicit stream, int status).
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/sgv_b200_aux.h"

namespace sgv {

constexpr int kDenseThreads = 128;
constexpr int kDenseCols = 8;       // output columns per CTA (forward) = weight rows per CTA (weight gradient)
constexpr int kDenseRows = 32;      // A rows per CTA (forward, data gradient): 4 warps x 8

struct DenseArgs
{
    const float* a; const long long* a_row_off; long long lda;
    const float* w; const float* bias; float* y; long long ldy;
    int m, n, k;
    float w_gain, b_gain; int act; float alpha, gain;
    int groups; const int* group_col; const long long* group_off;
    const float* dy; long long lddy;
    float* da; long long ldda;
    float* dw; float* db; int accumulate;
    int nsplit;
};

__device__ __forceinline__ long long dense_group_offset(const DenseArgs& p, int col)
{
    long long off = 0;
    for (int g = 0; g < p.groups; g++)
        if (col >= __ldg(p.group_col + g)) off = __ldg(p.group_off + g);
    return off;
}

__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v)
{
    acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}

// dz = dy * gain * act'(y) for the two activations the path uses (bias_act.py: the slope is recovered from the saved output)
__device__ __forceinline__ float dense_dz(const DenseArgs& p, int m, int n)
{
    float g = __ldg(p.dy + (long long)m * p.lddy + n) * p.gain;
    if (p.act == 3 && __ldg(p.y + (long long)m * p.ldy + n) <= 0.f) g *= p.alpha;
    return g;
}

// ---- forward: CTA = 8 columns x 32 rows; warp = 8 rows; lane = 4 consecutive k per 128-wide trip ----
__global__ void __launch_bounds__(kDenseThreads) dense_fwd_kernel(const DenseArgs p)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * kDenseCols;
    const int m0 = blockIdx.y * kDenseRows + warp * 8;
    if (m0 >= p.m) return;
    const long long goff = dense_group_offset(p, n0);
    long long aoff[8], woff[8];
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        const int m = min(m0 + r, p.m - 1);
        aoff[r] = (p.a_row_off ? __ldg(p.a_row_off + m) : (long long)m * p.lda) + goff;
        woff[r] = (long long)min(n0 + r, p.n - 1) * p.k;
    }
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] = 0.f;

    for (int k0 = lane * 4; k0 < p.k; k0 += 128)
    {
        float4 av[8], wv[8];
#pragma unroll
        for (int r = 0; r < 8; r++) av[r] = __ldg(reinterpret_cast<const float4*>(p.a + aoff[r] + k0));
#pragma unroll
        for (int c = 0; c < 8; c++) wv[c] = __ldg(reinterpret_cast<const float4*>(p.w + woff[c] + k0));
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c < 8; c++)
            {
                float s = acc[r][c];
                s = fmaf(av[r].x, wv[c].x, s); s = fmaf(av[r].y, wv[c].y, s); s = fmaf(av[r].z, wv[c].z, s); s = fmaf(av[r].w, wv[c].w, s);
                acc[r][c] = s;
            }
    }
    // lane l ends up with the totals of (row l / 4, column l % 4) and (row l / 4, column 4 + l % 4)
    float v0[32], v1[32];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) { v0[r * 4 + c] = acc[r][c]; v1[r * 4 + c] = acc[r][4 + c]; }
    const float t0 = ptx::warp_reduce_32x32(v0, lane);
    const float t1 = ptx::warp_reduce_32x32(v1, lane);
    const int m = m0 + (lane >> 2);
    if (m >= p.m) return;
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        const int n = n0 + (lane & 3) + 4 * h;
        if (n >= p.n) continue;
        float f = (h ? t1 : t0) * p.w_gain;
        if (p.bias) f = fmaf(__ldg(p.bias + n), p.b_gain, f);
        if (p.act == 3) f = f > 0.f ? f : f * p.alpha;
        p.y[(long long)m * p.ldy + n] = f * p.gain;
    }
}

// ---- data gradient: CTA = 32 rows x 128 k over one slice of the group's columns; lane = 4 consecutive k; atomics into dA ----
__global__ void __launch_bounds__(kDenseThreads) dense_dgrad_kernel(const DenseArgs p)
{
    __shared__ __align__(16) float dzs[kDenseRows][36];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128 + lane * 4;
    const int m0 = blockIdx.y * kDenseRows;
    const int g = (int)blockIdx.z / p.nsplit, sp = (int)blockIdx.z - g * p.nsplit;
    const int cb = p.groups ? __ldg(p.group_col + g) : 0, ce = p.groups ? __ldg(p.group_col + g + 1) : p.n;
    const long long goff = p.groups ? __ldg(p.group_off + g) : 0;
    const int chunk = ((ce - cb + p.nsplit - 1) / p.nsplit + 31) & ~31;
    const int nb = cb + sp * chunk, ne = min(ce, nb + chunk);
    if (nb >= ne) return;
    float4 acc[8];
#pragma unroll
    for (int r = 0; r < 8; r++) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool kok = k0 < p.k;
    const int srow = threadIdx.x >> 2, sseg = (threadIdx.x & 3) * 8;

    for (int nc = nb; nc < ne; nc += 32)
    {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const int m = m0 + srow, n = nc + sseg + j;
            dzs[srow][sseg + j] = (m < p.m && n < ne) ? dense_dz(p, m, n) : 0.f;
        }
        __syncthreads();
        // 8 weight rows (16-byte loads) in flight per lane; columns past the slice read as zero (their dz entries are zero as well)
#pragma unroll
        for (int j = 0; j < 32; j += 8)
        {
            if (nc + j >= ne) break;
            float4 wv[8];
#pragma unroll
            for (int jj = 0; jj < 8; jj++)
            {
                const int n = nc + j + jj;
                wv[jj] = (kok && n < ne) ? __ldg(reinterpret_cast<const float4*>(p.w + (long long)n * p.k + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
            {
                const float4 d0 = *reinterpret_cast<const float4*>(&dzs[warp * 8 + r][j]);
                const float4 d1 = *reinterpret_cast<const float4*>(&dzs[warp * 8 + r][j + 4]);
                fma4(acc[r], d0.x, wv[0]); fma4(acc[r], d0.y, wv[1]); fma4(acc[r], d0.z, wv[2]); fma4(acc[r], d0.w, wv[3]);
                fma4(acc[r], d1.x, wv[4]); fma4(acc[r], d1.y, wv[5]); fma4(acc[r], d1.z, wv[6]); fma4(acc[r], d1.w, wv[7]);
            }
        }
    }
    if (!kok) return;
#pragma unroll
    for (int r = 0; r < 8; r++)
    {
        const int m = m0 + warp * 8 + r;
        if (m >= p.m) break;
        float* dst = p.da + (long long)m * p.ldda + goff + k0;
        atomicAdd(dst + 0, acc[r].x * p.w_gain); atomicAdd(dst + 1, acc[r].y * p.w_gain);
        atomicAdd(dst + 2, acc[r].z * p.w_gain); atomicAdd(dst + 3, acc[r].w * p.w_gain);
    }
}

// ---- weight gradient: CTA = NC weight rows x 512 k; warp = 128 k; lane = 4 consecutive k; loop over all A rows ----
// NC = 16 when the reduction is long (many A rows: the conv1d-as-windows layers): every A row is then re-read by half as many CTAs — the
// kernel is bound by that L2 traffic (M = 384, K = 5632, N = 512: 92 us at NC = 8, profiles/timeline_r2m_serial.txt).
template <int NC>
__global__ void __launch_bounds__(kDenseThreads) dense_wgrad_kernel(const DenseArgs p)
{
    __shared__ __align__(16) float dzs[kDenseRows][NC];
    __shared__ long long roff[kDenseRows];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * NC;
    const int k0 = blockIdx.x * 512 + warp * 128 + lane * 4;
    const bool kok = k0 < p.k;
    const long long goff = dense_group_offset(p, n0);
    float4 acc[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;

    for (int mc = 0; mc < p.m; mc += kDenseRows)
    {
        __syncthreads();
#pragma unroll
        for (int e = 0; e < kDenseRows * NC / kDenseThreads; e++)
        {
            const int idx = threadIdx.x * (kDenseRows * NC / kDenseThreads) + e, r = idx / NC, c = idx % NC;
            const int m = mc + r, n = n0 + c;
            dzs[r][c] = (m < p.m && n < p.n) ? dense_dz(p, m, n) : 0.f;
        }
        if (threadIdx.x < kDenseRows)
        {
            const int m = min(mc + (int)threadIdx.x, p.m - 1);
            roff[threadIdx.x] = (p.a_row_off ? __ldg(p.a_row_off + m) : (long long)m * p.lda) + goff;
        }
        __syncthreads();
        const int rows = min(kDenseRows, p.m - mc);
        if (kok)
        {
#pragma unroll 4
            for (int r = 0; r < rows; r++)
            {
                const float4 av = __ldg(reinterpret_cast<const float4*>(p.a + roff[r] + k0));
#pragma unroll
                for (int c4 = 0; c4 < NC / 4; c4++)
                {
                    const float4 d = *reinterpret_cast<const float4*>(&dzs[r][c4 * 4]);
                    fma4(acc[c4 * 4 + 0], d.x, av); fma4(acc[c4 * 4 + 1], d.y, av); fma4(acc[c4 * 4 + 2], d.z, av); fma4(acc[c4 * 4 + 3], d.w, av);
                }
            }
        }
        if (p.db && blockIdx.x == 0 && warp == 0 && lane < NC)
            for (int r = 0; r < rows; r++) bsum += dzs[r][lane];
    }
    if (kok)
    {
#pragma unroll
        for (int c = 0; c < NC; c++)
        {
            if (n0 + c >= p.n) break;
            float4* dst = reinterpret_cast<float4*>(p.dw + (long long)(n0 + c) * p.k + k0);
            float4 o = make_float4(acc[c].x * p.w_gain, acc[c].y * p.w_gain, acc[c].z * p.w_gain, acc[c].w * p.w_gain);
            if (p.accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            *dst = o;
        }
    }
    if (p.db && blockIdx.x == 0 && warp == 0 && lane < NC && n0 + lane < p.n)
    {
        const float o = bsum * p.b_gain;
        p.db[n0 + lane] = p.accumulate ? p.db[n0 + lane] + o : o;
    }
}

static int dense_args(DenseArgs* a, const sgv_dense_params* p, const char* who)
{
    SGV_CHECK_ARG(p != nullptr, "%s: NULL params", who);
    SGV_CHECK_ARG(p->m >= 1 && p->n >= 1 && p->k >= 4 && p->k % 4 == 0, "%s: bad extents m=%d n=%d k=%d (k must be a multiple of 4)", who, p->m, p->n, p->k);
    SGV_CHECK_ARG(p->w != nullptr && ((uintptr_t)p->w & 15) == 0, "%s: weight pointer NULL or not 16-byte aligned", who);
    SGV_CHECK_ARG(p->act == 1 || p->act == 3, "%s: activation must be linear (1) or lrelu (3)", who);
    SGV_CHECK_ARG(p->groups >= 0 && (p->groups == 0 || (p->group_col && p->group_off)), "%s: group tables missing", who);
    memset(a, 0, sizeof(*a));
    a->a = p->a; a->a_row_off = (const long long*)p->a_row_off; a->lda = p->lda;
    a->w = p->w; a->bias = p->bias; a->y = p->y; a->ldy = p->ldy;
    a->m = p->m; a->n = p->n; a->k = p->k;
    a->w_gain = p->w_gain; a->b_gain = p->b_gain; a->act = p->act; a->alpha = p->alpha; a->gain = p->gain;
    a->groups = p->groups; a->group_col = p->group_col; a->group_off = (const long long*)p->group_off;
    a->dy = p->dy; a->lddy = p->lddy; a->da = p->da; a->ldda = p->ldda; a->dw = p->dw; a->db = p->db; a->accumulate = p->accumulate;
    a->nsplit = 1;
    return SGV_OK;
}

static int check_a(const sgv_dense_params* p, const char* who)
{
    SGV_CHECK_ARG(p->a != nullptr && ((uintptr_t)p->a & 15) == 0, "%s: A pointer NULL or not 16-byte aligned", who);
    SGV_CHECK_ARG(p->a_row_off != nullptr || p->lda % 4 == 0, "%s: lda must be a multiple of 4 (rows are read as 16-byte vectors; window / group offsets too)", who);
    return SGV_OK;
}

} // namespace sgv

extern "C" int sgv_dense_f32_fwd(const sgv_dense_params* p, void* stream_)
{
    using namespace sgv;
    DenseArgs a;
    int rc = dense_args(&a, p, "sgv_dense_f32_fwd");
    if (rc != SGV_OK) return rc;
    rc = check_a(p, "sgv_dense_f32_fwd");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(p->y != nullptr, "sgv_dense_f32_fwd: NULL output");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    dim3 grid((unsigned)ceil_div(p->n, kDenseCols), (unsigned)ceil_div(p->m, kDenseRows));
    dense_fwd_kernel<<<grid, kDenseThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("dense_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_dense_f32_dgrad(const sgv_dense_params* p, void* stream_)
{
    using namespace sgv;
    DenseArgs a;
    int rc = dense_args(&a, p, "sgv_dense_f32_dgrad");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(p->dy && p->da && ((uintptr_t)p->da & 15) == 0 && p->ldda % 4 == 0, "sgv_dense_f32_dgrad: dy / dA missing or dA not 16-byte addressable");
    SGV_CHECK_ARG(p->act == 1 || p->y != nullptr, "sgv_dense_f32_dgrad: lrelu needs the saved output y");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    const int groups = p->groups ? p->groups : 1;
    const int base = ceil_div(p->k, 128) * ceil_div(p->m, kDenseRows) * groups;
    // split the column (reduction) range until the grid covers the SMs about twice; every slice keeps >= 64 columns
    int nsplit = 1;
    const int cols_per_group = ceil_div(p->n, groups);
    while (base * nsplit < 2 * num_sms() && cols_per_group / (nsplit * 2) >= 64) nsplit *= 2;
    a.nsplit = nsplit;
    dim3 grid((unsigned)ceil_div(p->k, 128), (unsigned)ceil_div(p->m, kDenseRows), (unsigned)(groups * nsplit));
    dense_dgrad_kernel<<<grid, kDenseThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("dense_dgrad_kernel");
    return SGV_OK;
}

extern "C" int sgv_dense_f32_wgrad(const sgv_dense_params* p, void* stream_)
{
    using namespace sgv;
    DenseArgs a;
    int rc = dense_args(&a, p, "sgv_dense_f32_wgrad");
    if (rc != SGV_OK) return rc;
    rc = check_a(p, "sgv_dense_f32_wgrad");
    if (rc != SGV_OK) return rc;
    SGV_CHECK_ARG(p->dy && p->dw && ((uintptr_t)p->dw & 15) == 0, "sgv_dense_f32_wgrad: dy / dW missing or dW not 16-byte aligned");
    SGV_CHECK_ARG(p->act == 1 || p->y != nullptr, "sgv_dense_f32_wgrad: lrelu needs the saved output y");
    rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    // 16 weight rows per CTA when the reduction is long and the grid still covers the SMs (group boundaries are multiples of 8 columns, so
    // grouped launches keep 8)
    if (p->m >= 128 && p->groups == 0 && p->n % 16 == 0 && ceil_div(p->k, 512) * (p->n / 16) >= num_sms())
    {
        dim3 grid((unsigned)ceil_div(p->k, 512), (unsigned)(p->n / 16));
        dense_wgrad_kernel<16><<<grid, kDenseThreads, 0, (cudaStream_t)stream_>>>(a);
    }
    else
    {
        dim3 grid((unsigned)ceil_div(p->k, 512), (unsigned)ceil_div(p->n, kDenseCols));
        dense_wgrad_kernel<8><<<grid, kDenseThreads, 0, (cudaStream_t)stream_>>>(a);
    }
    SGV_LAUNCH_OK("dense_wgrad_kernel");
    return SGV_OK;
}
