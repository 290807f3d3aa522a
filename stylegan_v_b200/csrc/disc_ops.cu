// Discriminator-side companions of the NHWC contraction kernels (sm_100a), all HBM-bound single passes:
//
//   sgv_fromrgb_fwd / _bwd   the 1x1 `fromrgb` layer of the first discriminator block (src/training/networks.py:447-449,467-470 ->
//                            layers.py:184-197 with 3 input channels): y[n,hw,c] = act(sum_j img[n,j,hw] * w[c,j] * wgain + b[c]) * gain.
//                            With 3 input channels this is not a tensor-core contraction but a 12-FMA-per-output streaming pass (the
//                            adjoint of the generator's ToRGB, csrc/layer_elementwise.cu); the reference runs it as a cuDNN conv on an
//                            NCHW tensor plus a separate bias_act pass.  The frames stay NCHW (as the data loader / generator deliver
//                            them), the 64-channel result is written NHWC, i.e. in the layout of the tcgen05 conv that follows.
//   sgv_mbstd_fwd / _bwd     MinibatchStdLayer (networks.py:492-516) fused with the channel concat that follows it and with the zero
//                            padding of the channel count to a multiple of 64 (513 -> 576), so that the epilogue's 3x3 convolution
//                            runs on the tcgen05 kernel instead of the library.
//
// Same contract as include/sgv_b200.h (caller-owned buffers, no allocation, no synchronisation, explicit stream, int status).
#include <string.h>
#include "common.cuh"
#include "../../include/sgv_b200_aux.h"

namespace sgv {

constexpr int kDiscThreads = 256;

struct FromRgbArgs
{
    const float* img; const float* w; const float* bias; float* y;          // fwd
    const float* dz; float* dimg; float* dw;                                 // bwd
    int n, hw, c, j;                                                         // j = image channels (<= 4)
    float wgain; int act; float alpha, gain;
    int cvecs, lanes;                                                        // threads per pixel (c / 4), pixels per block iteration
};

// thread = (pixel lane, 4 consecutive channels); a block walks pixels [chunk * lanes + lane + it * stride)
__global__ void __launch_bounds__(kDiscThreads) fromrgb_fwd_kernel(FromRgbArgs p)
{
    const int cv = threadIdx.x % p.cvecs, lane = threadIdx.x / p.cvecs;
    if (lane >= p.lanes) return;
    const int c0 = cv * 4;
    float wr[4][4];
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) wr[e][j] = j < p.j ? __ldg(p.w + (long long)(c0 + e) * p.j + j) * p.wgain : 0.f;
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + c0));
    const long long total = (long long)p.n * p.hw;
    const long long stride = (long long)gridDim.x * p.lanes;
    // four pixels per trip: the image loads of all four are issued before any store (the kernel is a pure stream: 12 B in, 4 * c B out per pixel)
    for (long long px0 = (long long)blockIdx.x * p.lanes + lane; px0 < total; px0 += 4 * stride)
    {
        float x[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const long long px = px0 + u * stride;
            const bool ok = px < total;
            const int n = ok ? (int)(px / p.hw) : 0, q = ok ? (int)(px - (long long)n * p.hw) : 0;
            const float* ip = p.img + (long long)n * p.j * p.hw + q;
#pragma unroll
            for (int j = 0; j < 4; j++) x[u][j] = (ok && j < p.j) ? __ldg(ip + (long long)j * p.hw) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const long long px = px0 + u * stride;
            if (px >= total) continue;
            float o[4] = {b.x, b.y, b.z, b.w};
            float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int e = 0; e < 4; e++) v[e] = fmaf(x[u][j], wr[e][j], v[e]);
#pragma unroll
            for (int e = 0; e < 4; e++)
            {
                float f = v[e] + o[e];
                if (p.act == 3) f = f > 0.f ? f : f * p.alpha;
                o[e] = f * p.gain;
            }
            __stcs(reinterpret_cast<float4*>(p.y + px * p.c + c0), make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// dimg[n,j,hw] = wgain * sum_c dz[n,hw,c] * w[c,j]   (optional);   dw[c,j] += wgain * sum_{n,hw} dz[n,hw,c] * img[n,j,hw]
__global__ void __launch_bounds__(kDiscThreads) fromrgb_bwd_kernel(FromRgbArgs p)
{
    extern __shared__ float sred[];                    // [lanes][cvecs][4 channels][4 j] partial weight gradients of the block
    const int cv = threadIdx.x % p.cvecs, lane = threadIdx.x / p.cvecs;
    const int c0 = cv * 4;
    const bool active = lane < p.lanes;
    float wr[4][4], acc[4][4];
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) { wr[e][j] = (active && j < p.j) ? __ldg(p.w + (long long)(c0 + e) * p.j + j) * p.wgain : 0.f; acc[e][j] = 0.f; }
    const long long total = (long long)p.n * p.hw;
    const long long stride = (long long)gridDim.x * p.lanes;
    const long long iters = (total + stride - 1) / stride;             // uniform trip count: the shuffles below need whole warps
    for (long long it0 = 0; it0 < iters; it0 += 4)
    {
        // four pixels per trip: all loads (4 x 128-bit gradient + 4 x j image scalars) are in flight before the arithmetic starts
        float4 d[4]; float x[4][4]; bool ok[4]; long long pxs[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            pxs[u] = (long long)blockIdx.x * p.lanes + lane + (it0 + u) * stride;
            ok[u] = active && (it0 + u) < iters && pxs[u] < total;
            d[u] = ok[u] ? __ldcs(reinterpret_cast<const float4*>(p.dz + pxs[u] * p.c + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const int n = ok[u] ? (int)(pxs[u] / p.hw) : 0, q = ok[u] ? (int)(pxs[u] - (long long)n * p.hw) : 0;
            const float* ip = p.img + (long long)n * p.j * p.hw + q;
#pragma unroll
            for (int j = 0; j < 4; j++) x[u][j] = (ok[u] && j < p.j) ? __ldg(ip + (long long)j * p.hw) : 0.f;
        }
        float r[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
            for (int j = 0; j < 4; j++)
            {
                r[u][j] = 0.f;
#pragma unroll
                for (int e = 0; e < 4; e++) { acc[e][j] = fmaf(dv[e], x[u][j], acc[e][j]); r[u][j] = fmaf(dv[e], wr[e][j], r[u][j]); }
            }
        }
        if (p.dimg)
        {
            // reduce over the cvecs threads of a pixel (cvecs is a power of two <= 32: a pixel's threads sit in one warp)
            for (int o = p.cvecs >> 1; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int j = 0; j < 4; j++) r[u][j] += __shfl_down_sync(0xffffffffu, r[u][j], o, 32);
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (ok[u] && cv == 0)
                {
                    const int n = (int)(pxs[u] / p.hw), q = (int)(pxs[u] - (long long)n * p.hw);
                    float* op = p.dimg + (long long)n * p.j * p.hw + q;
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (j < p.j) op[(long long)j * p.hw] = r[u][j];
                }
        }
    }
    // block reduction of the weight-gradient partials over the pixel lanes, then one atomic per (c, j)
    float* mine = sred + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int e = 0; e < 4; e++)
#pragma unroll
        for (int j = 0; j < 4; j++) mine[e * 4 + j] = active ? acc[e][j] : 0.f;
    __syncthreads();
    for (int idx = threadIdx.x; idx < p.cvecs * 16; idx += blockDim.x)
    {
        const int v = idx / 16, ej = idx % 16;
        float s = 0.f;
        for (int l = 0; l < p.lanes; l++) s += sred[(size_t)(l * p.cvecs + v) * 16 + ej];
        const int e = ej >> 2, j = ej & 3;
        if (j < p.j) atomicAdd(p.dw + (long long)(v * 4 + e) * p.j + j, s * p.wgain);
    }
}

static int fromrgb_geom(FromRgbArgs* a, int n, int hw, int c, int j)
{
    SGV_CHECK_ARG(n >= 1 && hw >= 1, "extents must be positive");
    SGV_CHECK_ARG(j >= 1 && j <= 4, "image channels must be in [1, 4] (got %d)", j);
    const int cvecs = c / 4;
    SGV_CHECK_ARG(c % 4 == 0 && cvecs >= 1 && cvecs <= 32 && (cvecs & (cvecs - 1)) == 0, "output channels must be 4 * a power of two <= 128 (got %d)", c);
    SGV_CHECK_ARG((long long)n * hw * c <= 0x7fffffffLL, "tensor too large");
    a->n = n; a->hw = hw; a->c = c; a->j = j; a->cvecs = cvecs; a->lanes = kDiscThreads / cvecs;
    return SGV_OK;
}

// ---- minibatch standard deviation + concat + channel padding ------------------------------------------------------------------
struct MbstdArgs
{
    const float* x; float* y; float* sd_mean;            // fwd: x [N, C, HW] via strides, y [N, HW, Cpad] NHWC, sd_mean [M, F]
    const float* dy; float* dx;                          // bwd: dy [N, HW, Cpad], dx [N, HW, C] NHWC
    long long xs_n, xs_c, xs_p;                          // element strides of x (sample, channel, pixel)
    int n, c, hw, cpad, G, M, F;
};

// one block per (m, f): sd[m,f,c',p] = sqrt(var_g x[g*M+m, f*c1+c', p] + 1e-8); s[m,f] = mean_{c',p} sd; also copies x into y[..., :C]
__global__ void __launch_bounds__(kDiscThreads) mbstd_fwd_kernel(MbstdArgs p)
{
    __shared__ float swarp[kDiscThreads / 32];
    __shared__ float s_bcast;
    const int m = blockIdx.x / p.F, f = blockIdx.x % p.F;
    const int c1 = p.c / p.F;
    const int elems = c1 * p.hw;
    float sum = 0.f;
    for (int i = threadIdx.x; i < elems; i += blockDim.x)
    {
        const int cc = f * c1 + i / p.hw, px = i % p.hw;
        float mu = 0.f;
        for (int g = 0; g < p.G; g++) mu += __ldg(p.x + (long long)(g * p.M + m) * p.xs_n + (long long)cc * p.xs_c + (long long)px * p.xs_p);
        mu /= (float)p.G;
        float var = 0.f;
        for (int g = 0; g < p.G; g++)
        {
            const float v = __ldg(p.x + (long long)(g * p.M + m) * p.xs_n + (long long)cc * p.xs_c + (long long)px * p.xs_p);
            var += (v - mu) * (v - mu);
            p.y[((long long)(g * p.M + m) * p.hw + px) * p.cpad + cc] = v;           // the concat's copy of x, into the NHWC result
        }
        sum += sqrtf(var / (float)p.G + 1e-8f);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        float t = 0.f;
        for (int w = 0; w < kDiscThreads / 32; w++) t += swarp[w];
        s_bcast = t / (float)elems;
        p.sd_mean[m * p.F + f] = s_bcast;
    }
    __syncthreads();
    const float s = s_bcast;
    // the statistic as channel C + f of every sample of the group column m, and zeros in the padding channels (written by the f = 0 block)
    for (int i = threadIdx.x; i < p.G * p.hw; i += blockDim.x)
    {
        const int g = i / p.hw, px = i % p.hw;
        float* row = p.y + ((long long)(g * p.M + m) * p.hw + px) * p.cpad;
        row[p.c + f] = s;
        if (f == 0)
            for (int z = p.c + p.F; z < p.cpad; z++) row[z] = 0.f;
    }
}

// dx[n,p,c] = dy[n,p,c] + ds[m,f] / (c1*HW) * (x - mu) / (G * sd),  ds[m,f] = sum_{g,p} dy[g*M+m, p, C+f]
__global__ void __launch_bounds__(kDiscThreads) mbstd_bwd_kernel(MbstdArgs p)
{
    __shared__ float swarp[kDiscThreads / 32];
    __shared__ float s_bcast;
    const int m = blockIdx.x / p.F, f = blockIdx.x % p.F;
    const int c1 = p.c / p.F;
    const int elems = c1 * p.hw;
    float ds = 0.f;
    for (int i = threadIdx.x; i < p.G * p.hw; i += blockDim.x)
    {
        const int g = i / p.hw, px = i % p.hw;
        ds += __ldg(p.dy + ((long long)(g * p.M + m) * p.hw + px) * p.cpad + p.c + f);
    }
    for (int o = 16; o > 0; o >>= 1) ds += __shfl_down_sync(0xffffffffu, ds, o);
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        float t = 0.f;
        for (int w = 0; w < kDiscThreads / 32; w++) t += swarp[w];
        s_bcast = t / (float)elems;
    }
    __syncthreads();
    const float k = s_bcast;
    for (int i = threadIdx.x; i < elems; i += blockDim.x)
    {
        const int cc = f * c1 + i / p.hw, px = i % p.hw;
        float mu = 0.f;
        for (int g = 0; g < p.G; g++) mu += __ldg(p.x + (long long)(g * p.M + m) * p.xs_n + (long long)cc * p.xs_c + (long long)px * p.xs_p);
        mu /= (float)p.G;
        float var = 0.f;
        for (int g = 0; g < p.G; g++)
        {
            const float v = __ldg(p.x + (long long)(g * p.M + m) * p.xs_n + (long long)cc * p.xs_c + (long long)px * p.xs_p);
            var += (v - mu) * (v - mu);
        }
        const float sd = sqrtf(var / (float)p.G + 1e-8f);
        for (int g = 0; g < p.G; g++)
        {
            const float v = __ldg(p.x + (long long)(g * p.M + m) * p.xs_n + (long long)cc * p.xs_c + (long long)px * p.xs_p);
            const long long at = ((long long)(g * p.M + m) * p.hw + px);
            p.dx[at * p.c + cc] = __ldg(p.dy + at * p.cpad + cc) + k * (v - mu) / ((float)p.G * sd);
        }
    }
}

} // namespace sgv

extern "C" int sgv_fromrgb_fwd(const float* img, const float* w, const float* bias, float* y, int32_t n, int32_t hw, int32_t c, int32_t img_channels,
                               float weight_gain, int32_t act, float alpha, float gain, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(img && w && y, "sgv_fromrgb_fwd: NULL argument");
    SGV_CHECK_ARG(act == 1 || act == 3, "act must be 1 (linear) or 3 (lrelu)");
    SGV_CHECK_ARG((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0), "y and bias must be 16-byte aligned");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    FromRgbArgs a;
    memset(&a, 0, sizeof(a));
    rc = fromrgb_geom(&a, n, hw, c, img_channels);
    if (rc != SGV_OK) return rc;
    a.img = img; a.w = w; a.bias = bias; a.y = y; a.wgain = weight_gain; a.act = act; a.alpha = alpha; a.gain = gain;
    const long long total = (long long)n * hw;
    const unsigned grid = (unsigned)min((long long)num_sms() * 8, (total + a.lanes - 1) / a.lanes);
    fromrgb_fwd_kernel<<<grid, kDiscThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("fromrgb_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_fromrgb_bwd(const float* dz, const float* img, const float* w, float* dimg, float* dw, int32_t n, int32_t hw, int32_t c,
                               int32_t img_channels, float weight_gain, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dz && img && w && dw, "sgv_fromrgb_bwd: NULL argument");
    SGV_CHECK_ARG((reinterpret_cast<uintptr_t>(dz) & 15) == 0, "dz must be 16-byte aligned");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    FromRgbArgs a;
    memset(&a, 0, sizeof(a));
    rc = fromrgb_geom(&a, n, hw, c, img_channels);
    if (rc != SGV_OK) return rc;
    a.dz = dz; a.img = img; a.w = w; a.dimg = dimg; a.dw = dw; a.wgain = weight_gain;
    const long long total = (long long)n * hw;
    const unsigned grid = (unsigned)min((long long)num_sms() * 4, (total + a.lanes - 1) / a.lanes);
    fromrgb_bwd_kernel<<<grid, kDiscThreads, kDiscThreads * 16 * sizeof(float), (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("fromrgb_bwd_kernel");
    return SGV_OK;
}

static int mbstd_args(sgv::MbstdArgs* a, int n, int c, int hw, int cpad, int group, int nf, int64_t xs_n, int64_t xs_c, int64_t xs_p)
{
    using namespace sgv;
    SGV_CHECK_ARG(n >= 1 && c >= 1 && hw >= 1 && nf >= 1 && c % nf == 0, "bad extents");
    SGV_CHECK_ARG(group >= 1 && n % group == 0, "the batch (%d) must be a multiple of the group size (%d)", n, group);
    SGV_CHECK_ARG(cpad >= c + nf, "padded channel count %d < %d + %d", cpad, c, nf);
    a->n = n; a->c = c; a->hw = hw; a->cpad = cpad; a->G = group; a->M = n / group; a->F = nf;
    a->xs_n = xs_n; a->xs_c = xs_c; a->xs_p = xs_p;
    return SGV_OK;
}

extern "C" int sgv_mbstd_fwd(const float* x, int64_t x_stride_n, int64_t x_stride_c, int64_t x_stride_p, float* y, float* sd_mean,
                             int32_t n, int32_t c, int32_t hw, int32_t cpad, int32_t group, int32_t num_channels, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(x && y && sd_mean, "sgv_mbstd_fwd: NULL argument");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    MbstdArgs a;
    memset(&a, 0, sizeof(a));
    rc = mbstd_args(&a, n, c, hw, cpad, group, num_channels, x_stride_n, x_stride_c, x_stride_p);
    if (rc != SGV_OK) return rc;
    a.x = x; a.y = y; a.sd_mean = sd_mean;
    mbstd_fwd_kernel<<<(unsigned)(a.M * a.F), kDiscThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("mbstd_fwd_kernel");
    return SGV_OK;
}

extern "C" int sgv_mbstd_bwd(const float* dy, const float* x, int64_t x_stride_n, int64_t x_stride_c, int64_t x_stride_p, float* dx,
                             int32_t n, int32_t c, int32_t hw, int32_t cpad, int32_t group, int32_t num_channels, void* stream_)
{
    using namespace sgv;
    SGV_CHECK_ARG(dy && x && dx, "sgv_mbstd_bwd: NULL argument");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    MbstdArgs a;
    memset(&a, 0, sizeof(a));
    rc = mbstd_args(&a, n, c, hw, cpad, group, num_channels, x_stride_n, x_stride_c, x_stride_p);
    if (rc != SGV_OK) return rc;
    a.x = x; a.dy = dy; a.dx = dx;
    mbstd_bwd_kernel<<<(unsigned)(a.M * a.F), kDiscThreads, 0, (cudaStream_t)stream_>>>(a);
    SGV_LAUNCH_OK("mbstd_bwd_kernel");
    return SGV_OK;
}
