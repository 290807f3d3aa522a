// upfirdn2d for sm_100a: pad / zero-insert upsample / 2-D FIR / decimate, HBM-bound.
//
// Replaces the reference plugin src/torch_utils/ops/upfirdn2d.{cpp,cu}.  Integer geometry is
// the reference's (upfirdn2d.cu:43-48,60-71,176-185; output size upfirdn2d.cpp:32-33) so tap selection is
// bit-exact; taps are accumulated in the reference's order (input row ascending, then input column ascending)
// with fused multiply-adds, then multiplied by `gain`, so fp32 results are bit-identical to the reference
// kernels as well.  Three kernels:
//
//   fir_nchw_tiled   W-contiguous fp32 tensors.  A CTA stages an input tile in shared memory with 128-bit
//                    global loads that keep the global 16-byte phase of every row (rows of odd width such as
//                    the 257-wide transposed-conv outputs stay vector-loadable), FIR taps live in shared
//                    memory / registers, each thread produces 4 consecutive outputs per row from a register
//                    window and writes them with one 128-bit store.
//   fir_nhwc         channels_last fp32 tensors: a thread owns 4 channels (one 128-bit lane) of 1-2 output
//                    pixels; taps stream through L1.
//   fir_generic      any dtype / stride / factor: one gather per output (same math as upfirdn2d_kernel_large).
//
// All three optionally apply the fused epilogue of sgv_upfirdn2d_params (scale[n,c], bias[c], lrelu, gain, clamp).
#include <stdlib.h>
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace sgv {

struct FirArgs
{
    const void* x; const float* f; void* y;
    int upx, upy, downx, downy, pad_x0, pad_y0, flip; float gain;
    int in_w, in_h, in_c, in_n; long long isx, isy, isc, isn;
    int f_w, f_h; long long fsx, fsy;
    int out_w, out_h; long long osx, osy, osc, osn;
    const float* escale; const float* ebias; int eact; float ealpha, egain, eclamp; int eround;
    const float* enoise; long long ensn, ensy, ensx;      // per-pixel noise plane(s) added after the scale (fma(x, dcoefs, noise), networks.py:68-69)
};

__device__ __forceinline__ float fir_noise(const FirArgs& p, int n, int oy, int ox)
{
    // callers may evaluate the epilogue for the (masked) columns past a ragged edge: keep the read in range
    return __ldg(p.enoise + (long long)n * p.ensn + (long long)min(oy, p.out_h - 1) * p.ensy + (long long)min(ox, p.out_w - 1) * p.ensx);
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

template <class A, bool EPI = true>
__device__ __forceinline__ A fir_epilogue(A v, const FirArgs& p, int n, int c, int oy, int ox)
{
    if (!EPI || p.eact == 0) return v;
    // separate roundings (no FMA contraction) so the fused result equals the unfused op sequence bit for bit
    if (p.escale) v = mul_rn(v, (A)p.escale[(long long)n * p.in_c + c]);
    if (p.enoise) v = add_rn(v, (A)fir_noise(p, n, oy, ox));
    if (p.ebias) v = add_rn(v, (A)p.ebias[c]);
    if (p.eact == 3) v = (v > 0) ? v : v * (A)p.ealpha;
    v *= (A)p.egain;
    if (p.eclamp >= 0) { A cl = (A)p.eclamp; v = (v > -cl && v < cl) ? v : (v >= 0 ? cl : -cl); }
    return v;
}

__device__ __forceinline__ float fmadd(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fmadd(double a, double b, double c) { return fma(a, b, c); }

//------------------------------------------------------------------------------------------------
// Generic gather kernel.

template <class T>
__global__ void __launch_bounds__(256) fir_generic(FirArgs p, int channels_fast, long long total)
{
    typedef typename acc_type<T>::type A;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        int n, c, outY, outX;
        long long r = idx;
        if (channels_fast) { c = (int)(r % p.in_c); r /= p.in_c; outX = (int)(r % p.out_w); r /= p.out_w; outY = (int)(r % p.out_h); n = (int)(r / p.out_h); }
        else               { outX = (int)(r % p.out_w); r /= p.out_w; outY = (int)(r % p.out_h); r /= p.out_h; c = (int)(r % p.in_c); n = (int)(r / p.in_c); }

        int midY = outY * p.downy + p.upy - 1 - p.pad_y0;
        int inY = min(max(floor_div(midY, p.upy), 0), p.in_h);
        int h = min(max(floor_div(midY + p.f_h, p.upy), 0), p.in_h) - inY;
        int filterY = midY + p.f_h - (inY + 1) * p.upy;
        if (p.flip) filterY = p.f_h - 1 - filterY;
        int midX = outX * p.downx + p.upx - 1 - p.pad_x0;
        int inX = min(max(floor_div(midX, p.upx), 0), p.in_w);
        int w = min(max(floor_div(midX + p.f_w, p.upx), 0), p.in_w) - inX;
        int filterX = midX + p.f_w - (inX + 1) * p.upx;
        if (p.flip) filterX = p.f_w - 1 - filterX;

        const T* xp = (const T*)p.x + inX * p.isx + inY * p.isy + c * p.isc + n * p.isn;
        const float* fp = p.f + filterX * p.fsx + filterY * p.fsy;
        long long stepX = (p.flip ? p.upx : -p.upx) * p.fsx;
        long long stepY = (p.flip ? p.upy : -p.upy) * p.fsy;
        A v = 0;
        for (int y = 0; y < h; y++)
        {
            for (int x = 0; x < w; x++)
            {
                v = fmadd((A)(*xp), (A)(*fp), v);
                xp += p.isx;
                fp += stepX;
            }
            xp += p.isy - w * p.isx;
            fp += stepY - w * stepX;
        }
        v *= (A)p.gain;
        v = fir_epilogue<A>(v, p, n, c, outY, outX);
        ((T*)p.y)[outX * p.osx + outY * p.osy + c * p.osc + n * p.osn] = (T)v;
    }
}

//------------------------------------------------------------------------------------------------
// Tiled kernel for dense NCHW fp32.
//   block = 256 threads = lanes_x (power of two, <= 32) x (256 / lanes_x) thread rows;
//   each thread: 4 consecutive outputs in x, RPT = 4 output rows.

constexpr int kFirThreads = 256;
constexpr int kFirRPT = 4;
constexpr int kFirMaxTaps = 32;   // per dimension, for the runtime-sized filter variants

struct FirTile
{
    int lanes_x_log2;     // threads along x = 1 << lanes_x_log2
    int tile_out_w, tile_out_h;
    int tile_in_w, tile_in_h;
    int vecs_per_row;     // float4 per staged input row
    int tiles_x, tiles_y;
    long long in_total;   // numel of x (for vector-load bounds)
    int out_vec_ok;       // 128-bit stores allowed
    FastDiv div_vpr;      // division by vecs_per_row
};

template <int UPX, int UPY, int DOWNX, int DOWNY, int FW_T, int FH_T, bool EPI>
__global__ void __launch_bounds__(kFirThreads) fir_nchw_tiled(FirArgs p, FirTile t)
{
    extern __shared__ __align__(16) float smem[];
    const int fw = FW_T > 0 ? FW_T : p.f_w;
    const int fh = FH_T > 0 ? FH_T : p.f_h;
    // filter extents padded up to a multiple of the upsampling factor (taps beyond the real filter are zero)
    const int fwp = ceil_div(fw, UPX) * UPX;
    const int fhp = ceil_div(fh, UPY) * UPY;
    float* sf = smem;                               // [fhp][fwp]   (flipped like upfirdn2d.cu:118-130)
    float* sx = smem + ((fhp * fwp + 3) & ~3);      // [tile_in_h][pitch]
    const int pitch = t.vecs_per_row * 4;

    const int tid = threadIdx.x;
    long long bid = blockIdx.x;
    const int tile_x = (int)(bid % t.tiles_x); bid /= t.tiles_x;
    const int tile_y = (int)(bid % t.tiles_y); bid /= t.tiles_y;
    const long long plane = bid;                     // n * C + c
    const int n = (int)(plane / p.in_c), c = (int)(plane % p.in_c);

    for (int i = tid; i < fhp * fwp; i += kFirThreads)
    {
        int fy = i / fwp, fx = i - fy * fwp;
        float v = 0.f;
        if (fx < p.f_w && fy < p.f_h)
        {
            int ffx = p.flip ? fx : p.f_w - 1 - fx;
            int ffy = p.flip ? fy : p.f_h - 1 - fy;
            v = p.f[ffx * p.fsx + ffy * p.fsy];
        }
        sf[i] = v;
    }

    const int tileOutX = tile_x * t.tile_out_w;
    const int tileOutY = tile_y * t.tile_out_h;
    const int tileMidX = tileOutX * DOWNX + UPX - 1 - p.pad_x0;
    const int tileMidY = tileOutY * DOWNY + UPY - 1 - p.pad_y0;
    const int tileInX = floor_div(tileMidX, UPX);
    const int tileInY = floor_div(tileMidY, UPY);

    // ---- stage the input tile: aligned 128-bit loads, each row keeps its global 16-byte phase ----
    const float* xg = (const float*)p.x;
    const long long plane_off = plane * (long long)p.in_h * p.in_w;
    constexpr int LDU = 4;     // independent 128-bit loads in flight per thread
    const int nvec_tile = t.tile_in_h * t.vecs_per_row;
    for (int base = 0; base < nvec_tile; base += kFirThreads * LDU)
    {
        float4 val[LDU];
        int xs[LDU], so[LDU];
#pragma unroll
        for (int u = 0; u < LDU; u++)
        {
            const int i = base + u * kFirThreads + tid;
            val[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            so[u] = -1;
            if (i < nvec_tile)
            {
                const int r = (int)t.div_vpr.div((uint32_t)i), v = i - r * t.vecs_per_row;
                const int inY = tileInY + r;
                const long long rowbase = plane_off + (long long)inY * p.in_w + tileInX;
                const int phase = (int)(rowbase & 3);
                const long long e0 = rowbase - phase + 4 * v;
                const int x0 = tileInX - phase + 4 * v;
                xs[u] = x0;
                so[u] = r * pitch + 4 * v;
                if (inY >= 0 && inY < p.in_h && x0 + 3 >= 0 && x0 < p.in_w)
                {
                    if (e0 >= 0 && e0 + 3 < t.in_total)
                        val[u] = __ldg(reinterpret_cast<const float4*>(xg + e0));
                    else
                    {
                        if (e0 + 0 >= 0 && e0 + 0 < t.in_total) val[u].x = __ldg(xg + e0 + 0);
                        if (e0 + 1 >= 0 && e0 + 1 < t.in_total) val[u].y = __ldg(xg + e0 + 1);
                        if (e0 + 2 >= 0 && e0 + 2 < t.in_total) val[u].z = __ldg(xg + e0 + 2);
                        if (e0 + 3 >= 0 && e0 + 3 < t.in_total) val[u].w = __ldg(xg + e0 + 3);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < LDU; u++)
        {
            if (so[u] < 0) continue;
            const int x0 = xs[u];
            if (x0 + 0 < 0 || x0 + 0 >= p.in_w) val[u].x = 0.f;
            if (x0 + 1 < 0 || x0 + 1 >= p.in_w) val[u].y = 0.f;
            if (x0 + 2 < 0 || x0 + 2 >= p.in_w) val[u].z = 0.f;
            if (x0 + 3 < 0 || x0 + 3 >= p.in_w) val[u].w = 0.f;
            *reinterpret_cast<float4*>(sx + so[u]) = val[u];
        }
    }
    __syncthreads();

    // ---- compute ----
    const int lanes_x = 1 << t.lanes_x_log2;
    const int tx = tid & (lanes_x - 1), ty = tid >> t.lanes_x_log2;
    const int relOutX0 = tx * 4;
    const int relOutY0 = ty * kFirRPT;
    if (relOutX0 >= t.tile_out_w || relOutY0 >= t.tile_out_h) return;
    const int outX0 = tileOutX + relOutX0;
    const int outY0 = tileOutY + relOutY0;
    if (outX0 >= p.out_w || outY0 >= p.out_h) return;

    float acc[kFirRPT][4];
#pragma unroll
    for (int r = 0; r < kFirRPT; r++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[r][k] = 0.f;

    if constexpr (UPX == 1 && UPY == 1 && FW_T > 0 && FH_T > 0 && FW_T * FH_T <= 16)
    {
        // register window: rows stream through, every staged value is read from shared memory once per thread
        float fr[FH_T][FW_T];
#pragma unroll
        for (int y = 0; y < FH_T; y++)
#pragma unroll
            for (int x = 0; x < FW_T; x++) fr[y][x] = sf[y * FW_T + x];
        constexpr int WIN_W = 3 * DOWNX + FW_T;
        constexpr int WIN_H = (kFirRPT - 1) * DOWNY + FH_T;
        const int relInX = relOutX0 * DOWNX;         // up == 1: in = mid, tileIn = tileMid
        const int relInY = relOutY0 * DOWNY;
#pragma unroll
        for (int wr = 0; wr < WIN_H; wr++)
        {
            const int row = relInY + wr;
            const long long rowbase = plane_off + (long long)(tileInY + row) * p.in_w + tileInX;
            const float* srow = sx + row * pitch + (int)(rowbase & 3) + relInX;
            float win[WIN_W];
#pragma unroll
            for (int i = 0; i < WIN_W; i++) win[i] = srow[i];
#pragma unroll
            for (int r = 0; r < kFirRPT; r++)
            {
                const int fy = wr - r * DOWNY;
                if (fy >= 0 && fy < FH_T)
                {
#pragma unroll
                    for (int k = 0; k < 4; k++)
#pragma unroll
                        for (int x = 0; x < FW_T; x++)
                            acc[r][k] = fmaf(win[k * DOWNX + x], fr[fy][x], acc[r][k]);
                }
            }
        }
    }
    else
    {
        const int ntx = fwp / UPX, nty = fhp / UPY;
#pragma unroll
        for (int r = 0; r < kFirRPT; r++)
        {
            const int midY = tileMidY + (relOutY0 + r) * DOWNY;
            const int inY = floor_div(midY, UPY);
            const int relInY = inY - tileInY;
            const int filterY = (inY + 1) * UPY - midY - 1;
            if (relOutY0 + r >= t.tile_out_h) continue;
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                if (relOutX0 + k >= t.tile_out_w) continue;
                const int midX = tileMidX + (relOutX0 + k) * DOWNX;
                const int inX = floor_div(midX, UPX);
                const int relInX = inX - tileInX;
                const int filterX = (inX + 1) * UPX - midX - 1;
                float v = 0.f;
                for (int y = 0; y < nty; y++)
                {
                    const int row = relInY + y;
                    const long long rowbase = plane_off + (long long)(tileInY + row) * p.in_w + tileInX;
                    const float* srow = sx + row * pitch + (int)(rowbase & 3) + relInX;
                    const float* frow = sf + (filterY + y * UPY) * fwp + filterX;
                    for (int x = 0; x < ntx; x++)
                        v = fmaf(srow[x], frow[x * UPX], v);
                }
                acc[r][k] = v;
            }
        }
    }

    float* yg = (float*)p.y;
#pragma unroll
    for (int r = 0; r < kFirRPT; r++)
    {
        const int outY = outY0 + r;
        if (relOutY0 + r >= t.tile_out_h || outY >= p.out_h) continue;
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) o[k] = fir_epilogue<float, EPI>(acc[r][k] * p.gain, p, n, c, outY, outX0 + k);
        float* dst = yg + (plane * p.out_h + outY) * (long long)p.out_w + outX0;
        if (t.out_vec_ok && outX0 + 3 < p.out_w && relOutX0 + 3 < t.tile_out_w)
            __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
        else
        {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (outX0 + k < p.out_w && relOutX0 + k < t.tile_out_w) dst[k] = o[k];
        }
    }
}

//------------------------------------------------------------------------------------------------
// channels_last fp32 kernel: thread = (4 channels) x (1 output column) x (ROWS output rows).

template <int VEC> struct vec_t;
template <> struct vec_t<4> { typedef float4 type; };
template <> struct vec_t<1> { typedef float type; };

__device__ __forceinline__ void vfma(float4& a, const float4& x, float f) { a.x = fmaf(x.x, f, a.x); a.y = fmaf(x.y, f, a.y); a.z = fmaf(x.z, f, a.z); a.w = fmaf(x.w, f, a.w); }
__device__ __forceinline__ void vfma(float& a, const float& x, float f) { a = fmaf(x, f, a); }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float& a) { a = 0.f; }

template <int VEC, bool EPI>
__device__ __forceinline__ void nhwc_store(const FirArgs& p, typename vec_t<VEC>::type v, int n, int c0, int outY, int outX)
{
    float* dst = (float*)p.y + n * p.osn + outY * p.osy + outX * p.osx + c0;
    if constexpr (VEC == 4)
    {
        float4 o;
        o.x = fir_epilogue<float, EPI>(v.x * p.gain, p, n, c0 + 0, outY, outX);
        o.y = fir_epilogue<float, EPI>(v.y * p.gain, p, n, c0 + 1, outY, outX);
        o.z = fir_epilogue<float, EPI>(v.z * p.gain, p, n, c0 + 2, outY, outX);
        o.w = fir_epilogue<float, EPI>(v.w * p.gain, p, n, c0 + 3, outY, outX);
        __stcs(reinterpret_cast<float4*>(dst), o);
    }
    else
        *dst = fir_epilogue<float, EPI>(v * p.gain, p, n, c0, outY, outX);
}

// FAST: up = 1 in both dims, compile-time FWxFH filter, ROWS = 2 output rows per thread.
template <int VEC, int DOWN, int FW_T, int FH_T, bool EPI>
__global__ void __launch_bounds__(256) fir_nhwc_fast(FirArgs p, long long total)
{
    typedef typename vec_t<VEC>::type V;
    __shared__ float sf[FH_T * FW_T];
    if (threadIdx.x < FH_T * FW_T)
    {
        int fy = threadIdx.x / FW_T, fx = threadIdx.x - fy * FW_T;
        int ffx = p.flip ? fx : FW_T - 1 - fx;
        int ffy = p.flip ? fy : FH_T - 1 - fy;
        sf[threadIdx.x] = p.f[ffx * p.fsx + ffy * p.fsy];
    }
    __syncthreads();
    const int cvecs = p.in_c / VEC;
    const int rows2 = (p.out_h + 1) / 2;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        long long r = idx;
        const int cv = (int)(r % cvecs); r /= cvecs;
        const int outX = (int)(r % p.out_w); r /= p.out_w;
        const int oy2 = (int)(r % rows2);
        const int n = (int)(r / rows2);
        const int outY0 = oy2 * 2;
        const int c0 = cv * VEC;
        const int inX0 = outX * DOWN - p.pad_x0;          // up == 1: mid = in
        const int inY0 = outY0 * DOWN - p.pad_y0;
        V acc0, acc1;
        vzero(acc0); vzero(acc1);
        const float* xb = (const float*)p.x + n * p.isn + c0;
        constexpr int WIN_H = DOWN + FH_T;
#pragma unroll
        for (int wr = 0; wr < WIN_H; wr++)
        {
            const int inY = inY0 + wr;
            if (inY < 0 || inY >= p.in_h) continue;
            V win[FW_T];
#pragma unroll
            for (int x = 0; x < FW_T; x++)
            {
                const int inX = inX0 + x;
                vzero(win[x]);
                if (inX >= 0 && inX < p.in_w)
                    win[x] = __ldg(reinterpret_cast<const V*>(xb + inY * p.isy + inX * p.isx));
            }
            if (wr < FH_T)
            {
#pragma unroll
                for (int x = 0; x < FW_T; x++) vfma(acc0, win[x], sf[wr * FW_T + x]);
            }
            if (wr >= DOWN)
            {
#pragma unroll
                for (int x = 0; x < FW_T; x++) vfma(acc1, win[x], sf[(wr - DOWN) * FW_T + x]);
            }
        }
        nhwc_store<VEC, EPI>(p, acc0, n, c0, outY0, outX);
        if (outY0 + 1 < p.out_h) nhwc_store<VEC, EPI>(p, acc1, n, c0, outY0 + 1, outX);
    }
}

// Sliding-window channels_last kernel for the hot geometry (up = down = 1, 4x4 taps): a thread owns 4 channels of TWO output
// rows and walks SEG output columns; the 5x4 input window lives in registers and only ONE new column (5 x 128-bit loads) is
// fetched per step, i.e. 2.5 loads per output vector instead of 10.  Tap order per output is unchanged (rows, then columns, ascending) => same bits.
template <bool EPI>
__global__ void __launch_bounds__(256, 2) fir_nhwc_slide44(FirArgs p, long long total, int seg, int nseg)
{
    __shared__ float sf[16];
    if (threadIdx.x < 16)
    {
        int fy = threadIdx.x >> 2, fx = threadIdx.x & 3;
        int ffx = p.flip ? fx : 3 - fx;
        int ffy = p.flip ? fy : 3 - fy;
        sf[threadIdx.x] = p.f[ffx * p.fsx + ffy * p.fsy];
    }
    __syncthreads();
    float fr[4][4];
#pragma unroll
    for (int i = 0; i < 16; i++) fr[i >> 2][i & 3] = sf[i];
    const int cvecs = p.in_c >> 2;
    const int rows2 = (p.out_h + 1) >> 1;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        long long r = idx;
        const int cv = (int)(r % cvecs); r /= cvecs;
        const int sg = (int)(r % nseg); r /= nseg;
        const int oy2 = (int)(r % rows2);
        const int n = (int)(r / rows2);
        const int c0 = cv * 4;
        const int outY0 = oy2 * 2;
        const int x_begin = sg * seg;
        const int x_end = min(x_begin + seg, p.out_w);
        const int inY0 = outY0 - p.pad_y0;
        const float* xb = (const float*)p.x + n * p.isn + c0;
        const float* rowp[5]; bool rowok[5];
#pragma unroll
        for (int k = 0; k < 5; k++) { const int iy = inY0 + k; rowok[k] = iy >= 0 && iy < p.in_h; rowp[k] = xb + (long long)(rowok[k] ? iy : 0) * p.isy; }
        auto load_col = [&](int inX, float4 (&col)[5]) {
            const bool cok = inX >= 0 && inX < p.in_w;
#pragma unroll
            for (int k = 0; k < 5; k++)
                col[k] = (cok && rowok[k]) ? __ldg(reinterpret_cast<const float4*>(rowp[k] + (long long)inX * p.isx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // fused epilogue constants of this item (n and the 4 channels are fixed): loaded once instead of per output
        float4 esc = make_float4(1.f, 1.f, 1.f, 1.f), ebi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI && p.eact != 0)
        {
            if (p.escale) esc = __ldg(reinterpret_cast<const float4*>(p.escale + (long long)n * p.in_c + c0));
            if (p.ebias) ebi = __ldg(reinterpret_cast<const float4*>(p.ebias + c0));
        }
        auto store = [&](float4 a, int outY, int outX) {
            float4 o = make_float4(a.x * p.gain, a.y * p.gain, a.z * p.gain, a.w * p.gain);
            if (EPI && p.eact != 0)
            {
                // separate roundings (no FMA contraction): bit-identical to the unfused op sequence, like fir_epilogue
                if (p.escale) { o.x = __fmul_rn(o.x, esc.x); o.y = __fmul_rn(o.y, esc.y); o.z = __fmul_rn(o.z, esc.z); o.w = __fmul_rn(o.w, esc.w); }
                if (p.enoise) { const float nz = fir_noise(p, n, outY, outX); o.x = __fadd_rn(o.x, nz); o.y = __fadd_rn(o.y, nz); o.z = __fadd_rn(o.z, nz); o.w = __fadd_rn(o.w, nz); }
                if (p.ebias) { o.x = __fadd_rn(o.x, ebi.x); o.y = __fadd_rn(o.y, ebi.y); o.z = __fadd_rn(o.z, ebi.z); o.w = __fadd_rn(o.w, ebi.w); }
                if (p.eact == 3) { o.x = o.x > 0.f ? o.x : o.x * p.ealpha; o.y = o.y > 0.f ? o.y : o.y * p.ealpha; o.z = o.z > 0.f ? o.z : o.z * p.ealpha; o.w = o.w > 0.f ? o.w : o.w * p.ealpha; }
                o.x *= p.egain; o.y *= p.egain; o.z *= p.egain; o.w *= p.egain;
                if (p.eclamp >= 0.f)
                {
                    const float cl = p.eclamp;
                    o.x = (o.x > -cl && o.x < cl) ? o.x : (o.x >= 0.f ? cl : -cl); o.y = (o.y > -cl && o.y < cl) ? o.y : (o.y >= 0.f ? cl : -cl);
                    o.z = (o.z > -cl && o.z < cl) ? o.z : (o.z >= 0.f ? cl : -cl); o.w = (o.w > -cl && o.w < cl) ? o.w : (o.w >= 0.f ? cl : -cl);
                }
                if (p.eround) { o.x = ptx::tf32_rn(o.x); o.y = ptx::tf32_rn(o.y); o.z = ptx::tf32_rn(o.z); o.w = ptx::tf32_rn(o.w); }
            }
            __stcs(reinterpret_cast<float4*>((float*)p.y + n * p.osn + outY * p.osy + outX * p.osx + c0), o);
        };
        float4 win[4][5];                          // win[slot][row]: columns inX .. inX+3 of the current output, slots rotate
        const int inX0 = x_begin - p.pad_x0;
        load_col(inX0 + 0, win[0]); load_col(inX0 + 1, win[1]); load_col(inX0 + 2, win[2]);
        for (int xo = x_begin; xo < x_end; xo += 4)
        {
#pragma unroll
            for (int u = 0; u < 4; u++)
            {
                const int outX = xo + u;
                load_col(outX - p.pad_x0 + 3, win[(u + 3) & 3]);      // slot (u+3)%4 receives the newest column
                if (outX < x_end)
                {
                    float4 a0, a1; vzero(a0); vzero(a1);
#pragma unroll
                    for (int ry = 0; ry < 4; ry++)
#pragma unroll
                        for (int cx = 0; cx < 4; cx++)
                        {
                            vfma(a0, win[(u + cx) & 3][ry], fr[ry][cx]);
                        }
#pragma unroll
                    for (int ry = 0; ry < 4; ry++)
#pragma unroll
                        for (int cx = 0; cx < 4; cx++)
                        {
                            vfma(a1, win[(u + cx) & 3][ry + 1], fr[ry][cx]);
                        }
                    store(a0, outY0, outX);
                    if (outY0 + 1 < p.out_h) store(a1, outY0 + 1, outX);
                }
            }
        }
    }
}

// channels_last, up = down = 1, 4x4 filter, C % 32 == 0: TMA-fed persistent kernel.
// A CTA walks tiles of 8 x 32 output pixels x 32 channels.  The (8+3) x (32+3) pixel x 128 B input box of tile i+1 (and i+2) is in
// flight (cp.async.bulk.tensor, out-of-range pixels = the zero padding) while tile i is filtered out of shared memory, so ~100 KB
// of reads per SM are always outstanding — the LDG sliding-window kernel above issues its loads right before it needs them and
// reaches about half of the HBM rate (profiles/launches_r1ac summary: 2.1-3.3 TB/s).  Thread = 4 channels x one output column,
// walking down the 8 rows with a 4 x 4 register window (4 LDS.128 per output); tap order = fir_nhwc_slide44's, results identical.
constexpr int kFtTW = 32, kFtTH = 8;
constexpr int kFtBoxW = kFtTW + 3, kFtBoxH = kFtTH + 3;
constexpr int kFtStage = ((kFtBoxW * kFtBoxH * 128) + 1023) & ~1023;
constexpr int kFtSmem = 2 * kFtStage + 2 * 8 + 1024;

template <bool EPI>
__global__ void __launch_bounds__(256, 2) fir_nhwc_tma44(const __grid_constant__ CUtensorMap tmap_x, FirArgs p, int tiles_x, int tiles_y, int cblocks, int total_tiles)
{
    using namespace ptx;
    extern __shared__ uint8_t fsmem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fsmem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kFtStage);
    __shared__ float sf[16];
    if (threadIdx.x < 16)
    {
        int fy = threadIdx.x >> 2, fx = threadIdx.x & 3;
        int ffx = p.flip ? fx : 3 - fx;
        int ffy = p.flip ? fy : 3 - fy;
        sf[threadIdx.x] = p.f[ffx * p.fsx + ffy * p.fsy];
    }
    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_x);
        mbar_init(full + 0, 1); mbar_init(full + 1, 1);
        fence_mbar_init();
    }
    __syncthreads();
    float fr[4][4];
#pragma unroll
    for (int i = 0; i < 16; i++) fr[i >> 2][i & 3] = sf[i];

    auto decode = [&](int t, int& n, int& oy0, int& ox0, int& cb) {
        cb = t % cblocks; t /= cblocks;
        ox0 = (t % tiles_x) * kFtTW; t /= tiles_x;
        oy0 = (t % tiles_y) * kFtTH;
        n = t / tiles_y;
    };
    auto issue = [&](int t, int s) {
        int n, oy0, ox0, cb;
        decode(t, n, oy0, ox0, cb);
        mbar_expect_tx(full + s, (uint32_t)(kFtBoxW * kFtBoxH * 128));
        tma_load_4d(smem + s * kFtStage, &tmap_x, full + s, cb * 32, ox0 - p.pad_x0, oy0 - p.pad_y0, n);
    };
    if (threadIdx.x == 0)
    {
        if ((int)blockIdx.x < total_tiles) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < total_tiles) issue(blockIdx.x + gridDim.x, 1);
    }
    const int cv = threadIdx.x & 7, col = threadIdx.x >> 3;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++)
    {
        const int s = it & 1;
        int n, oy0, ox0, cb;
        decode(t, n, oy0, ox0, cb);
        const int c0 = cb * 32 + cv * 4;
        float4 esc = make_float4(1.f, 1.f, 1.f, 1.f), ebi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI && p.eact != 0)
        {
            if (p.escale) esc = __ldg(reinterpret_cast<const float4*>(p.escale + (long long)n * p.in_c + c0));
            if (p.ebias) ebi = __ldg(reinterpret_cast<const float4*>(p.ebias + c0));
        }
        mbar_wait(full + s, (uint32_t)(it >> 1) & 1u);
        const uint32_t base = smem_u32(smem + s * kFtStage);
        auto load_row = [&](int r, float4 (&w)[4]) {
#pragma unroll
            for (int cx = 0; cx < 4; cx++)
            {
                const uint32_t line = (uint32_t)(r * kFtBoxW + col + cx);
                w[cx] = lds128(base + line * 128u + (uint32_t)((cv ^ (line & 7u)) << 4));
            }
        };
        float4 win[4][4];                              // win[slot][cx]: input row (slot rotates), columns col .. col+3
        load_row(0, win[0]); load_row(1, win[1]); load_row(2, win[2]);
        const int ox = ox0 + col;
        float* ycol = (float*)p.y + n * p.osn + ox * p.osx + c0;
#pragma unroll
        for (int r = 0; r < kFtTH; r++)
        {
            load_row(r + 3, win[(r + 3) & 3]);
            float4 a; vzero(a);
#pragma unroll
            for (int ry = 0; ry < 4; ry++)
#pragma unroll
                for (int cx = 0; cx < 4; cx++) vfma(a, win[(r + ry) & 3][cx], fr[ry][cx]);
            const int oy = oy0 + r;
            if (oy < p.out_h && ox < p.out_w)
            {
                float4 o = make_float4(a.x * p.gain, a.y * p.gain, a.z * p.gain, a.w * p.gain);
                if (EPI && p.eact != 0)
                {
                    // separate roundings (no FMA contraction): bit-identical to the unfused op sequence, like fir_epilogue
                    if (p.escale) { o.x = __fmul_rn(o.x, esc.x); o.y = __fmul_rn(o.y, esc.y); o.z = __fmul_rn(o.z, esc.z); o.w = __fmul_rn(o.w, esc.w); }
                    if (p.enoise) { const float nz = fir_noise(p, n, oy, ox); o.x = __fadd_rn(o.x, nz); o.y = __fadd_rn(o.y, nz); o.z = __fadd_rn(o.z, nz); o.w = __fadd_rn(o.w, nz); }
                    if (p.ebias) { o.x = __fadd_rn(o.x, ebi.x); o.y = __fadd_rn(o.y, ebi.y); o.z = __fadd_rn(o.z, ebi.z); o.w = __fadd_rn(o.w, ebi.w); }
                    if (p.eact == 3) { o.x = o.x > 0.f ? o.x : o.x * p.ealpha; o.y = o.y > 0.f ? o.y : o.y * p.ealpha; o.z = o.z > 0.f ? o.z : o.z * p.ealpha; o.w = o.w > 0.f ? o.w : o.w * p.ealpha; }
                    o.x *= p.egain; o.y *= p.egain; o.z *= p.egain; o.w *= p.egain;
                    if (p.eclamp >= 0.f)
                    {
                        const float cl = p.eclamp;
                        o.x = (o.x > -cl && o.x < cl) ? o.x : (o.x >= 0.f ? cl : -cl); o.y = (o.y > -cl && o.y < cl) ? o.y : (o.y >= 0.f ? cl : -cl);
                        o.z = (o.z > -cl && o.z < cl) ? o.z : (o.z >= 0.f ? cl : -cl); o.w = (o.w > -cl && o.w < cl) ? o.w : (o.w >= 0.f ? cl : -cl);
                    }
                    if (p.eround) { o.x = ptx::tf32_rn(o.x); o.y = ptx::tf32_rn(o.y); o.z = ptx::tf32_rn(o.z); o.w = ptx::tf32_rn(o.w); }
                }
                __stcs(reinterpret_cast<float4*>(ycol + oy * p.osy), o);
            }
        }
        __syncthreads();                               // every thread is done reading stage s
        if (threadIdx.x == 0)
        {
            const long long nxt = (long long)t + 2LL * gridDim.x;
            if (nxt < total_tiles) issue((int)nxt, s);
        }
    }
}

// Zero-insertion x2 with 4x4 taps on channels_last tensors, even leading pads: the ADJOINT of the discriminator's decimating FIR
// (conv2d_resample.py:100-110 forward with down = 2 -> upfirdn2d.py:205-213 backward: up = 2, pad0 = 2) — 5.1 ms of a G + D step on the
// general kernel below (one output per thread, runtime tap loops: 21 % of the HBM rate, profiles/timeline_gd_step_r2h.txt).
// A thread owns 4 channels of a 2x2 output quad: with an even pad the quad (2qy + a, 2qx + b) reads the 3x3 input window starting at
// (qy - pad_y0/2, qx - pad_x0/2); output parity a uses window rows a, a + 1 with taps 3 - a, 1 - a (mirrored when flip).  9 vector loads
// for 4 vector stores; per output the taps are accumulated rows-then-columns ascending with the same FMAs as fir_nhwc_any and out-of-range
// taps are skipped, not added as zeros => identical bits.
__global__ void __launch_bounds__(256) fir_nhwc_up2_44(FirArgs p, long long total, int qw, int qh)
{
    __shared__ float sf[16];
    if (threadIdx.x < 16)
    {
        const int fy = threadIdx.x >> 2, fx = threadIdx.x & 3;
        sf[threadIdx.x] = p.f[(p.flip ? 3 - fx : fx) * p.fsx + (p.flip ? 3 - fy : fy) * p.fsy];
    }
    __syncthreads();
    const int cvecs = p.in_c / 4;
    const int sx = -(p.pad_x0 / 2), sy = -(p.pad_y0 / 2);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        long long r = idx;
        const int cv = (int)(r % cvecs); r /= cvecs;
        const int qx = (int)(r % qw); r /= qw;
        const int qy = (int)(r % qh);
        const int n = (int)(r / qh);
        const int c0 = cv * 4;
        const int ix0 = qx + sx, iy0 = qy + sy;
        const float* xb = (const float*)p.x + n * p.isn + c0;
        float4 win[3][3];
        bool rok[3], cok[3];
#pragma unroll
        for (int w = 0; w < 3; w++) { rok[w] = (iy0 + w >= 0) && (iy0 + w < p.in_h); cok[w] = (ix0 + w >= 0) && (ix0 + w < p.in_w); }
#pragma unroll
        for (int wy = 0; wy < 3; wy++)
#pragma unroll
            for (int wx = 0; wx < 3; wx++)
            {
                vzero(win[wy][wx]);
                if (rok[wy] && cok[wx]) win[wy][wx] = __ldg(reinterpret_cast<const float4*>(xb + (iy0 + wy) * p.isy + (ix0 + wx) * p.isx));
            }
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++)
            {
                const int oy = 2 * qy + a, ox = 2 * qx + b;
                if (oy >= p.out_h || ox >= p.out_w) continue;
                float4 acc; vzero(acc);
#pragma unroll
                for (int ty = 0; ty < 2; ty++)
                {
                    if (!rok[a + ty]) continue;
#pragma unroll
                    for (int tx = 0; tx < 2; tx++)
                        if (cok[b + tx]) vfma(acc, win[a + ty][b + tx], sf[(3 - a - 2 * ty) * 4 + (3 - b - 2 * tx)]);
                }
                nhwc_store<4, true>(p, acc, n, c0, oy, ox);
            }
    }
}

// General channels_last kernel (any up/down/filter), one output pixel x VEC channels per thread.
template <int VEC>
__global__ void __launch_bounds__(256) fir_nhwc_any(FirArgs p, long long total)
{
    typedef typename vec_t<VEC>::type V;
    const int cvecs = p.in_c / VEC;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    {
        long long r = idx;
        const int cv = (int)(r % cvecs); r /= cvecs;
        const int outX = (int)(r % p.out_w); r /= p.out_w;
        const int outY = (int)(r % p.out_h);
        const int n = (int)(r / p.out_h);
        const int c0 = cv * VEC;
        int midY = outY * p.downy + p.upy - 1 - p.pad_y0;
        int inY = min(max(floor_div(midY, p.upy), 0), p.in_h);
        int h = min(max(floor_div(midY + p.f_h, p.upy), 0), p.in_h) - inY;
        int filterY = midY + p.f_h - (inY + 1) * p.upy;
        if (p.flip) filterY = p.f_h - 1 - filterY;
        int midX = outX * p.downx + p.upx - 1 - p.pad_x0;
        int inX = min(max(floor_div(midX, p.upx), 0), p.in_w);
        int w = min(max(floor_div(midX + p.f_w, p.upx), 0), p.in_w) - inX;
        int filterX = midX + p.f_w - (inX + 1) * p.upx;
        if (p.flip) filterX = p.f_w - 1 - filterX;
        const long long stepX = (p.flip ? p.upx : -p.upx) * p.fsx;
        const long long stepY = (p.flip ? p.upy : -p.upy) * p.fsy;
        const float* xp = (const float*)p.x + n * p.isn + inY * p.isy + inX * p.isx + c0;
        const float* fp = p.f + filterX * p.fsx + filterY * p.fsy;
        V acc; vzero(acc);
        for (int y = 0; y < h; y++)
        {
            for (int x = 0; x < w; x++)
                vfma(acc, __ldg(reinterpret_cast<const V*>(xp + x * p.isx)), __ldg(fp + x * stepX));
            xp += p.isy;
            fp += stepY;
        }
        nhwc_store<VEC, true>(p, acc, n, c0, outY, outX);
    }
}

//------------------------------------------------------------------------------------------------
// Host side.

static bool dense_nchw(int w, int h, int c, long long sx, long long sy, long long sc, long long sn)
{
    return sx == 1 && sy == w && sc == (long long)w * h && sn == (long long)w * h * c;
}
static bool dense_nhwc(int w, int h, int c, long long sx, long long sy, long long sc, long long sn)
{
    return sc == 1 && sx == c && sy == (long long)w * c && sn == (long long)w * h * c;
}

template <int UPX, int UPY, int DOWNX, int DOWNY, int FW_T, int FH_T>
static int launch_tiled(const FirArgs& a, cudaStream_t stream)
{
    const int fw = FW_T > 0 ? FW_T : a.f_w, fh = FH_T > 0 ? FH_T : a.f_h;
    const int fwp = ceil_div(fw, UPX) * UPX, fhp = ceil_div(fh, UPY) * UPY;
    FirTile t;
    int need_lanes = ceil_div(a.out_w, 4);
    int lg = 0;
    const int lg_max = (DOWNX > 1) ? 4 : 5;      // keep the staged tile of decimating variants under 48 KB
    while ((1 << lg) < need_lanes && lg < lg_max) lg++;
    t.lanes_x_log2 = lg;
    const int lanes_x = 1 << lg;
    t.tile_out_w = 4 * lanes_x;
    int rows = (kFirThreads / lanes_x) * kFirRPT;
    // keep the staged tile modest for tiny-width images
    const int max_rows = 64;
    if (rows > max_rows) rows = max_rows;
    if (rows > ceil_div(a.out_h, kFirRPT) * kFirRPT) rows = ceil_div(a.out_h, kFirRPT) * kFirRPT;
    t.tile_out_h = rows;
    t.tile_in_w = ((t.tile_out_w - 1) * DOWNX + fwp - 1) / UPX + 1;
    t.tile_in_h = ((t.tile_out_h - 1) * DOWNY + fhp - 1) / UPY + 1;
    t.vecs_per_row = (t.tile_in_w + 3 + 3) / 4;
    t.div_vpr = FastDiv((uint32_t)t.vecs_per_row);
    t.tiles_x = ceil_div(a.out_w, t.tile_out_w);
    t.tiles_y = ceil_div(a.out_h, t.tile_out_h);
    t.in_total = (long long)a.in_n * a.in_c * a.in_h * a.in_w;
    t.out_vec_ok = (a.out_w % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0);
    size_t smem = (size_t)(((fhp * fwp + 3) & ~3) + t.tile_in_h * t.vecs_per_row * 4) * sizeof(float);
    if (smem > 96 * 1024) return -1;
    long long blocks = (long long)a.in_n * a.in_c * t.tiles_x * t.tiles_y;
    if (blocks > 0x7fffffffLL) return -1;
    auto kern = (a.eact != 0) ? fir_nchw_tiled<UPX, UPY, DOWNX, DOWNY, FW_T, FH_T, true> : fir_nchw_tiled<UPX, UPY, DOWNX, DOWNY, FW_T, FH_T, false>;
    if (smem > 48 * 1024)
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kern<<<(unsigned)blocks, kFirThreads, smem, stream>>>(a, t);
    return 0;
}

static int dispatch_tiled(const FirArgs& a, cudaStream_t s)
{
    const int ux = a.upx, uy = a.upy, dx = a.downx, dy = a.downy;
    const bool f44 = (a.f_w == 4 && a.f_h == 4);
    if (a.f_w > kFirMaxTaps || a.f_h > kFirMaxTaps) return -1;
#define SGV_FIR_CASE(UX, UY, DX, DY) if (ux == UX && uy == UY && dx == DX && dy == DY) { \
        if (f44) return launch_tiled<UX, UY, DX, DY, 4, 4>(a, s); return launch_tiled<UX, UY, DX, DY, 0, 0>(a, s); }
    SGV_FIR_CASE(1, 1, 1, 1)
    SGV_FIR_CASE(2, 2, 1, 1)
    SGV_FIR_CASE(1, 1, 2, 2)
    SGV_FIR_CASE(2, 1, 1, 1)
    SGV_FIR_CASE(1, 2, 1, 1)
    SGV_FIR_CASE(1, 1, 2, 1)
    SGV_FIR_CASE(1, 1, 1, 2)
#undef SGV_FIR_CASE
    return -1;
}

} // namespace sgv

extern "C" int sgv_upfirdn2d_out_size(int in_size, int up, int pad0, int pad1, int fsize, int down)
{
    return (in_size * up + pad0 + pad1 - fsize + down) / down;
}

extern "C" int sgv_upfirdn2d(const sgv_upfirdn2d_params* p, void* stream_)
{
    using namespace sgv;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGV_CHECK_ARG(p != nullptr, "sgv_upfirdn2d: params is NULL");
    SGV_CHECK_ARG(p->x && p->f && p->y, "sgv_upfirdn2d: x, f and y must be non-NULL");
    SGV_CHECK_ARG(p->dtype == SGV_F32 || p->dtype == SGV_F16 || p->dtype == SGV_F64, "sgv_upfirdn2d: bad dtype %d", p->dtype);
    SGV_CHECK_ARG(p->up_x >= 1 && p->up_y >= 1, "upsampling factor must be at least 1");
    SGV_CHECK_ARG(p->down_x >= 1 && p->down_y >= 1, "downsampling factor must be at least 1");
    SGV_CHECK_ARG(p->f_w >= 1 && p->f_h >= 1, "f must be at least 1x1");
    SGV_CHECK_ARG(p->in_w >= 1 && p->in_h >= 1 && p->in_c >= 1 && p->in_n >= 1, "x must have non-empty extents");
    SGV_CHECK_ARG((long long)p->in_w * p->in_h * p->in_c * p->in_n <= 0x7fffffffLL, "x is too large");
    const int ow = sgv_upfirdn2d_out_size(p->in_w, p->up_x, p->pad_x0, p->pad_x1, p->f_w, p->down_x);
    const int oh = sgv_upfirdn2d_out_size(p->in_h, p->up_y, p->pad_y0, p->pad_y1, p->f_h, p->down_y);
    SGV_CHECK_ARG(ow >= 1 && oh >= 1, "output must be at least 1x1");
    SGV_CHECK_ARG(ow == p->out_w && oh == p->out_h, "out size mismatch: expected %dx%d (w x h), got %dx%d", ow, oh, p->out_w, p->out_h);
    SGV_CHECK_ARG((long long)ow * oh * p->in_c * p->in_n <= 0x7fffffffLL, "output is too large");
    SGV_CHECK_ARG(p->epi_act == 0 || p->epi_act == 1 || p->epi_act == 3, "fused epilogue supports act 0/1/3 only");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;

    FirArgs a;
    a.x = p->x; a.f = p->f; a.y = p->y;
    a.upx = p->up_x; a.upy = p->up_y; a.downx = p->down_x; a.downy = p->down_y;
    a.pad_x0 = p->pad_x0; a.pad_y0 = p->pad_y0; a.flip = p->flip ? 1 : 0; a.gain = p->gain;
    a.in_w = p->in_w; a.in_h = p->in_h; a.in_c = p->in_c; a.in_n = p->in_n;
    a.isx = p->in_stride_x; a.isy = p->in_stride_y; a.isc = p->in_stride_c; a.isn = p->in_stride_n;
    a.f_w = p->f_w; a.f_h = p->f_h; a.fsx = p->f_stride_x; a.fsy = p->f_stride_y;
    a.out_w = ow; a.out_h = oh;
    a.osx = p->out_stride_x; a.osy = p->out_stride_y; a.osc = p->out_stride_c; a.osn = p->out_stride_n;
    a.escale = p->epi_scale; a.ebias = p->epi_bias; a.eact = p->epi_act;
    a.ealpha = p->epi_alpha; a.egain = p->epi_gain; a.eclamp = p->epi_clamp; a.eround = p->epi_round_tf32;
    a.enoise = p->epi_noise; a.ensn = p->epi_noise_stride_n; a.ensy = p->epi_noise_stride_y; a.ensx = p->epi_noise_stride_x;
    SGV_CHECK_ARG(!a.enoise || a.eact != 0, "epi_noise needs a fused epilogue (epi_act != 0)");
    SGV_CHECK_ARG(!a.eround || a.eact != 0, "epi_round_tf32 needs a fused epilogue (epi_act != 0)");
    // the rounding flag is honoured by the two channels_last 4x4 kernels below; any other route rejects it rather than ignore it
    const bool round_ok = p->dtype == SGV_F32 && a.upx == 1 && a.upy == 1 && a.downx == 1 && a.downy == 1 && a.f_w == 4 && a.f_h == 4 && a.in_c % 4 == 0 && a.in_c > 1
                          && dense_nhwc(a.in_w, a.in_h, a.in_c, a.isx, a.isy, a.isc, a.isn) && dense_nhwc(ow, oh, a.in_c, a.osx, a.osy, a.osc, a.osn)
                          && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->y) & 15) == 0;
    SGV_CHECK_ARG(!a.eround || round_ok, "epi_round_tf32: only for float32 channels_last tensors, up = down = 1, 4x4 filter, C %% 4 == 0");

    const long long total = (long long)ow * oh * p->in_c * p->in_n;
    const int sms = num_sms();

    if (p->dtype == SGV_F32)
    {
        const bool x16 = (reinterpret_cast<uintptr_t>(p->x) & 15) == 0;
        const bool y16 = (reinterpret_cast<uintptr_t>(p->y) & 15) == 0;
        const bool in_nchw = dense_nchw(a.in_w, a.in_h, a.in_c, a.isx, a.isy, a.isc, a.isn);
        const bool out_nchw = dense_nchw(ow, oh, a.in_c, a.osx, a.osy, a.osc, a.osn);
        const bool in_nhwc = dense_nhwc(a.in_w, a.in_h, a.in_c, a.isx, a.isy, a.isc, a.isn);
        const bool out_nhwc = dense_nhwc(ow, oh, a.in_c, a.osx, a.osy, a.osc, a.osn);
        if (in_nhwc && out_nhwc && a.in_c > 1)
        {
            const bool v4 = (a.in_c % 4 == 0) && x16 && y16;
            const bool fast = a.upx == 1 && a.upy == 1 && a.downx == a.downy && (a.downx == 1 || a.downx == 2) && a.f_w == 4 && a.f_h == 4;
            const int cvecs = v4 ? a.in_c / 4 : a.in_c;
            if (fast)
            {
                const long long work = (long long)a.in_n * ((oh + 1) / 2) * ow * cvecs;
                const unsigned grid = (unsigned)min((long long)sms * 32, (work + 255) / 256);
                const bool epi = a.eact != 0;
#define SGV_NHWC_FAST(V, D) do { if (epi) fir_nhwc_fast<V, D, 4, 4, true><<<grid, 256, 0, stream>>>(a, work); \
                                 else fir_nhwc_fast<V, D, 4, 4, false><<<grid, 256, 0, stream>>>(a, work); } while (0)
                static const int use_tma = env_int("SGV_FIR_NO_TMA", 0) ? 0 : 1;
                if (v4 && a.downx == 1 && use_tma && a.in_c % 32 == 0 && ow >= kFtTW && oh >= kFtTH)
                {
                    CUtensorMap tm;
                    const uint64_t dims[4] = {(uint64_t)a.in_c, (uint64_t)a.in_w, (uint64_t)a.in_h, (uint64_t)a.in_n};
                    const uint64_t strides[3] = {(uint64_t)a.isx * 4, (uint64_t)a.isy * 4, (uint64_t)a.isn * 4};
                    const uint32_t box[4] = {32, (uint32_t)kFtBoxW, (uint32_t)kFtBoxH, 1};
                    const uint32_t es[4] = {1, 1, 1, 1};
                    rc = make_tmap_f32(&tm, a.x, 4, dims, strides, box, es);
                    if (rc != SGV_OK) return rc;
                    const int tiles_x = ceil_div(ow, kFtTW), tiles_y = ceil_div(oh, kFtTH), cblocks = a.in_c / 32;
                    const long long tiles = (long long)tiles_x * tiles_y * cblocks * a.in_n;
                    SGV_CHECK_ARG(tiles <= 0x7fffffffLL, "too many tiles");
                    const unsigned g3 = (unsigned)min((long long)sms * 2, tiles);
                    SGV_OPT_IN_SMEM(fir_nhwc_tma44<true>, kFtSmem);
                    SGV_OPT_IN_SMEM(fir_nhwc_tma44<false>, kFtSmem);
                    if (epi) fir_nhwc_tma44<true><<<g3, 256, kFtSmem, stream>>>(tm, a, tiles_x, tiles_y, cblocks, (int)tiles);
                    else fir_nhwc_tma44<false><<<g3, 256, kFtSmem, stream>>>(tm, a, tiles_x, tiles_y, cblocks, (int)tiles);
                }
                else if (v4 && a.downx == 1)
                {
                    const int seg = 32;
                    const int nseg = ceil_div(ow, seg);
                    const long long items = (long long)a.in_n * ((oh + 1) / 2) * nseg * cvecs;
                    const unsigned g2 = (unsigned)min((long long)sms * 16, (items + 255) / 256);
                    if (epi) fir_nhwc_slide44<true><<<g2, 256, 0, stream>>>(a, items, seg, nseg);
                    else fir_nhwc_slide44<false><<<g2, 256, 0, stream>>>(a, items, seg, nseg);
                }
                else if (v4) SGV_NHWC_FAST(4, 2);
                else if (a.downx == 1) SGV_NHWC_FAST(1, 1);
                else SGV_NHWC_FAST(1, 2);
#undef SGV_NHWC_FAST
                SGV_LAUNCH_OK("fir_nhwc_fast");
                return SGV_OK;
            }
            if (v4 && a.upx == 2 && a.upy == 2 && a.downx == 1 && a.downy == 1 && a.f_w == 4 && a.f_h == 4
                && a.pad_x0 >= 0 && a.pad_y0 >= 0 && a.pad_x0 % 2 == 0 && a.pad_y0 % 2 == 0)
            {
                const int qw = (ow + 1) / 2, qh = (oh + 1) / 2;
                const long long quads = (long long)a.in_n * qh * qw * cvecs;
                const unsigned gq = (unsigned)min((long long)sms * 32, (quads + 255) / 256);
                fir_nhwc_up2_44<<<gq, 256, 0, stream>>>(a, quads, qw, qh);
                SGV_LAUNCH_OK("fir_nhwc_up2_44");
                return SGV_OK;
            }
            const long long work = (long long)a.in_n * oh * ow * cvecs;
            const unsigned grid = (unsigned)min((long long)sms * 32, (work + 255) / 256);
            if (v4) fir_nhwc_any<4><<<grid, 256, 0, stream>>>(a, work);
            else fir_nhwc_any<1><<<grid, 256, 0, stream>>>(a, work);
            SGV_LAUNCH_OK("fir_nhwc_any");
            return SGV_OK;
        }
        if (in_nchw && out_nchw && x16)
        {
            if (dispatch_tiled(a, stream) == 0)
            {
                SGV_LAUNCH_OK("fir_nchw_tiled");
                return SGV_OK;
            }
        }
    }

    // generic path (any dtype, any strides)
    const int channels_fast = (a.isc == 1 && a.in_c > 1) ? 1 : 0;
    const unsigned grid = (unsigned)min((long long)sms * 16, (total + 255) / 256);
    if (p->dtype == SGV_F32) fir_generic<float><<<grid, 256, 0, stream>>>(a, channels_fast, total);
    else if (p->dtype == SGV_F64) fir_generic<double><<<grid, 256, 0, stream>>>(a, channels_fast, total);
    else fir_generic<__half><<<grid, 256, 0, stream>>>(a, channels_fast, total);
    SGV_LAUNCH_OK("fir_generic");
    return SGV_OK;
}
