// Weight gradient of the modulated convolution on tcgen05 tensor cores (sm_100a), split-K over pixels.
//
//   dw[t][o][i] += sum_{n,p} (g[n, p*gs + dg_t, o] * g_scale[n,o]) * (x[n, p*xs + dx_t, i] * x_scale[n,i])     (include/sgv_b200_conv.h)
//
// GEMM view per tap t: D[o][i] (M = 128 output channels, N = BN input channels) accumulated over K = pixels.  With NHWC
// tensors both operands are "MN-major" (channels contiguous, pixels strided): a TMA box {32 channels x 32 pixels} lands as
// 32 rows of 128 bytes = eight 4-pixel x 32-channel swizzle atoms (TMA SWIZZLE_128B_ATOM_32B), the canonical MN-major
// SWIZZLE_128B_BASE32B layout of the UMMA descriptor (the only MN-major layout for 32-bit operands).  A 5-D tensor map {32, W, H, N, C/32} delivers all channel blocks of a stage with ONE bulk load
// ([c_blk][pixel][32 ch] in shared memory); taps are shifted / strided boxes with TMA zero fill at the borders.
//
// Same warp roles as conv_tf32.cu: warp 0 TMA producer, warp 1 MMA issuer (+TMEM owner), warps 2-5 scale the staged rows
// by g_scale / x_scale and round to TF32 in place, then run the epilogue (TMEM -> red.global.add.v4.f32 into dw).
// Grid: (M tiles x N tiles, taps, K splits).  Replaces aten::cudnn_convolution_backward_weight (conv2d_gradfix.py:140-148).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/sgv_b200_conv.h"

#include <stdlib.h>
#include <string.h>

namespace sgv {

using namespace ptx;

// g_lo / x_lo: stage the TF32 residual tf32(v - tf32(v)) of that operand instead of tf32(v) (one pass of the tf32x3 mode)
int conv2d_wgrad_tf32_v2(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query);
int conv2d_wgrad_tf32_s64(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query);     // wgrad_tf32_s64.cu

constexpr int kWgThreads = 192;
constexpr int kWgM = 128;
constexpr int kWgKc = 32;                       // pixels per k-step
constexpr int kWgATile = kWgM * kWgKc * 4;      // 16 KB: [4 channel blocks][32 pixels][32 ch]

struct WgArgs
{
    float* dw; const float* g_scale; const float* x_scale;
    int n, cin, cout, out_h, out_w;
    int g_stride, x_stride, ntaps;
    int g_dy[SGV_CONV_MAX_TAPS], g_dx[SGV_CONV_MAX_TAPS], x_dy[SGV_CONV_MAX_TAPS], x_dx[SGV_CONV_MAX_TAPS];
    int tw, th, tn;                   // pixel box: tw*th*tn == 32
    int tiles_x, tiles_y, tiles_nb;
    int mtiles, ktiles, ksplit;
    int dw_slot[SGV_CONV_MAX_TAPS];   // dw block each tap accumulates into
    int g_lo, x_lo;                   // tf32x3 pass: stage the TF32 residual of that operand
};

template <int BN, int STAGES>
struct WgSmem
{
    static constexpr int kBTile = BN * kWgKc * 4;
    static constexpr int kStage = kWgATile + kBTile;
    static constexpr int kBarOffset = STAGES * kStage;
    static constexpr int kTotal = kBarOffset + (3 * STAGES + 1) * 8 + 16 + 1024;
};

// scales one staged 128-byte row (32 channels of one pixel) by sc[0..31] and rounds to TF32, in place
__device__ __forceinline__ void wg_transform_row(uint8_t* rowp_generic, int row, const float* __restrict__ sc, bool lo)
{
    const uint32_t rowp = smem_u32(rowp_generic);
    const int flip = (row >> 2) & 1;          // chunk order that keeps a quarter-warp's 8 rows on 8 distinct bank groups
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const int jj = j ^ flip;
        // SWIZZLE_128B_ATOM_32B: logical 32-byte chunk (jj >> 1) lives at physical chunk (jj >> 1) ^ (row & 3)
        v[j] = lds128(rowp + (uint32_t)(((((jj >> 1) ^ (row & 3)) << 1) | (jj & 1)) << 4));
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
    {
        const int jj = j ^ flip;
        const float4 s = sc ? __ldg(reinterpret_cast<const float4*>(sc) + jj) : make_float4(1.f, 1.f, 1.f, 1.f);
        v[j].x = __fmul_rn(v[j].x, s.x); v[j].y = __fmul_rn(v[j].y, s.y); v[j].z = __fmul_rn(v[j].z, s.z); v[j].w = __fmul_rn(v[j].w, s.w);
        if (lo) { v[j].x = tf32_lo(v[j].x); v[j].y = tf32_lo(v[j].y); v[j].z = tf32_lo(v[j].z); v[j].w = tf32_lo(v[j].w); }
        else { v[j].x = tf32_rn(v[j].x); v[j].y = tf32_rn(v[j].y); v[j].z = tf32_rn(v[j].z); v[j].w = tf32_rn(v[j].w); }
        sts128(rowp + (uint32_t)(((((jj >> 1) ^ (row & 3)) << 1) | (jj & 1)) << 4), v[j]);
    }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tf32_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x, const WgArgs p)
{
    using L = WgSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* ready_bar = full_bar + STAGES;
    uint64_t* empty_bar = ready_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % p.mtiles, nt = blockIdx.x / p.mtiles;
    const int m0 = mt * kWgM;                 // first output channel (rows of dw)
    const int c0 = nt * BN;                   // first input channel (cols of dw)
    const int tap = blockIdx.y;
    const int per = (p.ktiles + p.ksplit - 1) / p.ksplit;
    const int kt0 = blockIdx.z * per;
    const int kt1 = min(kt0 + per, p.ktiles);
    const int ksteps = kt1 - kt0;             // may be <= 0 for trailing splits

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_g);
        prefetch_tmap(&tmap_x);
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(ready_bar + s, 4); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    constexpr int kTmemCols = BN < 32 ? 32 : BN;
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (ksteps > 0)
    {
        if (warp == 0)
        {
            if (elect_one())
            {
                int stage = 0; uint32_t phase = 0;
                for (int kt = kt0; kt < kt1; kt++)
                {
                    int r = kt;
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y; r /= p.tiles_y;
                    const int px0 = tx * p.tw, py0 = ty * p.th, nb0 = r * p.tn;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStage;
                    mbar_expect_tx(full_bar + stage, L::kStage);
                    tma_load_5d(sa, &tmap_g, full_bar + stage, 0, px0 * p.g_stride + p.g_dx[tap], py0 * p.g_stride + p.g_dy[tap], nb0, m0 / 32);
                    tma_load_5d(sa + kWgATile, &tmap_x, full_bar + stage, 0, px0 * p.x_stride + p.x_dx[tap], py0 * p.x_stride + p.x_dy[tap], nb0, c0 / 32);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        else if (warp == 1)
        {
            constexpr uint32_t idesc = umma_idesc_tf32(kWgM, BN, 1, 1);     // both operands MN-major
            int stage = 0; uint32_t phase = 0;
            for (int ks = 0; ks < ksteps; ks++)
            {
                mbar_wait(ready_bar + stage, phase);
                tc_fence_after();
                if (elect_one())
                {
                    const uint32_t sa = smem_u32(smem + stage * L::kStage);
#pragma unroll
                    for (int k = 0; k < kWgKc / 8; k++)         // one 8-pixel swizzle atom (1024 B) per instruction
                    {
                        const uint64_t da = umma_desc_mn_sw128_32b(sa + k * 1024, kWgKc * 128, 512);
                        const uint64_t db = umma_desc_mn_sw128_32b(sa + kWgATile + k * 1024, kWgKc * 128, 512);
                        mma_tf32(tmem_base, da, db, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(empty_bar + stage);
                    if (ks == ksteps - 1) mma_commit(accum_bar);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
        else
        {
            const int tid = threadIdx.x - 64;                  // 0..127
            const int box_hw = p.tw * p.th;
            {
                int stage = 0; uint32_t phase = 0;
                for (int kt = kt0; kt < kt1; kt++)
                {
                    const int nb0 = (kt / (p.tiles_x * p.tiles_y)) * p.tn;
                    mbar_wait(full_bar + stage, phase);
                    uint8_t* sa = smem + stage * L::kStage;
                    {   // gradient rows: row = blk*32 + pixel, 128 rows = one per thread
                        const int blk = tid >> 5, pix = tid & 31;
                        const int n = min(nb0 + pix / box_hw, p.n - 1);
                        const int ch = m0 + blk * 32;
                        const float* sc = (p.g_scale && ch < p.cout) ? p.g_scale + (long long)n * p.cout + ch : nullptr;
                        if (ch < p.cout) wg_transform_row(sa + tid * 128, tid, sc, p.g_lo != 0);
                    }
#pragma unroll
                    for (int rr = 0; rr < (BN + 127) / 128; rr++)
                    {
                        const int row = rr * 128 + tid;
                        if (row < BN)
                        {
                            const int blk = row >> 5, pix = row & 31;
                            const int n = min(nb0 + pix / box_hw, p.n - 1);
                            const int ch = c0 + blk * 32;
                            const float* sc = (p.x_scale && ch < p.cin) ? p.x_scale + (long long)n * p.cin + ch : nullptr;
                            if (ch < p.cin) wg_transform_row(sa + kWgATile + row * 128, row, sc, p.x_lo != 0);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready_bar + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            // ---- epilogue: accumulate the partial [128 x BN] block into dw[tap] ----
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int o = m0 + q * 32 + lane;
            float* drow = p.dw + ((long long)p.dw_slot[tap] * p.cout + o) * p.cin + c0;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; cc++)
            {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v);
                tmem_ld_wait();
                if (o < p.cout && c0 + cc * 32 < p.cin)
                {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        atomicAdd(reinterpret_cast<float4*>(drow + cc * 32 + j * 4),
                                  make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN, int STAGES>
static int launch_wgrad(const CUtensorMap& tg, const CUtensorMap& tx, const WgArgs& a, dim3 grid, cudaStream_t stream)
{
    using L = WgSmem<BN, STAGES>;
    auto kern = wgrad_tf32_kernel<BN, STAGES>;
    SGV_OPT_IN_SMEM(kern, L::kTotal);
    kern<<<grid, kWgThreads, L::kTotal, stream>>>(tg, tx, a);
    SGV_LAUNCH_OK("wgrad_tf32_kernel");
    return SGV_OK;
}

static int make_blocked_tmap(CUtensorMap* m, const float* base, int c, int w, int h, int n, int tw, int th, int tn, int stride, int blocks)
{
    const uint64_t dims[5] = {32, (uint64_t)w, (uint64_t)h, (uint64_t)n, (uint64_t)(c / 32)};
    const uint64_t strides[4] = {(uint64_t)c * 4, (uint64_t)w * c * 4, (uint64_t)h * w * c * 4, 128};
    const uint32_t box[5] = {32, (uint32_t)(tw * stride), (uint32_t)(th * stride), (uint32_t)tn, (uint32_t)blocks};
    const uint32_t es[5] = {1, (uint32_t)stride, (uint32_t)stride, 1, 1};
    return make_tmap_f32(m, base, 5, dims, strides, box, es, /*atom32=*/true);
}

} // namespace sgv

static int wgrad_pass(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query);

static int wgrad_dispatch(const sgv_wgrad_params* p, cudaStream_t stream, sgv_wgrad_variant* query)
{
    using namespace sgv;
    SGV_CHECK_ARG(p != nullptr, "sgv_conv2d_wgrad_tf32: params is NULL");
    SGV_CHECK_ARG(p->precision == 0 || p->precision == 1, "precision must be 0 (tf32x1) or 1 (tf32x3)");
    if (p->precision == 0) return wgrad_pass(p, 0, 0, stream, query);
    // tf32x3: dw += g_hi*x_hi + g_lo*x_hi + g_hi*x_lo — three passes of the same kernel into the same accumulation buffer, each staging
    // the hi or lo part of its operands (the staging warps form the parts from the fp32 values in shared memory)
    SGV_CHECK_ARG(!p->g_ready && !p->x_ready, "tf32x3 mode needs unrounded operands: g_ready / x_ready must be 0");
    int rc = wgrad_pass(p, 1, 0, stream, query);           // small terms first
    if (rc != SGV_OK || query) { if (query) query->passes = 3; return rc; }
    rc = wgrad_pass(p, 0, 1, stream, nullptr);
    if (rc != SGV_OK) return rc;
    return wgrad_pass(p, 0, 0, stream, nullptr);
}

extern "C" int sgv_conv2d_wgrad_tf32(const sgv_wgrad_params* p, void* stream_)
{
    return wgrad_dispatch(p, (cudaStream_t)stream_, nullptr);
}

extern "C" int sgv_conv2d_wgrad_tf32_variant(const sgv_wgrad_params* p, sgv_wgrad_variant* out)
{
    SGV_CHECK_ARG(out != nullptr, "sgv_conv2d_wgrad_tf32_variant: out is NULL");
    memset(out, 0, sizeof(*out));
    return wgrad_dispatch(p, nullptr, out);
}

static int wgrad_pass(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query)
{
    using namespace sgv;
    SGV_CHECK_ARG(p->g && p->x && p->dw, "sgv_conv2d_wgrad_tf32: g, x and dw must be non-NULL");
    SGV_CHECK_ARG(p->cin >= 32 && p->cin % 32 == 0 && p->cout >= 32 && p->cout % 32 == 0, "cin and cout must be multiples of 32 (got %d, %d)", p->cin, p->cout);
    SGV_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= SGV_CONV_MAX_TAPS, "ntaps must be in [1, %d]", SGV_CONV_MAX_TAPS);
    SGV_CHECK_ARG((p->g_stride == 1 || p->g_stride == 2) && (p->x_stride == 1 || p->x_stride == 2), "strides must be 1 or 2");
    SGV_CHECK_ARG(p->n >= 1 && p->out_h >= 1 && p->out_w >= 1, "extents must be positive");
    SGV_CHECK_ARG((reinterpret_cast<uintptr_t>(p->g) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->dw) & 15) == 0, "g, x and dw must be 16-byte aligned");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;

    static const bool force_v1 = env_int("SGV_WGRAD_V1", 0) != 0;
    if (!force_v1)
    {
        rc = conv2d_wgrad_tf32_s64(p, g_lo, x_lo, stream, query);        // 64 output channels, full 3x3: stacked-M kernel
        if (rc != SGV_ERR_UNSUPPORTED) return rc;
        rc = conv2d_wgrad_tf32_v2(p, g_lo, x_lo, stream, query);
        if (rc != SGV_ERR_UNSUPPORTED) return rc;
    }
    if (p->x_stride_x != 0) return sgv::fail(SGV_ERR_UNSUPPORTED, "strided x views are only supported by the grouped-tap kernel (stride 1, out_w >= 8, out_h >= 4)");

    WgArgs a;
    a.dw = p->dw; a.g_scale = p->g_scale; a.x_scale = p->x_scale;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    a.g_stride = p->g_stride; a.x_stride = p->x_stride; a.ntaps = p->ntaps;
    for (int t = 0; t < SGV_CONV_MAX_TAPS; t++) { a.g_dy[t] = p->g_dy[t]; a.g_dx[t] = p->g_dx[t]; a.x_dy[t] = p->x_dy[t]; a.x_dx[t] = p->x_dx[t]; a.dw_slot[t] = p->use_dw_slot ? p->dw_slot[t] : t; }
    // 32-pixel box, spilling into the batch dimension for tiny planes
    int tw = 1; while (tw * 2 <= p->out_w && tw < 8) tw *= 2;
    int th = 32 / tw; { int ph = 1; while (ph < p->out_h) ph *= 2; if (th > ph) th = ph; }
    int tn = 32 / (tw * th);
    a.tw = tw; a.th = th; a.tn = tn;
    a.tiles_x = ceil_div(p->out_w, tw); a.tiles_y = ceil_div(p->out_h, th); a.tiles_nb = ceil_div(p->n, tn);
    a.ktiles = a.tiles_x * a.tiles_y * a.tiles_nb;
    const int bn = (p->cin % 256 == 0) ? 256 : (p->cin % 128 == 0) ? 128 : (p->cin % 64 == 0) ? 64 : 32;
    a.mtiles = ceil_div(p->cout, kWgM);
    const int base_ctas = a.mtiles * (p->cin / bn) * p->ntaps;
    int ksplit = ceil_div(2 * num_sms(), base_ctas);
    if (ksplit > a.ktiles) ksplit = a.ktiles;
    if (ksplit < 1) ksplit = 1;
    a.ksplit = ksplit;
    a.g_lo = g_lo; a.x_lo = x_lo;
    if (query)
    {
        query->kernel = 1; query->nt = bn; query->stages = bn == 256 ? 4 : bn == 128 ? 6 : 8; query->ksplit = ksplit; query->passes = 1;
        return SGV_OK;
    }

    CUtensorMap tg, tx;
    rc = make_blocked_tmap(&tg, p->g, p->cout, p->gw, p->gh, p->n, tw, th, tn, p->g_stride, kWgM / 32);
    if (rc != SGV_OK) return rc;
    rc = make_blocked_tmap(&tx, p->x, p->cin, p->xw, p->xh, p->n, tw, th, tn, p->x_stride, bn / 32);
    if (rc != SGV_OK) return rc;
    dim3 grid((unsigned)(a.mtiles * (p->cin / bn)), (unsigned)p->ntaps, (unsigned)ksplit);
    switch (bn)
    {
        case 256: return launch_wgrad<256, 4>(tg, tx, a, grid, stream);
        case 128: return launch_wgrad<128, 6>(tg, tx, a, grid, stream);
        case 64:  return launch_wgrad<64, 8>(tg, tx, a, grid, stream);
        default:  return launch_wgrad<32, 8>(tg, tx, a, grid, stream);
    }
}

// =====================================================================================================================
// v2 (stride-1 correlations, e.g. 3x3 pad 1): one CTA owns a GROUP of taps that share dy (one kernel row: up to 3 taps).
//   Each tap has its own NT-column TMEM accumulator (3 x 128 = 384 columns).  Per k-step (an 8x4 pixel tile) the gradient
//   tile is loaded once and the input once as a (8 + dx-span) x 4 patch of NT channels; tap t's B operand is that patch
//   addressed through a shifted MN-major descriptor (start row = y*PW + dx_t; 8 consecutive pixel rows = one image row).
//   tcgen05 swizzles on absolute address bits, so unaligned starts are legal (profiles/umma_probe_r1.txt).
//   Why groups of 3 and NT = 128: an SS-mode tcgen05.mma re-reads its A tile (128 x 8 x 4 B) from shared memory for every
//   instruction, so N must be >= 128 for the 128 B/clk shared-memory port to keep up (measured: an all-9-taps variant with
//   N = 32 ran at 110 TFLOP/s); TMEM (512 columns) then holds 3 taps.
// =====================================================================================================================
namespace sgv {

constexpr int kW2GTile = 128 * 32 * 4;            // 16 KB
constexpr int kW2MaxGroupTaps = 3;
constexpr int kWg2Threads = 64 + 256;          // TMA warp, MMA warp, 8 transform/epilogue warps

struct Wg2Args
{
    float* dw; const float* g_scale; const float* x_scale;
    int n, cin, cout, out_h, out_w;
    int ngroups;
    int grp_ntaps[SGV_CONV_MAX_TAPS];                       // taps in group
    int grp_dy[SGV_CONV_MAX_TAPS];                          // common dy of the group
    int grp_tap[SGV_CONV_MAX_TAPS][kW2MaxGroupTaps];        // global tap index (row of dw)
    int grp_col[SGV_CONV_MAX_TAPS][kW2MaxGroupTaps];        // dx_t - dx_min: column offset inside the patch
    int dx_min, pw;
    int tiles_x, tiles_y, mtiles, ktiles, ksplit;
    int g_ready, x_ready;     // operand needs no staging pass
    int g_lo, x_lo;           // tf32x3 pass: stage the TF32 residual of that operand
    int debug;      // ablation switches, only honoured by -DSGV_ABLATION builds (profiles/wgrad_ablation_r1.txt): 1 skip transform, 2 skip epilogue, 4 skip MMAs
};

template <int NT, int STAGES, bool PAIR = false>
struct Wg2Smem
{
    static constexpr int kXBlock = 10 * 4 * 128;                  // one 32-channel block of the (8+2) x 4 patch (5120 B, multiple of the 512 B swizzle atom)
    static constexpr int kNLocal = PAIR ? NT / 2 : NT;            // CTA pair: each CTA holds (stages, reads) half of the N tile's channel blocks
    static constexpr int kXTile = ((kNLocal / 32) * kXBlock + 1023) & ~1023;
    static constexpr int kStage = kW2GTile + kXTile;
    static constexpr int kBarOffset = STAGES * kStage;
    static constexpr int kTotal = kBarOffset + (3 * STAGES + 1) * 8 + 16 + 1024;
};

// PAIR: the CTAs (2i, 2i+1) of a 2-CTA cluster own the output-channel tiles (mt, mt + 1) of the SAME input-channel tile and tap group and
// issue ONE tcgen05.mma.cta_group::2 of M = 256 per k-row: each CTA supplies its own gradient tile (A) and HALF of the input patch's channel
// blocks (B) — half the patch TMA, half the staging work and half the B-operand reads per CTA (the shared-memory port is this kernel's
// limiter: DESIGN.md §4).  Leader-issued MMAs, remote barrier arrivals and multicast commits as in conv_tf32_v3.cu.
template <int NT, int STAGES, bool PAIR>
__global__ void __launch_bounds__(kWg2Threads, 1)
wgrad_tf32_v2_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x, const Wg2Args p)
{
    using L = Wg2Smem<NT, STAGES, PAIR>;
    constexpr int NL = L::kNLocal;
    const int crank = PAIR ? (int)cluster_ctarank() : 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* ready_bar = full_bar + STAGES;
    uint64_t* empty_bar = ready_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x % p.mtiles, nt = blockIdx.x / p.mtiles;
    const int m0 = mt * kWgM, c0 = nt * NT;
    const int grp = blockIdx.z;
    const int gtaps = p.grp_ntaps[grp];
    const int per = (p.ktiles + p.ksplit - 1) / p.ksplit;
    const int kt0 = blockIdx.y * per;
    const int kt1 = min(kt0 + per, p.ktiles);
    const int ksteps = kt1 - kt0;
    const int xblock = p.pw * 4 * 128;                         // bytes of one 32-channel block of the patch
    const uint32_t x_bytes = (uint32_t)(xblock * (NL / 32));

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_g);
        prefetch_tmap(&tmap_x);
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(ready_bar + s, PAIR ? 16 : 8); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    constexpr int kCols = (kW2MaxGroupTaps * NT) <= 256 ? 256 : 512;
    if (warp == 1)
    {
        if (PAIR) { tmem_alloc_pair(tmem_slot, kCols); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, kCols); tmem_relinquish(); }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (ksteps > 0)
    {
        if (warp == 0)
        {
            if (elect_one())
            {
                int stage = 0; uint32_t phase = 0;
                for (int kt = kt0; kt < kt1; kt++)
                {
                    int r = kt;
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y; r /= p.tiles_y;
                    const int px0 = tx * 8, py0 = ty * 4, n = r;
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* sg = smem + stage * L::kStage;
                    mbar_expect_tx(full_bar + stage, kW2GTile + x_bytes);
                    tma_load_5d(sg, &tmap_g, full_bar + stage, 0, px0, py0, n, m0 / 32);
                    tma_load_5d(sg + kW2GTile, &tmap_x, full_bar + stage, 0, px0 + p.dx_min, py0 + p.grp_dy[grp], n, c0 / 32 + crank * (NL / 32));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        else if (warp == 1)
        {
            constexpr uint32_t idesc = umma_idesc_tf32(PAIR ? 2 * kWgM : kWgM, NT, 1, 1);
            if (elect_one() && crank == 0)      // one thread (of the leader CTA) runs the whole issue loop (no per-step election / reconvergence)
            {
                int stage = 0; uint32_t phase = 0;
                for (int ks = 0; ks < ksteps; ks++)
                {
                    mbar_wait(ready_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t sg = smem_u32(smem + stage * L::kStage);
                    const uint32_t sx = sg + kW2GTile;
                    for (int t = 0; t < gtaps; t++)
                    {
#pragma unroll
                        for (int k = 0; k < 4; k++)        // image row k of the 8x4 tile = 8 consecutive pixel rows
                        {
                            if (SGV_ABL(p.debug, 4)) continue;
                            const uint64_t da = umma_desc_mn_sw128_32b(sg + k * 1024, 32 * 128, 512);
                            const uint64_t db = umma_desc_mn_sw128_32b(sx + (uint32_t)(p.grp_col[grp][t] + k * p.pw) * 128u, (uint32_t)xblock, 512);
                            if (PAIR) mma_tf32_pair(tmem_base + (uint32_t)(t * NT), da, db, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                            else mma_tf32(tmem_base + (uint32_t)(t * NT), da, db, idesc, (ks > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    if (PAIR) mma_commit_pair_mc(empty_bar + stage, 3); else mma_commit(empty_bar + stage);
                    if (ks == ksteps - 1) { if (PAIR) mma_commit_pair_mc(accum_bar, 3); else mma_commit(accum_bar); }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            __syncwarp();
        }
        else
        {
            // 8 transform warps.  Unified row space: rows [0,128) = gradient tile, rows [128, 128 + xrows) = input patch.
            // A thread owns rows tid and tid + 256; their channel block is fixed, so the 32 scale factors stay in registers and
            // are reloaded only when the sample index changes.
            const int tid = threadIdx.x - 64;                   // 0..255
            const int xrows_blk = p.pw * 4;
            const int total_rows = 128 + xrows_blk * (NL / 32);
            float sv[2][32];
            int cur_n = -1;
            const FastDiv div_plane((uint32_t)(p.tiles_x * p.tiles_y));
            int rr[2]; const float* sbase[2]; int sstride[2]; bool act[2]; bool lo[2];
#pragma unroll
            for (int s = 0; s < 2; s++)
            {
                rr[s] = tid + s * 256;
                act[s] = rr[s] < total_rows;
                sbase[s] = nullptr; sstride[s] = 0;
                lo[s] = (rr[s] < 128 ? p.g_lo : p.x_lo) != 0;
                if (act[s])
                {
                    if (rr[s] < 128) { const int ch = m0 + (rr[s] >> 5) * 32; act[s] = ch < p.cout && !p.g_ready; if (p.g_scale) { sbase[s] = p.g_scale + ch; sstride[s] = p.cout; } }
                    else { const int ch = c0 + crank * NL + ((rr[s] - 128) / xrows_blk) * 32; act[s] = ch < p.cin && !p.x_ready; if (p.x_scale) { sbase[s] = p.x_scale + ch; sstride[s] = p.cin; } }
                }
#pragma unroll
                for (int j = 0; j < 32; j++) sv[s][j] = 1.f;
            }
            {
                int stage = 0; uint32_t phase = 0;
                for (int kt = kt0; kt < kt1; kt++)
                {
                    const int n = min((int)div_plane.div((uint32_t)kt), p.n - 1);
                    if (n != cur_n)
                    {
                        cur_n = n;
#pragma unroll
                        for (int s = 0; s < 2; s++)
                            if (act[s] && sbase[s])
                            {
                                const float4* sp = reinterpret_cast<const float4*>(sbase[s] + (long long)n * sstride[s]);
#pragma unroll
                                for (int j = 0; j < 8; j++) { float4 v = __ldg(sp + j); sv[s][4 * j] = v.x; sv[s][4 * j + 1] = v.y; sv[s][4 * j + 2] = v.z; sv[s][4 * j + 3] = v.w; }
                            }
                    }
                    mbar_wait(full_bar + stage, phase);
                    uint8_t* sg = smem + stage * L::kStage;
#pragma unroll
                    for (int s = 0; s < 2; s++)
                    {
                        if (!act[s] || SGV_ABL(p.debug, 1)) continue;
                        const int row = rr[s] < 128 ? rr[s] : rr[s] - 128;
                        const uint32_t rowp = smem_u32(sg) + (uint32_t)(rr[s] < 128 ? 0 : kW2GTile) + (uint32_t)row * 128u;
                        // logical 16-byte chunk jj = j ^ bit2(row): the 8 rows a quarter-warp touches then hit 8 distinct physical
                        // chunks of the 32-byte-atom swizzle (row & 3 only permutes 32 B pairs) -> no shared-memory bank conflicts
                        const int flip = (row >> 2) & 1;
                        float4 v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++)
                        {
                            const int jj = j ^ flip;
                            v[j] = lds128(rowp + (uint32_t)(((((jj >> 1) ^ (row & 3)) << 1) | (jj & 1)) << 4));
                        }
#pragma unroll
                        for (int j = 0; j < 8; j++)
                        {
                            const int jj = j ^ flip;
                            const float s0 = flip ? sv[s][4 * (j ^ 1) + 0] : sv[s][4 * j + 0];
                            const float s1 = flip ? sv[s][4 * (j ^ 1) + 1] : sv[s][4 * j + 1];
                            const float s2 = flip ? sv[s][4 * (j ^ 1) + 2] : sv[s][4 * j + 2];
                            const float s3 = flip ? sv[s][4 * (j ^ 1) + 3] : sv[s][4 * j + 3];
                            v[j].x = __fmul_rn(v[j].x, s0); v[j].y = __fmul_rn(v[j].y, s1); v[j].z = __fmul_rn(v[j].z, s2); v[j].w = __fmul_rn(v[j].w, s3);
                            if (lo[s]) { v[j].x = tf32_lo(v[j].x); v[j].y = tf32_lo(v[j].y); v[j].z = tf32_lo(v[j].z); v[j].w = tf32_lo(v[j].w); }
                            else { v[j].x = tf32_rn(v[j].x); v[j].y = tf32_rn(v[j].y); v[j].z = tf32_rn(v[j].z); v[j].w = tf32_rn(v[j].w); }
                            sts128(rowp + (uint32_t)(((((jj >> 1) ^ (row & 3)) << 1) | (jj & 1)) << 4), v[j]);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_leader(ready_bar + stage); else mbar_arrive(ready_bar + stage); }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int q = warp & 3;
            const int half = (warp - 2) >> 2;                   // two warps share a TMEM lane quarter and split the column chunks
            const int o = m0 + q * 32 + lane;
#pragma unroll 1
            for (int t = 0; t < gtaps; t++)
            {
                float* drow = p.dw + ((long long)p.grp_tap[grp][t] * p.cout + o) * p.cin + c0;
#pragma unroll 1
                for (int cc = half; cc < NT / 32; cc += 2)
                {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * NT + cc * 32), v);
                    tmem_ld_wait();
                    if (o < p.cout && c0 + cc * 32 < p.cin && !SGV_ABL(p.debug, 2))
                    {
#pragma unroll
                        for (int j = 0; j < 8; j++)
                            atomicAdd(reinterpret_cast<float4*>(drow + cc * 32 + j * 4),
                                      make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    }
                }
            }
        }
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, kCols); else tmem_dealloc(tmem_base, kCols); }
}

template <int NT, int STAGES, bool PAIR = false>
static int launch_wgrad_v2(const CUtensorMap& tg, const CUtensorMap& tx, const Wg2Args& a, dim3 grid, cudaStream_t stream)
{
    using L = Wg2Smem<NT, STAGES, PAIR>;
    auto kern = wgrad_tf32_v2_kernel<NT, STAGES, PAIR>;
    SGV_OPT_IN_SMEM(kern, L::kTotal);
    if (PAIR)
    {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.gridDim = grid; cfg.blockDim = dim3(kWg2Threads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SGV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tg, tx, a));
    }
    else
        kern<<<grid, kWg2Threads, L::kTotal, stream>>>(tg, tx, a);
    SGV_LAUNCH_OK("wgrad_tf32_v2_kernel");
    return SGV_OK;
}

int conv2d_wgrad_tf32_v2(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query)
{
    if (p->g_stride != 1 || p->x_stride != 1 || p->out_w < 8 || p->out_h < 4) return SGV_ERR_UNSUPPORTED;
    int dx_min = p->x_dx[0], dx_max = p->x_dx[0];
    for (int t = 0; t < p->ntaps; t++)
    {
        if (p->g_dy[t] != 0 || p->g_dx[t] != 0) return SGV_ERR_UNSUPPORTED;
        dx_min = min(dx_min, p->x_dx[t]); dx_max = max(dx_max, p->x_dx[t]);
    }
    if (dx_max - dx_min > 2) return SGV_ERR_UNSUPPORTED;
    Wg2Args a;
    memset(&a, 0, sizeof(a));
    a.dw = p->dw; a.g_scale = p->g_scale; a.x_scale = p->x_scale;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    a.dx_min = dx_min; a.pw = 8 + (dx_max - dx_min);
    // group taps by dy
    a.ngroups = 0;
    for (int t = 0; t < p->ntaps; t++)
    {
        int g = -1;
        for (int j = 0; j < a.ngroups; j++) if (a.grp_dy[j] == p->x_dy[t] && a.grp_ntaps[j] < kW2MaxGroupTaps) { g = j; break; }
        if (g < 0) { g = a.ngroups++; a.grp_dy[g] = p->x_dy[t]; a.grp_ntaps[g] = 0; }
        a.grp_tap[g][a.grp_ntaps[g]] = p->use_dw_slot ? p->dw_slot[t] : t;
        a.grp_col[g][a.grp_ntaps[g]] = p->x_dx[t] - dx_min;
        a.grp_ntaps[g]++;
    }
    a.tiles_x = ceil_div(p->out_w, 8); a.tiles_y = ceil_div(p->out_h, 4);
    a.ktiles = a.tiles_x * a.tiles_y * p->n;
    a.mtiles = ceil_div(p->cout, kWgM);
    const int nt = (p->cin % 128 == 0) ? 128 : (p->cin % 64 == 0) ? 64 : 32;
    const int base_ctas = a.mtiles * (p->cin / nt) * a.ngroups;
    // one CTA is resident per SM: size the split so that the grid is (just under) a whole number of waves, and keep >= 48
    // k-steps per CTA so the prologue/epilogue (TMEM alloc, 3 x NT-column atomics) stays amortised
    int waves = 1;
    while (waves < 4 && a.ktiles / max(1, (waves * num_sms()) / base_ctas) > 1024) waves++;
    int ksplit = (waves * num_sms()) / base_ctas;
    if (ksplit < 1) ksplit = 1;
    while (ksplit > 1 && a.ktiles / ksplit < 48) ksplit--;
    if (ksplit > a.ktiles) ksplit = a.ktiles;
    a.ksplit = ksplit;
    static const int debug_flags = env_int("SGV_WG_DEBUG", 0);
    a.debug = debug_flags;
    a.g_ready = p->g_ready && !p->g_scale && !p->precision; a.x_ready = p->x_ready && !p->x_scale && !p->precision;
    a.g_lo = g_lo; a.x_lo = x_lo;

    // CTA pairs (tcgen05 cta_group::2): two output-channel tiles of the same (input-channel tile, tap group, K split) — needs an even number of
    // M tiles (adjacent blockIdx.x = adjacent M tiles form the cluster) and the 128-column N tile; SGV_WGRAD_PAIR=0 disables
    static const int pair_ok = env_int("SGV_WGRAD_PAIR", 1);
    const bool pair = pair_ok && nt == 128 && a.mtiles % 2 == 0 && p->cout % 256 == 0;
    if (query)
    {
        query->kernel = 2; query->nt = nt; query->stages = nt == 128 ? (pair ? 6 : 5) : nt == 64 ? 7 : 8; query->ksplit = a.ksplit; query->passes = 1;
        return SGV_OK;
    }
    CUtensorMap tg, tx;
    int rc = make_blocked_tmap(&tg, p->g, p->cout, p->gw, p->gh, p->n, 8, 4, 1, 1, kWgM / 32);
    if (rc != SGV_OK) return rc;
    {
        const uint64_t dims[5] = {32, (uint64_t)p->xw, (uint64_t)p->xh, (uint64_t)p->n, (uint64_t)(p->cin / 32)};
        const bool view = p->x_stride_x != 0;
        const uint64_t strides[4] = {(uint64_t)(view ? p->x_stride_x : p->cin) * 4, (uint64_t)(view ? p->x_stride_y : (int64_t)p->xw * p->cin) * 4,
                                     (uint64_t)(view ? p->x_stride_n : (int64_t)p->xh * p->xw * p->cin) * 4, 128};
        const uint32_t box[5] = {32, (uint32_t)a.pw, 4, 1, (uint32_t)((pair ? nt / 2 : nt) / 32)};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        rc = make_tmap_f32(&tx, p->x, 5, dims, strides, box, es, /*atom32=*/true);
        if (rc != SGV_OK) return rc;
    }
    dim3 grid((unsigned)(a.mtiles * (p->cin / nt)), (unsigned)ksplit, (unsigned)a.ngroups);
    switch (nt)
    {
        case 128: return pair ? launch_wgrad_v2<128, 6, true>(tg, tx, a, grid, stream) : launch_wgrad_v2<128, 5>(tg, tx, a, grid, stream);
        case 64:  return launch_wgrad_v2<64, 7>(tg, tx, a, grid, stream);
        default:  return launch_wgrad_v2<32, 8>(tg, tx, a, grid, stream);
    }
}

} // namespace sgv
