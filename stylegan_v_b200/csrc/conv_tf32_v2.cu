// Implicit-GEMM convolution v2 for sm_100a: one haloed activation patch per channel chunk feeds ALL taps.
//
// Same contract as conv_tf32.cu (include/sgv_b200_conv.h, stride-1 inputs).  What changed and why (profiles/: v1 was bound
// by L2->SM operand traffic, 96-192 B/clk/SM wanted vs ~40 available, and by per-tap operand transforms):
//   * A operand: per 32-channel chunk ONE TMA box {32 ch, PW, PH} = the output tile (16 rows x 8*MH cols) plus the halo
//     the taps reach.  Each tap's A matrix is that same patch addressed through a SHIFTED shared-memory descriptor:
//     start = patch + ((dy_t-dy_min)*PW + (dx_t-dx_min) + 8*half)*128 B, SBO = PW*128 B (16 row-groups = 16 image rows of
//     8 pixels).  tcgen05 applies the 128B swizzle to absolute smem address bits, so any 128-byte start row and any SBO
//     are legal (measured: profiles/umma_probe_r1.txt).  A traffic and the style-modulation/TF32-rounding transform drop
//     from 9 x 128 rows to <= 324 rows per chunk.
//   * M = 256 pixels per CTA (MH = 2 halves, two TMEM accumulators) share every weight slab: B traffic per FLOP halves.
//   * separate rings for patches (per chunk) and weight slabs (per chunk x tap).
// Warp roles as in v1: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 transform then epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/sgv_b200_conv.h"

namespace sgv {

using namespace ptx;

constexpr int kV2Threads = 192;
constexpr int kV2TileH = 16;          // output rows per tile; one UMMA row-group (8 pixels) per image row
constexpr int kV2MaxPatch = 18 * 18 * 128;

struct ConvV2Args
{
    float* y; const float* a_scale; const float* o_scale; const float* bias;
    int n, cin, cout, out_h, out_w;
    long long osn, osy, osx;
    int ntaps;
    int tap_row[SGV_CONV_MAX_TAPS];   // (dy_t - dy_min) * PW + (dx_t - dx_min): row offset of the tap inside the patch
    int dy_min, dx_min, pw, ph;
    int tiles_x, tiles_y;
    int act; float alpha, gain, clamp;
};

template <int BN, int MH, int SA, int SB>
struct ConvV2Smem
{
    static constexpr int kPatch = (((16 + 2) * (8 * MH + 2) * 128) + 1023) & ~1023;   // sized for 3x3 windows (largest supported halo)
    static constexpr int kBTile = BN * 128;
    static constexpr int kBOffset = SA * kPatch;
    static constexpr int kBarOffset = kBOffset + SB * kBTile;
    static constexpr int kNumBars = 3 * SA + 2 * SB + 1;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
};

template <int BN, int MH, int SA, int SB>
__global__ void __launch_bounds__(kV2Threads, 1)
conv_tf32_v2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvV2Args p)
{
    using L = ConvV2Smem<BN, MH, SA, SB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* ready_a = full_a + SA;
    uint64_t* empty_a = ready_a + SA;
    uint64_t* full_b = empty_a + SA;
    uint64_t* empty_b = full_b + SB;
    uint64_t* accum_bar = empty_b + SB;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int mt = blockIdx.x;
    const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
    const int tile_y = mt % p.tiles_y; mt /= p.tiles_y;
    const int n = mt;
    const int ox0 = tile_x * 8 * MH, oy0 = tile_y * kV2TileH;
    const int nb0 = blockIdx.y * BN;
    const int kchunks = p.cin / 32;
    const uint32_t patch_bytes = (uint32_t)(p.pw * p.ph * 128);

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < SA; s++) { mbar_init(full_a + s, 1); mbar_init(ready_a + s, 4); mbar_init(empty_a + s, 1); }
        for (int s = 0; s < SB; s++) { mbar_init(full_b + s, 1); mbar_init(empty_b + s, 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    constexpr int kTmemCols = (BN * MH) < 32 ? 32 : (BN * MH);
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        if (elect_one())
        {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            for (int kc = 0; kc < kchunks; kc++)
            {
                mbar_wait(empty_a + sa, pa ^ 1);
                mbar_expect_tx(full_a + sa, patch_bytes);
                tma_load_4d(smem + sa * L::kPatch, &tmap_x, full_a + sa, kc * 32, ox0 + p.dx_min, oy0 + p.dy_min, n);
                if (++sa == SA) { sa = 0; pa ^= 1; }
                for (int t = 0; t < p.ntaps; t++)
                {
                    mbar_wait(empty_b + sb, pb ^ 1);
                    mbar_expect_tx(full_b + sb, L::kBTile);
                    tma_load_2d(smem + L::kBOffset + sb * L::kBTile, &tmap_w, full_b + sb, kc * 32, t * p.cout + nb0);
                    if (++sb == SB) { sb = 0; pb ^= 1; }
                }
            }
        }
    }
    else if (warp == 1)
    {
        constexpr uint32_t idesc = umma_idesc_tf32(128, BN);
        const uint64_t sbo_field = (uint64_t)((uint32_t)(p.pw * 128) >> 4) << 32;
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        for (int kc = 0; kc < kchunks; kc++)
        {
            mbar_wait(ready_a + sa, pa);
            tc_fence_after();
            const uint32_t patch = smem_u32(smem + sa * L::kPatch);
            for (int t = 0; t < p.ntaps; t++)
            {
                mbar_wait(full_b + sb, pb);
                tc_fence_after();
                if (elect_one())
                {
                    const uint64_t db = umma_desc_k_sw128(smem_u32(smem + L::kBOffset + sb * L::kBTile));
#pragma unroll
                    for (int h = 0; h < MH; h++)
                    {
                        // K-major SWIZZLE_128B descriptor with SBO = patch row pitch: row-group g = image row g of this half
                        uint64_t da = umma_desc_k_sw128(patch + (uint32_t)(p.tap_row[t] + 8 * h) * 128u);
                        da = (da & ~((uint64_t)0x3FFF << 32)) | sbo_field;
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            mma_tf32(tmem_base + (uint32_t)(h * BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kc > 0 || t > 0 || k > 0) ? 1u : 0u);
                    }
                    mma_commit(empty_b + sb);
                    if (t == p.ntaps - 1)
                    {
                        mma_commit(empty_a + sa);
                        if (kc == kchunks - 1) mma_commit(accum_bar);
                    }
                }
                __syncwarp();
                if (++sb == SB) { sb = 0; pb ^= 1; }
            }
            if (++sa == SA) { sa = 0; pa ^= 1; }
        }
    }
    else
    {
        const int tid = threadIdx.x - 64;
        const int q = warp & 3;
        const int nrows = p.pw * p.ph;
        {
            int sa = 0; uint32_t pa = 0;
            for (int kc = 0; kc < kchunks; kc++)
            {
                float sv[32];
                if (p.a_scale)
                {
                    const float4* sp = reinterpret_cast<const float4*>(p.a_scale + (long long)n * p.cin + kc * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++) { float4 v = __ldg(sp + j); sv[4 * j] = v.x; sv[4 * j + 1] = v.y; sv[4 * j + 2] = v.z; sv[4 * j + 3] = v.w; }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) sv[j] = 1.f;
                }
                mbar_wait(full_a + sa, pa);
                uint8_t* patch = smem + sa * L::kPatch;
                for (int row = tid; row < nrows; row += 128)
                {
                    uint8_t* arow = patch + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        float4* ptr = reinterpret_cast<float4*>(arow + ((j ^ (row & 7)) << 4));
                        float4 v = *ptr;
                        v.x = tf32_rn(v.x * sv[4 * j + 0]); v.y = tf32_rn(v.y * sv[4 * j + 1]);
                        v.z = tf32_rn(v.z * sv[4 * j + 2]); v.w = tf32_rn(v.w * sv[4 * j + 3]);
                        *ptr = v;
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(ready_a + sa);
                if (++sa == SA) { sa = 0; pa ^= 1; }
            }
        }
        // ---- epilogue ----
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int row = q * 32 + lane;               // accumulator row: image row row/8, column row%8 of the half
        const int oy = oy0 + (row >> 3);
        const float* osc = p.o_scale ? p.o_scale + (long long)n * p.cout + nb0 : nullptr;
        const float* bia = p.bias ? p.bias + nb0 : nullptr;
#pragma unroll 1
        for (int h = 0; h < MH; h++)
        {
            const int ox = ox0 + 8 * h + (row & 7);
            const bool valid = (oy < p.out_h) && (ox < p.out_w);
            float* yrow = p.y + (long long)n * p.osn + (long long)oy * p.osy + (long long)ox * p.osx + nb0;
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; cc++)
            {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * BN + cc * 32), v);
                tmem_ld_wait();
                if (valid)
                {
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            const int col = cc * 32 + j * 4 + e;
                            float f = __uint_as_float(v[j * 4 + e]);
                            if (osc) f = __fmul_rn(f, __ldg(osc + col));
                            if (bia) f = __fadd_rn(f, __ldg(bia + col));
                            if (p.act == 3) f = (f > 0.f) ? f : f * p.alpha;
                            f *= p.gain;
                            if (p.clamp >= 0.f) f = (f > -p.clamp && f < p.clamp) ? f : (f >= 0.f ? p.clamp : -p.clamp);
                            o[e] = f;
                        }
                        *reinterpret_cast<float4*>(yrow + cc * 32 + j * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

template <int BN, int MH, int SA, int SB>
static int launch_v2(const CUtensorMap& tx, const CUtensorMap& tw, const ConvV2Args& a, dim3 grid, cudaStream_t stream)
{
    using L = ConvV2Smem<BN, MH, SA, SB>;
    auto kern = conv_tf32_v2_kernel<BN, MH, SA, SB>;
    static bool attr_set = false;
    if (!attr_set)
    {
        SGV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        attr_set = true;
    }
    kern<<<grid, kV2Threads, L::kTotal, stream>>>(tx, tw, a);
    SGV_LAUNCH_OK("conv_tf32_v2_kernel");
    return SGV_OK;
}

// Returns SGV_ERR_UNSUPPORTED (without touching the error text) when the shape is outside v2's envelope; the caller
// then uses the v1 kernel.
int conv2d_tf32_v2(const sgv_conv_params* p, cudaStream_t stream)
{
    if (p->in_stride != 1 || p->out_h < 12 || p->out_w < 12 || p->cout % 64 != 0) return SGV_ERR_UNSUPPORTED;
    if (p->in_stride_x != 0 || p->accumulate || p->red_out) return SGV_ERR_UNSUPPORTED;
    int dy_min = p->tap_dy[0], dy_max = p->tap_dy[0], dx_min = p->tap_dx[0], dx_max = p->tap_dx[0];
    for (int t = 1; t < p->ntaps; t++)
    {
        dy_min = min(dy_min, p->tap_dy[t]); dy_max = max(dy_max, p->tap_dy[t]);
        dx_min = min(dx_min, p->tap_dx[t]); dx_max = max(dx_max, p->tap_dx[t]);
    }
    if (dy_max - dy_min > 2 || dx_max - dx_min > 2) return SGV_ERR_UNSUPPORTED;
    const int mh = 2;
    ConvV2Args a;
    a.y = p->y; a.a_scale = p->a_scale; a.o_scale = p->o_scale; a.bias = p->bias;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    a.osn = p->out_stride_n; a.osy = p->out_stride_y; a.osx = p->out_stride_x;
    a.ntaps = p->ntaps;
    a.dy_min = dy_min; a.dx_min = dx_min;
    a.pw = 8 * mh + (dx_max - dx_min); a.ph = kV2TileH + (dy_max - dy_min);
    for (int t = 0; t < SGV_CONV_MAX_TAPS; t++) a.tap_row[t] = t < p->ntaps ? (p->tap_dy[t] - dy_min) * a.pw + (p->tap_dx[t] - dx_min) : 0;
    a.tiles_x = ceil_div(p->out_w, 8 * mh); a.tiles_y = ceil_div(p->out_h, kV2TileH);
    a.act = p->act; a.alpha = p->alpha; a.gain = p->gain; a.clamp = p->clamp;
    const int bn = (p->cout % 256 == 0) ? 256 : (p->cout % 128 == 0) ? 128 : 64;

    CUtensorMap tmx, tmw;
    {
        const uint64_t dims[4] = {(uint64_t)p->cin, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
        const uint64_t strides[3] = {(uint64_t)p->cin * 4, (uint64_t)p->w * p->cin * 4, (uint64_t)p->h * p->w * p->cin * 4};
        const uint32_t box[4] = {32, (uint32_t)a.pw, (uint32_t)a.ph, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        int rc = make_tmap_f32(&tmx, p->x, 4, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p->cin, (uint64_t)p->ntaps * p->cout};
        const uint64_t strides[1] = {(uint64_t)p->cin * 4};
        const uint32_t box[2] = {32, (uint32_t)bn};
        const uint32_t es[2] = {1, 1};
        int rc = make_tmap_f32(&tmw, p->wp, 2, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * p->n), (unsigned)(p->cout / bn), 1);
    switch (bn)
    {
        case 256: return launch_v2<256, 2, 2, 3>(tmx, tmw, a, grid, stream);   // 84 KB patches + 96 KB slabs
        case 128: return launch_v2<128, 2, 2, 6>(tmx, tmw, a, grid, stream);
        default:  return launch_v2<64, 2, 2, 3>(tmx, tmw, a, grid, stream);    // 108 KB: two CTAs per SM
    }
}

} // namespace sgv
