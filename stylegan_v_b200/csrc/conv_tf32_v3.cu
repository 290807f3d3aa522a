// Implicit-GEMM convolution for sm_100a, persistent halo-patch kernel ("v3"; haloed patch + shifted descriptors + 2-4 M-halves).
//
// Its round-1 predecessor launched one CTA per 256-pixel tile.  For the high-resolution, few-channel layers a tile is only ~2.4 us
// of MMA work, so the serial chain  TMA latency -> transform -> MMA -> TMEM read-out -> stores  (plus barrier init, TMEM alloc and
// descriptor fetch per CTA) left the tensor pipe 24 % active (ncu: profiles/ncu_tc_r1l_summary.txt).  This kernel keeps one CTA per
// SM alive and walks tiles round-robin (tile = blockIdx.x + i * gridDim.x):
//   * the TMA producer, the transform warps and the MMA issuer run ahead ACROSS tile boundaries through the same rings;
//   * the accumulator is double-buffered in TMEM (when 2 * MH * BN <= 512 columns);
//   * four dedicated epilogue warps drain tile i (TMEM -> regs -> dcoefs/bias/lrelu/gain -> NHWC stores) while tile i+1
//     is being multiplied.
// Operand staging: per 32-channel chunk ONE TMA box = output tile (16 rows x 16 cols) + halo; every tap's A matrix is
// that patch through a shifted K-major SWIZZLE_128B descriptor (start row = (dy-dy_min)*PW + (dx-dx_min) + 8*half,
// SBO = PW*128; legal because tcgen05 swizzles on absolute smem address bits — profiles/umma_probe_r1.txt).
// Warps: 0 activation-patch TMA producer, 1 MMA issuer + TMEM owner, 2-5 transform (styles * x, round to TF32), 6-13 epilogue
// (two groups of four warps = one per TMEM lane quarter; the groups take alternate 32-column units of a tile), 14 weight-slab TMA producer.  The two producers are separate threads on purpose: with one thread issuing "patch, then its
// ntaps slabs" the next patch could not be requested until the MMAs had freed slab slots, and the transform warps spent 2/3 of
// their time waiting for patches (ncu source view, profiles/ncu_v3_r1y_summary.txt).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/sgv_b200_conv.h"

namespace sgv {

using namespace ptx;

constexpr int kV3Threads = 64 + 128 + 256 + 32;
constexpr int kV3TileH = 16;
constexpr int kV3MaxCls = 8;                              // 4 parity classes (stride 2) x {hi, lo} operand part (tf32x3)
constexpr int kV3MaxTaps = 3 * SGV_CONV_MAX_TAPS;         // tf32x3: every tap is issued three times (hi*hi, hi*lo, lo*hi)

struct ConvV3Args
{
    float* y; const float* a_scale; const float* o_scale; const float* bias;
    int n, cin, cout, out_h, out_w;
    long long osn, osy, osx;
    int ntaps;      // class-ordered MMA taps (tf32x3: 3 per tap of the contraction)
    // Taps are grouped into patch CLASSES: all taps of a class read the same TMA box (origin cls_ox/oy relative to the tile's first
    // input pixel, sampled with the input stride) through shifted descriptors.  Stride 1: one class.  Stride 2 (data gradient of the
    // stride-2 transposed conv): one class per (dy, dx) parity, i.e. 4 boxes of every-other pixel per 32-channel chunk.
    // tf32x3 (fp32-grade) mode doubles the classes: class (c, hi) stages tf32(x*s) and is multiplied with the hi AND the lo weight slabs,
    // class (c, lo) stages tf32(x*s - tf32(x*s)) from a second TMA load of the same box and is multiplied with the hi slabs only.
    int in_stride, ncls;
    int cls_ntaps[kV3MaxCls], cls_ox[kV3MaxCls], cls_oy[kV3MaxCls], cls_lo[kV3MaxCls];
    int tap_row[kV3MaxTaps];               // class-ordered: start row of the tap's A matrix inside its class patch
    int tap_slab[kV3MaxTaps];              // class-ordered: first row of the tap's weight slab in the [rows, cin] slab tensor (hi set, then lo set)
    int pw, ph;
    int tiles_x, tiles_y, ntiles_n, total_groups;      // group = CL pixel tiles x one n-tile, one tile per CTA of a cluster
    int act; float alpha, gain, clamp;
    int accumulate;
    const float* red_x; float* red_out;
    int a_ready;    // activation patches need no staging pass
    const float* noise; long long nsn, nsy, nsx;      // per-pixel noise plane(s), added after o_scale (element strides; nsn = 0: shared plane)
    int debug;      // ablation switches, only honoured by -DSGV_ABLATION builds (profiles/conv_v3_ablation_r1.txt): 1 skip transform, 2 skip epilogue, 4 skip MMAs
};

template <int BN, int MH, int SA, int SB, bool PAIR = false>
struct ConvV3Smem
{
    static constexpr int kPatch = (((16 + 2) * (8 * MH + 2) * 128) + 1023) & ~1023;
    static constexpr int kBTile = (PAIR ? BN / 2 : BN) * 128;      // CTA-pair MMAs: each CTA of the pair holds half of the slab's rows
    static constexpr int kBOffset = SA * kPatch;
    static constexpr int kStageOffset = kBOffset + SB * kBTile;      // 2 x 16 KB output staging (128 pixels x 32 channels, 128B-swizzled)
    static constexpr int kBarOffset = kStageOffset + 2 * 16384;
    static constexpr int kNumBars = 3 * SA + 2 * SB + 4;
    static constexpr int kTotal = kBarOffset + kNumBars * 8 + 16 + 1024;
    static constexpr int kAccBufs = (2 * MH * BN <= 512) ? 2 : 1;
    static constexpr int kTmemCols = (kAccBufs * MH * BN) <= 32 ? 32 : (kAccBufs * MH * BN) <= 64 ? 64 : (kAccBufs * MH * BN) <= 128 ? 128
                                     : (kAccBufs * MH * BN) <= 256 ? 256 : 512;
};

struct TileCoord { int n, ox0, oy0, nb0; };

template <int BN, int MH, int CL>
__device__ __forceinline__ TileCoord tile_coord(const ConvV3Args& p, int group, int crank)
{
    // n-tile fastest: the CTAs working on the same pixels at the same time share the activation patch through L2;
    // the CL CTAs of a cluster take CL consecutive pixel tiles of the SAME n-tile (they share every weight slab)
    TileCoord c;
    const int nt = group % p.ntiles_n;
    int tile = (group / p.ntiles_n) * CL + crank;
    const int tx = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty = tile % p.tiles_y; tile /= p.tiles_y;
    c.n = tile; c.ox0 = tx * 8 * MH; c.oy0 = ty * kV3TileH; c.nb0 = nt * BN;
    return c;
}

// PAIR (needs CL == 2): the two CTAs of a cluster issue their MMAs as ONE tcgen05.mma.cta_group::2 of M = 256 — CTA r supplies the 128 pixel
// rows of ITS tile (own patches, own staging warps, own accumulator rows in its own TMEM, own epilogue) and rows [r*BN/2, (r+1)*BN/2) of every
// weight slab, so a CTA reads only half of B from shared memory per instruction (the B-operand traffic that makes the 128 B/clk
// shared-memory port the limiter in the one-CTA form: DESIGN.md §4).  Only the even CTA (the leader) issues MMAs; synchronisation:
//   ready_a  (leader's copy, 8 arrivals)   the 4 staging warps of BOTH CTAs arrive here (remote arrive from the odd CTA)
//   full_b   (leader's copy)               both CTAs' slab TMAs complete their bytes on it (cta_group::2 TMA form); the leader expects both halves
//   empty_a / empty_b / acc_full           tcgen05.commit.cta_group::2 multicast: one arrival on the barrier in each CTA
//   acc_empty (leader's copy, 16 arrivals) the 8 epilogue warps of both CTAs
template <int BN, int MH, int SA, int SB, int CL, bool PAIR>
__global__ void __launch_bounds__(kV3Threads, 1)
conv_tf32_v3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const __grid_constant__ CUtensorMap tmap_y, const ConvV3Args p)
{
    static_assert(!PAIR || CL == 2, "CTA-pair MMAs need clusters of exactly two CTAs");
    using L = ConvV3Smem<BN, MH, SA, SB, PAIR>;
    constexpr int NB = L::kAccBufs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_a = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* ready_a = full_a + SA;
    uint64_t* empty_a = ready_a + SA;
    uint64_t* full_b = empty_a + SA;
    uint64_t* empty_b = full_b + SB;
    uint64_t* acc_full = empty_b + SB;       // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / CL, ncl = gridDim.x / CL;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    const int kchunks = p.cin / 32;
    const uint32_t patch_bytes = (uint32_t)(p.pw * p.ph * 128);

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        prefetch_tmap(&tmap_y);
        for (int s = 0; s < SA; s++) { mbar_init(full_a + s, 1); mbar_init(ready_a + s, PAIR ? 8 : 4); mbar_init(empty_a + s, 1); }
        for (int s = 0; s < SB; s++) { mbar_init(full_b + s, 1); mbar_init(empty_b + s, PAIR ? 1 : CL); }   // a slab slot is freed by the MMAs of all CL CTAs (pair: by the leader's, multicast)
        for (int s = 0; s < 2; s++) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, PAIR ? 16 : 8); }
        fence_mbar_init();
    }
    if (warp == 1)
    {
        if (PAIR) { tmem_alloc_pair(tmem_slot, L::kTmemCols); tmem_relinquish_pair(); }
        else { tmem_alloc(tmem_slot, L::kTmemCols); tmem_relinquish(); }
    }
    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();      // peers' barriers are initialised before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        // ===== activation-patch producer =====
        if (elect_one())
        {
            int sa = 0; uint32_t pa = 0;
            for (int g = cid; g < p.total_groups; g += ncl)
            {
                const TileCoord tc = tile_coord<BN, MH, CL>(p, g, crank);
                for (int kc = 0; kc < kchunks; kc++)
                    for (int c = 0; c < p.ncls; c++)
                    {
                        mbar_wait(empty_a + sa, pa ^ 1);
                        mbar_expect_tx(full_a + sa, patch_bytes);
                        tma_load_4d(smem + sa * L::kPatch, &tmap_x, full_a + sa, kc * 32, tc.ox0 * p.in_stride + p.cls_ox[c],
                                    tc.oy0 * p.in_stride + p.cls_oy[c], tc.n);
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
            }
        }
    }
    else if (warp == 14)
    {
        // ===== weight-slab producer =====
        if (elect_one())
        {
            int sb = 0; uint32_t pb = 0;
            for (int g = cid; g < p.total_groups; g += ncl)
            {
                const int nb0 = (g % p.ntiles_n) * BN;
                for (int kc = 0; kc < kchunks; kc++)
                {
                    for (int t = 0; t < p.ntaps; t++)
                    {
                        mbar_wait(empty_b + sb, pb ^ 1);
                        if (PAIR)
                        {
                            // this CTA's half of the slab rows into ITS shared memory; both halves complete on the leader's barrier
                            if (crank == 0) mbar_expect_tx(full_b + sb, 2 * L::kBTile);
                            tma_load_2d_pair(smem + L::kBOffset + sb * L::kBTile, &tmap_w, full_b + sb, kc * 32, p.tap_slab[t] + nb0 + crank * (BN / 2));
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(full_b + sb, L::kBTile);
                        if (CL == 1)
                            tma_load_2d(smem + L::kBOffset + sb * L::kBTile, &tmap_w, full_b + sb, kc * 32, p.tap_slab[t] + nb0);
                        else    // this CTA fetches rows [crank * BN/CL, +BN/CL) of the slab once and multicasts them to the whole cluster
                            tma_load_2d_mc(smem + L::kBOffset + sb * L::kBTile + crank * (BN / CL) * 128, &tmap_w, full_b + sb, kc * 32,
                                           p.tap_slab[t] + nb0 + crank * (BN / CL), kMask);
                        if (++sb == SB) { sb = 0; pb ^= 1; }
                    }
                }
            }
        }
    }
    else if (warp == 1)
    {
        // ===== MMA issuer: ONE elected thread runs the whole role (no per-tap election / reconvergence), descriptors are built
        //       from a constant high word and a 32-bit low word that only needs integer adds per tap / half / k-step =====
        if (elect_one() && (!PAIR || crank == 0))
        {
            constexpr uint32_t idesc = umma_idesc_tf32(PAIR ? 256 : 128, BN);
            const uint64_t proto = umma_desc_k_sw128(0);
            const uint32_t b_hi = (uint32_t)(proto >> 32);
            const uint32_t a_hi = (b_hi & ~0x3FFFu) | (((uint32_t)(p.pw * 128) >> 4) & 0x3FFFu);      // SBO = patch row pitch
            const uint32_t lo_flags = (uint32_t)proto;                                                   // LBO field (start address bits are 0)
            const uint32_t smem_lo = (smem_u32(smem) & 0x3FFFFu) >> 4;
            auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int it = 0;
            for (int g = cid; g < p.total_groups; g += ncl, it++)
            {
                const int buf = it % NB;
                const uint32_t use = (uint32_t)(it / NB) & 1u;
                mbar_wait(acc_empty + buf, use ^ 1);                     // epilogue has drained this accumulator buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(buf * MH * BN);
                for (int kc = 0; kc < kchunks; kc++)
                {
                    int t = 0;
                    for (int c = 0; c < p.ncls; c++)
                    {
                        mbar_wait(ready_a + sa, pa);
                        tc_fence_after();
                        const uint32_t patch_lo = (smem_lo + (uint32_t)(sa * (L::kPatch >> 4))) | lo_flags;
                        const int t_end = t + p.cls_ntaps[c];
                        for (; t < t_end; t++)
                        {
                            const uint32_t a_lo = patch_lo + (uint32_t)p.tap_row[t] * 8u;                     // 128-byte rows -> 16-byte units
                            const uint32_t b_lo = (smem_lo + (uint32_t)((L::kBOffset + sb * L::kBTile) >> 4)) | lo_flags;
                            mbar_wait(full_b + sb, pb);
                            tc_fence_after();
#pragma unroll
                            for (int h = 0; h < MH; h++)
#pragma unroll
                                for (int k = 0; k < 4; k++)
                                    if (!SGV_ABL(p.debug, 4))
                                    {
                                        if (PAIR) mma_tf32_pair(acc + (uint32_t)(h * BN), desc(a_lo + (uint32_t)(h * 64 + k * 2), a_hi), desc(b_lo + (uint32_t)(k * 2), b_hi), idesc,
                                                                (kc > 0 || t > 0 || k > 0) ? 1u : 0u);
                                        else mma_tf32(acc + (uint32_t)(h * BN), desc(a_lo + (uint32_t)(h * 64 + k * 2), a_hi), desc(b_lo + (uint32_t)(k * 2), b_hi), idesc,
                                                      (kc > 0 || t > 0 || k > 0) ? 1u : 0u);
                                    }
                            if (PAIR) mma_commit_pair_mc(empty_b + sb, kMask);
                            else if (CL == 1) mma_commit(empty_b + sb);
                            else mma_commit_mc(empty_b + sb, kMask);
                            if (t == t_end - 1)
                            {
                                if (PAIR) mma_commit_pair_mc(empty_a + sa, kMask); else mma_commit(empty_a + sa);
                                if (kc == kchunks - 1 && c == p.ncls - 1) { if (PAIR) mma_commit_pair_mc(acc_full + buf, kMask); else mma_commit(acc_full + buf); }
                            }
                            if (++sb == SB) { sb = 0; pb ^= 1; }
                        }
                        if (++sa == SA) { sa = 0; pa ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    }
    else if (warp < 6)
    {
        // ===== operand transform (4 warps) =====
        const int tid = threadIdx.x - 64;
        const int nrows = p.pw * p.ph;
        int sa = 0; uint32_t pa = 0;
        for (int g = cid; g < p.total_groups; g += ncl)
        {
            const TileCoord tc = tile_coord<BN, MH, CL>(p, g, crank);
            for (int kc = 0; kc < kchunks; kc++)
            {
                float sv[32];
                if (p.a_scale)
                {
                    const float4* sp = reinterpret_cast<const float4*>(p.a_scale + (long long)tc.n * p.cin + kc * 32);
#pragma unroll
                    for (int j = 0; j < 8; j++) { float4 v = __ldg(sp + j); sv[4 * j] = v.x; sv[4 * j + 1] = v.y; sv[4 * j + 2] = v.z; sv[4 * j + 3] = v.w; }
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) sv[j] = 1.f;
                }
                for (int c = 0; c < p.ncls; c++)
                {
                    mbar_wait(full_a + sa, pa);
                    const uint32_t patch = smem_u32(smem + sa * L::kPatch);
                    const bool lo_part = p.cls_lo[c] != 0;
                    for (int row = tid; row < ((SGV_ABL(p.debug, 1) || p.a_ready) ? 0 : nrows); row += 128)
                    {
                        const uint32_t arow = patch + (uint32_t)row * 128u;
                        float4 v[8];
#pragma unroll
                        for (int j = 0; j < 8; j++) v[j] = lds128(arow + (uint32_t)((j ^ (row & 7)) << 4));
                        if (!lo_part)
                        {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                            {
                                v[j].x = tf32_rn(v[j].x * sv[4 * j + 0]); v[j].y = tf32_rn(v[j].y * sv[4 * j + 1]);
                                v[j].z = tf32_rn(v[j].z * sv[4 * j + 2]); v[j].w = tf32_rn(v[j].w * sv[4 * j + 3]);
                                sts128(arow + (uint32_t)((j ^ (row & 7)) << 4), v[j]);
                            }
                        }
                        else
                        {
#pragma unroll
                            for (int j = 0; j < 8; j++)
                            {
                                v[j].x = tf32_lo(__fmul_rn(v[j].x, sv[4 * j + 0])); v[j].y = tf32_lo(__fmul_rn(v[j].y, sv[4 * j + 1]));
                                v[j].z = tf32_lo(__fmul_rn(v[j].z, sv[4 * j + 2])); v[j].w = tf32_lo(__fmul_rn(v[j].w, sv[4 * j + 3]));
                                sts128(arow + (uint32_t)((j ^ (row & 7)) << 4), v[j]);
                            }
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_leader(ready_a + sa); else mbar_arrive(ready_a + sa); }
                    if (++sa == SA) { sa = 0; pa ^= 1; }
                }
            }
        }
    }
    else
    {
        // ===== epilogue (8 warps in two groups): TMEM -> registers -> dcoefs/bias/lrelu/gain -> swizzled staging -> TMA store =====
        // Thread `row` owns one output pixel of the 16 x 8 half tile and 32 consecutive channels per unit.  Writing those straight to
        // global memory made every STG.128 touch 32 different lines (16 B partial writes): the drain took 2-4x the MMA time of a tile
        // (ablation: profiles/conv_v3_ablation_r1.txt).  Instead the unit is staged as 128 rows x 128 B in the TMA 128B-swizzle
        // (conflict-free STS.128) and written by ONE bulk tensor store, which also clips the ragged right / bottom edge.
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int grp = (warp - 6) >> 2;                          // epilogue group 0 / 1: units with cc % 2 == grp, own staging buffer + named barrier
        const bool issuer = (threadIdx.x == 192 + grp * 128);
        const uint32_t stage = smem_u32(smem + L::kStageOffset) + (uint32_t)grp * 16384u;
        int it = 0;
        for (int g = cid; g < p.total_groups; g += ncl, it++)
        {
            const TileCoord tc = tile_coord<BN, MH, CL>(p, g, crank);
            const int buf = it % NB;
            const uint32_t use = (uint32_t)(it / NB) & 1u;
            mbar_wait(acc_full + buf, use);
            tc_fence_after();
            const uint32_t acc = tmem_base + (uint32_t)(buf * MH * BN);
            const int oy = tc.oy0 + (row >> 3);
#pragma unroll 1
            for (int h = 0; h < (SGV_ABL(p.debug, 2) ? 0 : MH); h++)
            {
                const int ox = tc.ox0 + 8 * h + (row & 7);
                const bool valid = (oy < p.out_h) && (ox < p.out_w);
                const float nz = (p.noise && valid) ? __ldg(p.noise + (long long)tc.n * p.nsn + (long long)oy * p.nsy + (long long)ox * p.nsx) : 0.f;
#pragma unroll 1
                for (int cc = grp; cc < BN / 32; cc += 2)
                {
                    uint32_t v[32];
                    tmem_ld_32x32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * BN + cc * 32), v);
                    tmem_ld_wait();
                    if (h == MH - 1 && cc == BN / 32 - 2 + grp)
                    {
                        // the accumulator buffer is in registers now: hand it back to the MMA issuer before the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) mbar_arrive_leader(acc_empty + buf); else mbar_arrive(acc_empty + buf); }
                    }
                    if (p.red_out)
                    {
                        // fused style-gradient reduction: sum over this warp's 32 pixels of raw * red_x, one channel per lane
                        float prod[32];
                        const float* rx = p.red_x + (long long)tc.n * p.osn + (long long)oy * p.osy + (long long)ox * p.osx + tc.nb0 + cc * 32;
#pragma unroll
                        for (int j = 0; j < 8; j++)
                        {
                            const float4 xv = valid ? __ldg(reinterpret_cast<const float4*>(rx) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            prod[4 * j + 0] = __uint_as_float(v[4 * j + 0]) * xv.x; prod[4 * j + 1] = __uint_as_float(v[4 * j + 1]) * xv.y;
                            prod[4 * j + 2] = __uint_as_float(v[4 * j + 2]) * xv.z; prod[4 * j + 3] = __uint_as_float(v[4 * j + 3]) * xv.w;
                        }
                        const float tot = warp_reduce_32x32(prod, lane);
                        atomicAdd(p.red_out + (long long)tc.n * p.cout + tc.nb0 + cc * 32 + lane, tot);
                    }
                    const float4* osc = p.o_scale ? reinterpret_cast<const float4*>(p.o_scale + (long long)tc.n * p.cout + tc.nb0 + cc * 32) : nullptr;
                    const float4* bia = p.bias ? reinterpret_cast<const float4*>(p.bias + tc.nb0 + cc * 32) : nullptr;
                    if (issuer) bulk_wait_read<0>();                 // this group's previous store has finished reading the staging buffer
                    named_bar_sync(1 + grp, 128);
#pragma unroll
                    for (int j = 0; j < 8; j++)
                    {
                        float o[4] = {__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])};
                        if (osc) { const float4 sv = __ldg(osc + j); o[0] = __fmul_rn(o[0], sv.x); o[1] = __fmul_rn(o[1], sv.y); o[2] = __fmul_rn(o[2], sv.z); o[3] = __fmul_rn(o[3], sv.w); }
                        if (p.noise) { o[0] = __fadd_rn(o[0], nz); o[1] = __fadd_rn(o[1], nz); o[2] = __fadd_rn(o[2], nz); o[3] = __fadd_rn(o[3], nz); }
                        if (bia) { const float4 bv = __ldg(bia + j); o[0] = __fadd_rn(o[0], bv.x); o[1] = __fadd_rn(o[1], bv.y); o[2] = __fadd_rn(o[2], bv.z); o[3] = __fadd_rn(o[3], bv.w); }
#pragma unroll
                        for (int e = 0; e < 4; e++)
                        {
                            float f = o[e];
                            if (p.act == 3) f = (f > 0.f) ? f : f * p.alpha;
                            f *= p.gain;
                            if (p.clamp >= 0.f) f = (f > -p.clamp && f < p.clamp) ? f : (f >= 0.f ? p.clamp : -p.clamp);
                            o[e] = f;
                        }
                        sts128(stage + (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4), make_float4(o[0], o[1], o[2], o[3]));
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(1 + grp, 128);
                    if (issuer)
                    {
                        if (p.accumulate) tma_reduce_add_4d(&tmap_y, stage, tc.nb0 + cc * 32, tc.ox0 + 8 * h, tc.oy0, tc.n);
                        else tma_store_4d(&tmap_y, stage, tc.nb0 + cc * 32, tc.ox0 + 8 * h, tc.oy0, tc.n);
                        bulk_commit();
                    }
                }
            }
            if (SGV_ABL(p.debug, 2)) { tc_fence_before(); __syncwarp(); if (lane == 0) { if (PAIR) mbar_arrive_leader(acc_empty + buf); else mbar_arrive(acc_empty + buf); } }
        }
        if (issuer) bulk_wait_all();
    }

    tc_fence_before();
    if (CL > 1) cluster_sync_all(); else __syncthreads();      // no CTA leaves while a peer may still multicast into it
    if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, L::kTmemCols); else tmem_dealloc(tmem_base, L::kTmemCols); }
}

template <int BN, int MH, int SA, int SB, int CL, bool PAIR = false>
static int launch_v3(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const ConvV3Args& a, cudaStream_t stream)
{
    using L = ConvV3Smem<BN, MH, SA, SB, PAIR>;
    auto kern = conv_tf32_v3_kernel<BN, MH, SA, SB, CL, PAIR>;
    static PerDeviceInt max_clusters;      // co-resident clusters of this variant, per device
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(kV3Threads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int dev = current_device_slot();
    int nc = max_clusters.get(dev);
    if (nc == 0)
    {
        SGV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        nc = num_sms() / CL;
        if (CL > 1)
        {
            cfg.gridDim = dim3((unsigned)(nc * CL));
            SGV_CUDA_OK(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));     // GPC granularity can leave fewer than num_sms / CL co-resident
        }
        SGV_CHECK_ARG(nc >= 1, "conv_tf32_v3: no co-resident cluster of this size");
        max_clusters.set(dev, nc);
    }
    const int clusters = a.total_groups < nc ? a.total_groups : nc;
    cfg.gridDim = dim3((unsigned)(clusters * CL));
    ConvV3Args args = a;
    SGV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, args));
    SGV_LAUNCH_OK("conv_tf32_v3_kernel");
    return SGV_OK;
}

// CTA-pair variants (cta_group::2 MMAs): same tiles; the halved slab stages buy deeper slab rings
static int launch_v3_pair(int bn, int mh, const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const ConvV3Args& a, cudaStream_t stream)
{
    switch (bn)
    {
        case 256: return launch_v3<256, 2, 2, 4, 2, true>(tx, tw, ty, a, stream);     // 84 KB patches + 4 x 16 KB slab halves
        case 128: return launch_v3<128, 2, 3, 6, 2, true>(tx, tw, ty, a, stream);     // 126 + 6 x 8 KB
        default:
            if (mh == 4) return launch_v3<64, 4, 2, 6, 2, true>(tx, tw, ty, a, stream);   // 154 + 6 x 4 KB
            return launch_v3<64, 2, 3, 8, 2, true>(tx, tw, ty, a, stream);
    }
}

template <int CL>
static int launch_v3_bn(int bn, int mh, const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const ConvV3Args& a, cudaStream_t stream)
{
    switch (bn)
    {
        case 256: return launch_v3<256, 2, 2, 3, CL>(tx, tw, ty, a, stream);     // 84 KB patches + 96 KB slabs, single accumulator buffer (512 cols)
        case 128: return launch_v3<128, 2, 3, 4, CL>(tx, tw, ty, a, stream);     // 126 + 64 KB, double-buffered accumulators (512 cols)
        default:
            if (mh == 4) return launch_v3<64, 4, 2, 4, CL>(tx, tw, ty, a, stream);   // 154 KB patches + 32 KB slabs, double-buffered accumulators (512 cols)
            return launch_v3<64, 2, 3, 8, CL>(tx, tw, ty, a, stream);                // 126 + 64 KB, double-buffered accumulators (256 cols)
    }
}

// Tuning switches, read from the environment ONCE per process (thread-safe static initialisation), never per launch.
struct V3Tuning
{
    int cluster;      // SGV_CONV_CLUSTER=1|2|4: CTAs sharing each weight slab through TMA multicast
    int mh4;          // SGV_V3_MH4=0: no 16x32-pixel tiles for 64-channel layers
    int max_bn;       // SGV_V3_MAXBN
    int wide_cin;     // SGV_V3_WIDE_CIN: smallest cin that takes the 256-column N tile
    int debug;        // SGV_V3_DEBUG: ablation switches (honoured by -DSGV_ABLATION builds only)
    int pair;         // SGV_CONV_PAIR=0: clusters of 2 share slabs by TMA multicast instead of issuing tcgen05 cta_group::2 MMAs (M = 256, B halved per CTA)
    V3Tuning()
    {
        // default ON: measured on the B200 (profiles/conv_pair_probe_r2_*.jsonl) bit-identical results and 0.215 -> 0.212 / 0.197 -> 0.178 /
        // 0.213 -> 0.191 / 0.297 -> 0.267 ms for the 512 / 256 / 128 / 64-channel conv1 layers of the 256^2 network at 32 frames
        pair = env_int("SGV_CONV_PAIR", 1) ? 1 : 0;
        cluster = env_int("SGV_CONV_CLUSTER", 2);
        if (cluster != 1 && cluster != 2 && cluster != 4) cluster = 2;
        mh4 = env_int("SGV_V3_MH4", 1) ? 1 : 0;
        max_bn = env_int("SGV_V3_MAXBN", 256);
        wide_cin = env_int("SGV_V3_WIDE_CIN", 512);
        debug = env_int("SGV_V3_DEBUG", 0);
    }
};
static const V3Tuning& v3_tuning() { static const V3Tuning t; return t; }

// Returns SGV_ERR_UNSUPPORTED when the shape is outside the envelope (caller falls back to the per-tap kernel).  With `query` the chosen variant is
// reported and nothing is launched.
int conv2d_tf32_v3(const sgv_conv_params* p, cudaStream_t stream, sgv_conv_variant* query)
{
    if ((p->in_stride != 1 && p->in_stride != 2) || p->out_h < 12 || p->out_w < 12 || p->cout % 64 != 0) return SGV_ERR_UNSUPPORTED;
    if (p->in_stride == 2 && p->in_stride_x != 0) return SGV_ERR_UNSUPPORTED;      // strided sampling of a strided view: not needed by any caller
    const V3Tuning& tune = v3_tuning();
    const int st = p->in_stride;
    const bool x3 = p->wp_lo != nullptr;
    // 64 output channels: each weight slab feeds only 128 x 64 MMAs, so slabs re-streamed per tile saturate the ~40 B/clk L2 -> SM path
    // (profiles/conv_v3_ablation_r1.txt); tiles of 4 halves (16 x 32 pixels) halve that traffic and still fit two accumulator buffers.
    const int mh = (tune.mh4 && p->cout % 128 != 0 && p->out_w >= 32) ? 4 : 2;
    ConvV3Args a;
    memset(&a, 0, sizeof(a));
    // group the taps into patch classes by the parity of (dy, dx) modulo the input stride
    int cls_of[SGV_CONV_MAX_TAPS], cmin_x[4], cmin_y[4], cmax_x[4], cmax_y[4], ckey[4];
    int ncls = 0;
    for (int t = 0; t < p->ntaps; t++)
    {
        const int key = (((p->tap_dy[t] % st) + st) % st) * st + (((p->tap_dx[t] % st) + st) % st);
        int c = -1;
        for (int j = 0; j < ncls; j++) if (ckey[j] == key) c = j;
        if (c < 0)
        {
            c = ncls++;
            ckey[c] = key; cmin_x[c] = cmax_x[c] = p->tap_dx[t]; cmin_y[c] = cmax_y[c] = p->tap_dy[t];
        }
        cmin_x[c] = min(cmin_x[c], p->tap_dx[t]); cmax_x[c] = max(cmax_x[c], p->tap_dx[t]);
        cmin_y[c] = min(cmin_y[c], p->tap_dy[t]); cmax_y[c] = max(cmax_y[c], p->tap_dy[t]);
        cls_of[t] = c;
    }
    int ext_x = 0, ext_y = 0;
    for (int c = 0; c < ncls; c++) { ext_x = max(ext_x, (cmax_x[c] - cmin_x[c]) / st); ext_y = max(ext_y, (cmax_y[c] - cmin_y[c]) / st); }
    if (ext_x > 2 || ext_y > 2) return SGV_ERR_UNSUPPORTED;
    a.y = p->y; a.a_scale = p->a_scale; a.o_scale = p->o_scale; a.bias = p->bias;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    a.osn = p->out_stride_n; a.osy = p->out_stride_y; a.osx = p->out_stride_x;
    a.in_stride = st;
    a.pw = 8 * mh + ext_x; a.ph = kV3TileH + ext_y;
    // rows of the slab tensor [rows, cin] the weight tensor map covers: the hi set, then (tf32x3) the lo set wherever the caller put it
    const long long lo_row0 = x3 ? (long long)(p->wp_lo - p->wp) / p->cin : 0;
    const long long slab_rows = (x3 ? lo_row0 : 0) + (long long)p->ntaps * p->cout;
    {
        int j = 0;
        a.ncls = 0;
        for (int c = 0; c < ncls; c++)
            for (int part = 0; part < (x3 ? 2 : 1); part++)      // part 0 = hi operand, part 1 = lo operand (tf32x3 only)
            {
                const int cc = a.ncls++;
                a.cls_ox[cc] = cmin_x[c]; a.cls_oy[cc] = cmin_y[c]; a.cls_ntaps[cc] = 0; a.cls_lo[cc] = part;
                for (int wpart = 0; wpart < ((x3 && part == 0) ? 2 : 1); wpart++)      // hi activations meet the hi and the lo slabs; lo activations the hi slabs
                    for (int t = 0; t < p->ntaps; t++)
                        if (cls_of[t] == c)
                        {
                            a.tap_row[j] = (p->tap_dy[t] - cmin_y[c]) / st * a.pw + (p->tap_dx[t] - cmin_x[c]) / st;
                            a.tap_slab[j] = (int)((wpart ? lo_row0 : 0) + (long long)t * p->cout);
                            a.cls_ntaps[cc]++; j++;
                        }
            }
        a.ntaps = j;
    }
    a.tiles_x = ceil_div(p->out_w, 8 * mh); a.tiles_y = ceil_div(p->out_h, kV3TileH);
    a.act = p->act; a.alpha = p->alpha; a.gain = p->gain; a.clamp = p->clamp;
    a.accumulate = p->accumulate;
    a.red_x = p->red_x; a.red_out = p->red_out;
    a.a_ready = p->a_ready && !p->a_scale && !x3;
    a.noise = p->noise; a.nsn = p->noise_stride_n; a.nsy = p->noise_stride_y; a.nsx = p->noise_stride_x;
    a.debug = tune.debug;
    const int pixel_tiles = a.tiles_x * a.tiles_y * p->n;
    // N tile: 256 columns leave no room to double-buffer the accumulator (2 halves x 256 = all 512 TMEM columns), so the drain of a
    // tile is exposed; that only pays off when K is long (cin >= 512) and there are enough tiles to fill the SMs.  Measured
    // (profiles/conv_v3_ablation_r1.txt): 256->256 ch 0.263 vs 0.242 ms, 512->512 ch 0.230 vs 0.267 ms for BN = 256 vs 128.
    const bool wide = p->cout % 256 == 0 && tune.max_bn >= 256 && p->cin >= tune.wide_cin && pixel_tiles * (p->cout / 256) >= num_sms();
    const int bn = wide ? 256 : (p->cout % 128 == 0 && tune.max_bn >= 128) ? 128 : 64;
    a.ntiles_n = p->cout / bn;
    if (st == 2 && pixel_tiles * a.ntiles_n < (3 * num_sms()) / 4) return SGV_ERR_UNSUPPORTED;     // too few tiles for one CTA per SM: the per-tap kernel's finer grid wins
    int cl = tune.cluster;
    while (cl > 1 && (pixel_tiles % cl != 0 || pixel_tiles / cl * a.ntiles_n < num_sms() / cl)) cl >>= 1;   // small problems: fill the SMs first
    a.total_groups = pixel_tiles / cl * a.ntiles_n;
    const bool pair = tune.pair && cl == 2;
    if (query)
    {
        query->kernel = 3; query->bn = bn; query->mh = mh; query->cluster = cl; query->cta_pair = pair ? 1 : 0; query->x3 = x3 ? 1 : 0;
        return SGV_OK;
    }

    CUtensorMap tmx, tmw, tmy;
    {
        const uint64_t dims[4] = {(uint64_t)p->cin, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
        const bool view = p->in_stride_x != 0;
        const uint64_t strides[3] = {(uint64_t)(view ? p->in_stride_x : p->cin) * 4, (uint64_t)(view ? p->in_stride_y : (int64_t)p->w * p->cin) * 4,
                                     (uint64_t)(view ? p->in_stride_n : (int64_t)p->h * p->w * p->cin) * 4};
        const uint32_t box[4] = {32, (uint32_t)(a.pw * st), (uint32_t)(a.ph * st), 1};      // ceil(box / elementStride) samples per dim
        const uint32_t es[4] = {1, (uint32_t)st, (uint32_t)st, 1};
        int rc = make_tmap_f32(&tmx, p->x, 4, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p->cin, (uint64_t)slab_rows};
        const uint64_t strides[1] = {(uint64_t)p->cin * 4};
        const uint32_t box[2] = {32, (uint32_t)(bn / cl)};      // cluster of 2 (multicast or CTA pair): each CTA fetches half of the slab rows
        const uint32_t es[2] = {1, 1};
        int rc = make_tmap_f32(&tmw, p->wp, 2, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    {
        // output (possibly a pixel-strided view): box = 32 channels x 8 columns x 16 rows, the unit one epilogue pass stages
        const uint64_t dims[4] = {(uint64_t)p->cout, (uint64_t)p->out_w, (uint64_t)p->out_h, (uint64_t)p->n};
        const uint64_t strides[3] = {(uint64_t)p->out_stride_x * 4, (uint64_t)p->out_stride_y * 4, (uint64_t)p->out_stride_n * 4};
        const uint32_t box[4] = {32, 8, (uint32_t)kV3TileH, 1};
        const uint32_t es[4] = {1, 1, 1, 1};
        int rc = make_tmap_f32(&tmy, p->y, 4, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    if (pair) return launch_v3_pair(bn, mh, tmx, tmw, tmy, a, stream);
    if (cl == 4) return launch_v3_bn<4>(bn, mh, tmx, tmw, tmy, a, stream);
    if (cl == 2) return launch_v3_bn<2>(bn, mh, tmx, tmw, tmy, a, stream);
    return launch_v3_bn<1>(bn, mh, tmx, tmw, tmy, a, stream);
}

} // namespace sgv
