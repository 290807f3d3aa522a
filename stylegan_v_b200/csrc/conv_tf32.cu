// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a) with StyleGAN modulation and bias/activation fused.
//
//   y[n,p,o] = epi( sum_{t,i} (x[n, p*stride + d_t, i] * a[n,i]) * W[t][o][i] )          (include/sgv_b200_conv.h)
//
// GEMM view: M = 128 output pixels per CTA (a TW x TH x TN box of the NHWC tensor), N = BN output channels,
// K = taps x Cin walked in chunks of 32 channels (= one 128-byte swizzle row of TF32).
//
// Warp roles (192 threads):
//   warp 0     TMA producer: per k-step one 4-D tile load of the activation box shifted by the tap offset
//              (out-of-image pixels are zero-filled by TMA = convolution padding for free) and one 2-D load of
//              the [BN x 32] weight slab; both land 128B-swizzled in shared memory and signal full[stage].
//   warps 2-5  operand transform: multiply the staged activation rows by the per-sample modulation
//              a[n, c..c+31] and round to TF32 (round-to-nearest; the tensor core would truncate), in place, then
//              fence.proxy.async and signal ready[stage].  After the k-loop the same warps run the epilogue:
//              tcgen05.ld the fp32 accumulator rows, apply o_scale[n,o] / bias / lrelu / gain / clamp, store NHWC.
//   warp 1     allocates TMEM, issues tcgen05.mma.kind::tf32 (one elected thread, 4 x K=8 per k-step) with the
//              accumulator in TMEM, releases stages with tcgen05.commit -> empty[stage].
//
// Split-K (small planes: the 4x4 ... 16x16 layers have only 4-64 pixel tiles but K = 9 * 512 ... 9 * 1024): the launch is a thread-block
// cluster of KS CTAs per output tile along gridDim.z; CTA r walks k-steps [r*S/KS, (r+1)*S/KS) into its own TMEM accumulator, every CTA
// parks the column chunks it does not own in its (now idle) pipeline stages, and after a cluster barrier CTA r sums chunk cc (cc % KS == r)
// over the cluster through distributed shared memory (ld.shared::cluster) and runs the epilogue for it — a reduce-scatter: no workspace,
// no second launch, no atomics, fixed summation order.
//
// Replaces cuDNN for the reference's conv2d_gradfix.conv2d / conv_transpose2d on the hot-path shapes
// (conv2d_gradfix.py:35-43) plus the x*styles, *dcoefs and bias_act passes around it (networks.py:64-74,141-143).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/sgv_b200_conv.h"

#include <stdlib.h>
#include <string.h>

namespace sgv {

using namespace ptx;

int conv2d_tf32_v3(const sgv_conv_params* p, cudaStream_t stream, sgv_conv_variant* query);     // conv_tf32_v3.cu

constexpr int kConvThreads = 192;
constexpr int kBM = 128;
constexpr int kBK = 32;                       // fp32/TF32 elements per k-step = 128 bytes
constexpr int kATileBytes = kBM * kBK * 4;    // 16 KB

struct ConvArgs
{
    float* y; const float* a_scale; const float* o_scale; const float* bias;
    int n, cin, cout, out_h, out_w;
    long long osn, osy, osx;
    int in_stride, ntaps;
    int tap_dy[SGV_CONV_MAX_TAPS], tap_dx[SGV_CONV_MAX_TAPS];
    int tw, th, tn;                   // activation box: tw*th*tn == 128
    int tiles_x, tiles_y, tiles_nb;
    int act; float alpha, gain, clamp;
    int accumulate;
    const float* red_x; float* red_out;
    // tf32x3 (fp32-grade) mode: every (chunk, tap) k-step is issued three times — part 0: tf32(x*s) x hi slab, part 1: tf32 residual of
    // x*s x hi slab, part 2: tf32(x*s) x lo slab; lo_row0 = first row of the lo slab set in the [rows, cin] slab tensor
    int parts; int lo_row0;
    const float* noise; long long nsn, nsy, nsx;
};

template <int BN, int STAGES>
struct ConvSmem
{
    static constexpr int kBTileBytes = BN * kBK * 4;
    static constexpr int kStageBytes = kATileBytes + kBTileBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (3 * STAGES + 1) * 8 + 16 + 1024;   // + barriers + tmem ptr + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const ConvArgs p)
{
    using L = ConvSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
    uint64_t* ready_bar = full_bar + STAGES;
    uint64_t* empty_bar = ready_bar + STAGES;
    uint64_t* accum_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int mt = blockIdx.x;
    const int tile_x = mt % p.tiles_x; mt /= p.tiles_x;
    const int tile_y = mt % p.tiles_y; mt /= p.tiles_y;
    const int n0 = mt * p.tn;
    const int ox0 = tile_x * p.tw, oy0 = tile_y * p.th;
    const int nb0 = blockIdx.y * BN;                 // first output channel of this CTA
    const int kchunks = p.cin / kBK;
    const int tp_per_chunk = p.ntaps * p.parts;
    const int ksteps_total = kchunks * tp_per_chunk;
    // split-K over the CTAs of the cluster (gridDim.z = cluster size; 1 = no split)
    const int ks_n = (int)gridDim.z, ks_r = (int)blockIdx.z;
    const int k_begin = (int)(((long long)ksteps_total * ks_r) / ks_n), k_end = (int)(((long long)ksteps_total * (ks_r + 1)) / ks_n);
    const int ksteps = k_end - k_begin;

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar + s, 1); mbar_init(ready_bar + s, 4); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    if (warp == 1)
    {
        tmem_alloc(tmem_slot, BN < 32 ? 32 : BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0)
    {
        // ===== TMA producer =====
        if (elect_one())
        {
            int stage = 0; uint32_t phase = 0;
            for (int ks = k_begin; ks < k_end; ks++)
                    {
                        const int kc = ks / tp_per_chunk, tp = ks - kc * tp_per_chunk;
                        const int t = tp / p.parts, part = tp - t * p.parts;
                        mbar_wait(empty_bar + stage, phase ^ 1);
                        uint8_t* sa = smem + stage * L::kStageBytes;
                        uint8_t* sb = sa + kATileBytes;
                        mbar_expect_tx(full_bar + stage, L::kStageBytes);
                        tma_load_4d(sa, &tmap_x, full_bar + stage, kc * kBK, ox0 * p.in_stride + p.tap_dx[t], oy0 * p.in_stride + p.tap_dy[t], n0);
                        tma_load_2d(sb, &tmap_w, full_bar + stage, kc * kBK, (part == 2 ? p.lo_row0 : 0) + t * p.cout + nb0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
        }
    }
    else if (warp == 1)
    {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
        int stage = 0; uint32_t phase = 0;
        for (int ks = 0; ks < ksteps; ks++)
        {
            mbar_wait(ready_bar + stage, phase);
            tc_fence_after();
            if (elect_one())
            {
                const uint32_t sa = smem_u32(smem + stage * L::kStageBytes);
                const uint64_t da = umma_desc_k_sw128(sa);
                const uint64_t db = umma_desc_k_sw128(sa + kATileBytes);
#pragma unroll
                for (int k = 0; k < kBK / 8; k++)      // K = 8 TF32 per instruction = 32 bytes along the swizzled row
                    mma_tf32(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (ks > 0 || k > 0) ? 1u : 0u);
                mma_commit(empty_bar + stage);
                if (ks == ksteps - 1) mma_commit(accum_bar);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    }
    else
    {
        // ===== operand transform, then epilogue =====
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;                 // accumulator row = A-tile row = pixel index inside the box
        const int box_hw = p.th * p.tw;
        const int tn_i = row / box_hw;
        const int rem = row - tn_i * box_hw;
        const int ty = rem / p.tw, tx = rem - ty * p.tw;
        const int n = n0 + tn_i;
        const int oy = oy0 + ty, ox = ox0 + tx;
        const bool valid = (n < p.n) && (oy < p.out_h) && (ox < p.out_w);
        const int nc = n < p.n ? n : p.n - 1;

        {
            int stage = 0; uint32_t phase = 0;
            float sv[kBK];
#pragma unroll
            for (int j = 0; j < kBK; j++) sv[j] = 1.f;
            int cur_kc = -1;
            {
                for (int ks = k_begin; ks < k_end; ks++)
                {
                    const int kc = ks / tp_per_chunk, tp = ks - kc * tp_per_chunk;
                    if (kc != cur_kc && p.a_scale)
                    {
                        const float4* sp = reinterpret_cast<const float4*>(p.a_scale + (long long)nc * p.cin + kc * kBK);
#pragma unroll
                        for (int j = 0; j < kBK / 4; j++) { float4 v = __ldg(sp + j); sv[4 * j] = v.x; sv[4 * j + 1] = v.y; sv[4 * j + 2] = v.z; sv[4 * j + 3] = v.w; }
                    }
                    cur_kc = kc;
                    const bool lo_part = (p.parts == 3) && (tp % 3 == 1);
                    mbar_wait(full_bar + stage, phase);
                    const uint32_t arow = smem_u32(smem + stage * L::kStageBytes) + (uint32_t)row * 128u;
                    float4 v[8];
#pragma unroll
                    for (int j = 0; j < 8; j++) v[j] = lds128(arow + (uint32_t)((j ^ (row & 7)) << 4));   // logical 16-byte chunk j
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
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready_bar + stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }

        // ---- split-K: every CTA of the cluster parks the 32-column chunks of its partial tile [128 rows][BN] that ANOTHER CTA owns
        //      (chunk cc belongs to CTA cc % KS) in its now idle pipeline stages ----
        if (ksteps > 0) { mbar_wait(accum_bar, 0); tc_fence_after(); }
        if (ks_n > 1)
        {
#pragma unroll 1
            for (int cc = 0; cc < BN / 32; cc++)
            {
                if (cc % ks_n == ks_r) continue;
                uint32_t v[32];
                if (ksteps > 0) { tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v); tmem_ld_wait(); }
                else {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] = 0u;
                }
                const uint32_t dst = smem_u32(smem) + (uint32_t)row * (uint32_t)(BN * 4) + (uint32_t)(cc * 128);
#pragma unroll
                for (int j = 0; j < 8; j++)
                    sts128(dst + (uint32_t)(((j + row) & 7) << 4), make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
            }
        }
    }
    if (ks_n > 1) { tc_fence_before(); cluster_sync_all(); tc_fence_after(); }      // partial tiles are visible cluster-wide
    if (warp >= 2)
    {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const int box_hw = p.th * p.tw;
        const int tn_i = row / box_hw;
        const int rem = row - tn_i * box_hw;
        const int ty = rem / p.tw, tx = rem - ty * p.tw;
        const int n = n0 + tn_i;
        const int oy = oy0 + ty, ox = ox0 + tx;
        const bool valid = (n < p.n) && (oy < p.out_h) && (ox < p.out_w);
        const int nc = n < p.n ? n : p.n - 1;
        // ---- epilogue; under split-K a reduce-scatter: each CTA sums and finishes the column chunks it owns, reading the other CTAs' parts of
        //      them through distributed shared memory (the serial chain per CTA is KS times shorter than a reduction in one leader CTA:
        //      b4.conv1 with 8-way split measured 0.099 ms with the leader-only reduction) ----
        float* yrow = p.y + (long long)nc * p.osn + (long long)oy * p.osy + (long long)ox * p.osx + nb0;
        const float* osc = p.o_scale ? p.o_scale + (long long)nc * p.cout + nb0 : nullptr;
        const float* bia = p.bias ? p.bias + nb0 : nullptr;
        const float nz = (p.noise && valid) ? __ldg(p.noise + (long long)nc * p.nsn + (long long)oy * p.nsy + (long long)ox * p.nsx) : 0.f;
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; cc++)
        {
            if (ks_n > 1 && cc % ks_n != ks_r) continue;
            uint32_t v[32];
            if (ksteps > 0) { tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v); tmem_ld_wait(); }
            else {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] = 0u;
            }
            for (int r = 0; r < ks_n; r++)
            {
                if (r == ks_r) continue;
                // partial tile of CTA r of the cluster (same layout, same rotation of the 16-byte chunks as it was written with)
                const uint32_t src = mapa_cluster(smem_u32(smem) + (uint32_t)row * (uint32_t)(BN * 4) + (uint32_t)(cc * 128), (uint32_t)r);
                float4 t4[8];
#pragma unroll
                for (int j = 0; j < 8; j++) t4[j] = ld_dsmem128(src + (uint32_t)(((j + row) & 7) << 4));
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + t4[j].x); v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + t4[j].y);
                    v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + t4[j].z); v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + t4[j].w);
                }
            }
            if (p.red_out)
            {
                // fused style-gradient reduction (see conv_tf32_v3.cu); a warp's 32 rows may straddle samples when the box spills
                // into the batch, so the sum is warp-reduced only when all rows share n, else accumulated per row
                float prod[32];
                const float* rx = p.red_x + (long long)nc * p.osn + (long long)oy * p.osy + (long long)ox * p.osx + nb0 + cc * 32;
#pragma unroll
                for (int j = 0; j < 8; j++)
                {
                    const float4 xv = valid ? __ldg(reinterpret_cast<const float4*>(rx) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    prod[4 * j + 0] = __uint_as_float(v[4 * j + 0]) * xv.x; prod[4 * j + 1] = __uint_as_float(v[4 * j + 1]) * xv.y;
                    prod[4 * j + 2] = __uint_as_float(v[4 * j + 2]) * xv.z; prod[4 * j + 3] = __uint_as_float(v[4 * j + 3]) * xv.w;
                }
                const int n_first = __shfl_sync(0xffffffffu, nc, 0);
                if (__all_sync(0xffffffffu, nc == n_first))
                {
                    const float tot = warp_reduce_32x32(prod, lane);
                    atomicAdd(p.red_out + (long long)n_first * p.cout + nb0 + cc * 32 + lane, tot);
                }
                else if (valid)
                {
#pragma unroll
                    for (int j = 0; j < 32; j++) atomicAdd(p.red_out + (long long)nc * p.cout + nb0 + cc * 32 + j, prod[j]);
                }
            }
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
                        if (p.noise) f = __fadd_rn(f, nz);
                        if (bia) f = __fadd_rn(f, __ldg(bia + col));
                        if (p.act == 3) f = (f > 0.f) ? f : f * p.alpha;
                        f *= p.gain;
                        if (p.clamp >= 0.f) f = (f > -p.clamp && f < p.clamp) ? f : (f >= 0.f ? p.clamp : -p.clamp);
                        o[e] = f;
                    }
                    float4* dst = reinterpret_cast<float4*>(yrow + cc * 32 + j * 4);
                    if (p.accumulate) { const float4 old = *dst; o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w; }
                    *dst = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
        }
    }

    tc_fence_before();
    if (ks_n > 1) cluster_sync_all(); else __syncthreads();      // no CTA leaves while the leader may still read its shared memory
    if (warp == 1) tmem_dealloc(tmem_base, BN < 32 ? 32 : BN);
}

// ---- weight preparation ----
struct TapTable { int ky[SGV_CONV_MAX_TAPS]; int kx[SGV_CONV_MAX_TAPS]; };

__global__ void __launch_bounds__(256) conv_prep_weights_kernel(const float* __restrict__ w, long long sr, long long sc, long long sky, long long skx,
                                                                  int rows, int cols, int ntaps, TapTable taps, float w_scale, float* __restrict__ wp,
                                                                  float* __restrict__ wp_lo)
{
    const long long total = (long long)ntaps * rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    {
        const int k = (int)(i % cols);
        const int r = (int)((i / cols) % rows);
        const int t = (int)(i / ((long long)cols * rows));
        const float v = __fmul_rn(w[r * sr + k * sc + taps.ky[t] * sky + taps.kx[t] * skx], w_scale);
        wp[i] = ptx::tf32_rn(v);
        if (wp_lo) wp_lo[i] = ptx::tf32_lo(v);
    }
}


// Both slab sets of one layer from ONE coalesced read of a dense [O][I][kh][kw] weight: wp_a[t][o][i] (forward: rows = O) and
// wp_b[t][i][o] (data gradient: rows = I), each with its own tap order.  A CTA stages a 32 (o) x 32 (i) x KK block — 32 runs of
// 32*KK contiguous floats — in shared memory and writes every slab row as 128 contiguous bytes.  (The gather kernel above reads
// with a 4*KK-byte stride: 16 us per call on a 512x512x3x3 weight, 44 calls per step.)
struct TapPair { int na, nb; int a[SGV_CONV_MAX_TAPS]; int b[SGV_CONV_MAX_TAPS]; };      // tap = ky * kw + kx

template <int KK>
__global__ void __launch_bounds__(256) conv_prep_pair_kernel(const float* __restrict__ w, int O, int I, TapPair taps, float* __restrict__ wa, float* __restrict__ wb,
                                                               float* __restrict__ wa_lo, float* __restrict__ wb_lo)
{
    __shared__ float tile[32][32 * KK + 1];
    const int i0 = blockIdx.x * 32, o0 = blockIdx.y * 32;
    const bool split = wa_lo != nullptr || wb_lo != nullptr;      // tf32x3: the tile keeps the fp32 value, hi / lo are formed at the store
    for (int idx = threadIdx.x; idx < 32 * 32 * KK; idx += 256)
    {
        const int ol = idx / (32 * KK), rem = idx - ol * (32 * KK);
        const float v = w[((long long)(o0 + ol) * I + i0) * KK + rem];
        tile[ol][rem] = split ? v : ptx::tf32_rn(v);
    }
    __syncthreads();
    const int lo = threadIdx.x & 31, hi = threadIdx.x >> 5;      // 8 row groups
    if (wa)
        for (int t = 0; t < taps.na; t++)
            for (int ol = hi; ol < 32; ol += 8)
            {
                const float v = tile[ol][lo * KK + taps.a[t]];
                const long long at = ((long long)t * O + o0 + ol) * I + i0 + lo;
                wa[at] = ptx::tf32_rn(v);
                if (wa_lo) wa_lo[at] = ptx::tf32_lo(v);
            }
    if (wb)
        for (int t = 0; t < taps.nb; t++)
            for (int il = hi; il < 32; il += 8)
            {
                const float v = tile[lo][il * KK + taps.b[t]];
                const long long at = ((long long)t * I + i0 + il) * O + o0 + lo;
                wb[at] = ptx::tf32_rn(v);
                if (wb_lo) wb_lo[at] = ptx::tf32_lo(v);
            }
}

// ---- host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn)
    {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides, bool atom32)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(SGV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base),
                    reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                    reinterpret_cast<const cuuint32_t*>(box), reinterpret_cast<const cuuint32_t*>(elem_strides),
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SGV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SGV_OK;
}

template <int BN, int STAGES>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvArgs& a, dim3 grid, cudaStream_t stream)
{
    using L = ConvSmem<BN, STAGES>;
    auto kern = conv_tf32_kernel<BN, STAGES>;
    SGV_OPT_IN_SMEM(kern, L::kTotal);
    if (grid.z > 1)
    {
        // split-K: the gridDim.z CTAs of an output tile form a thread-block cluster (partial tiles are reduced through distributed shared memory)
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = grid.z;
        cfg.gridDim = grid; cfg.blockDim = dim3(kConvThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = stream;
        cfg.attrs = attr; cfg.numAttrs = 1;
        SGV_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tx, tw, a));
    }
    else
        kern<<<grid, kConvThreads, L::kTotal, stream>>>(tx, tw, a);
    SGV_LAUNCH_OK("conv_tf32_kernel");
    return SGV_OK;
}

} // namespace sgv

extern "C" int sgv_conv_prep_weights_pair(const float* w, int32_t out_ch, int32_t in_ch, int32_t kh, int32_t kw,
                                          int32_t ntaps_a, const int32_t* a_ky, const int32_t* a_kx, float* wp_a,
                                          int32_t ntaps_b, const int32_t* b_ky, const int32_t* b_kx, float* wp_b, void* stream_)
{
    return sgv_conv_prep_weights_pair_x3(w, out_ch, in_ch, kh, kw, ntaps_a, a_ky, a_kx, wp_a, nullptr, ntaps_b, b_ky, b_kx, wp_b, nullptr, stream_);
}

extern "C" int sgv_conv_prep_weights_pair_x3(const float* w, int32_t out_ch, int32_t in_ch, int32_t kh, int32_t kw,
                                             int32_t ntaps_a, const int32_t* a_ky, const int32_t* a_kx, float* wp_a, float* wp_a_lo,
                                             int32_t ntaps_b, const int32_t* b_ky, const int32_t* b_kx, float* wp_b, float* wp_b_lo, void* stream_)
{
    using namespace sgv;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGV_CHECK_ARG(w && (wp_a || wp_b), "sgv_conv_prep_weights_pair: NULL argument");
    SGV_CHECK_ARG((wp_a || !wp_a_lo) && (wp_b || !wp_b_lo), "a lo slab set needs its hi set");
    SGV_CHECK_ARG((kh == 3 && kw == 3) || (kh == 1 && kw == 1), "kernel must be 1x1 or 3x3");
    SGV_CHECK_ARG(out_ch % 32 == 0 && in_ch % 32 == 0 && out_ch > 0 && in_ch > 0, "channel counts must be positive multiples of 32");
    SGV_CHECK_ARG(ntaps_a >= 0 && ntaps_a <= SGV_CONV_MAX_TAPS && ntaps_b >= 0 && ntaps_b <= SGV_CONV_MAX_TAPS, "ntaps must be in [0, %d]", SGV_CONV_MAX_TAPS);
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    TapPair tp;
    tp.na = wp_a ? ntaps_a : 0; tp.nb = wp_b ? ntaps_b : 0;
    for (int t = 0; t < tp.na; t++) { SGV_CHECK_ARG(a_ky[t] >= 0 && a_ky[t] < kh && a_kx[t] >= 0 && a_kx[t] < kw, "tap out of range"); tp.a[t] = a_ky[t] * kw + a_kx[t]; }
    for (int t = 0; t < tp.nb; t++) { SGV_CHECK_ARG(b_ky[t] >= 0 && b_ky[t] < kh && b_kx[t] >= 0 && b_kx[t] < kw, "tap out of range"); tp.b[t] = b_ky[t] * kw + b_kx[t]; }
    dim3 grid((unsigned)(in_ch / 32), (unsigned)(out_ch / 32));
    if (kh == 3) conv_prep_pair_kernel<9><<<grid, 256, 0, stream>>>(w, out_ch, in_ch, tp, wp_a, wp_b, wp_a_lo, wp_b_lo);
    else conv_prep_pair_kernel<1><<<grid, 256, 0, stream>>>(w, out_ch, in_ch, tp, wp_a, wp_b, wp_a_lo, wp_b_lo);
    SGV_LAUNCH_OK("conv_prep_pair_kernel");
    return SGV_OK;
}

extern "C" int sgv_conv_prep_weights(const float* w, int64_t stride_row, int64_t stride_col, int64_t stride_ky, int64_t stride_kx,
                                     int32_t rows, int32_t cols, int32_t ntaps, const int32_t* tap_ky, const int32_t* tap_kx,
                                     float* wp, void* stream_)
{
    return sgv_conv_prep_weights_ex(w, stride_row, stride_col, stride_ky, stride_kx, rows, cols, ntaps, tap_ky, tap_kx, 1.0f, wp, nullptr, stream_);
}

extern "C" int sgv_conv_prep_weights_ex(const float* w, int64_t stride_row, int64_t stride_col, int64_t stride_ky, int64_t stride_kx,
                                        int32_t rows, int32_t cols, int32_t ntaps, const int32_t* tap_ky, const int32_t* tap_kx,
                                        float w_scale, float* wp, float* wp_lo, void* stream_)
{
    using namespace sgv;
    cudaStream_t stream = (cudaStream_t)stream_;
    SGV_CHECK_ARG(w && wp && tap_ky && tap_kx, "sgv_conv_prep_weights: NULL argument");
    SGV_CHECK_ARG(ntaps >= 1 && ntaps <= SGV_CONV_MAX_TAPS, "ntaps must be in [1, %d]", SGV_CONV_MAX_TAPS);
    SGV_CHECK_ARG(rows >= 1 && cols >= 1, "rows/cols must be positive");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;
    TapTable taps;
    for (int t = 0; t < SGV_CONV_MAX_TAPS; t++) { taps.ky[t] = t < ntaps ? tap_ky[t] : 0; taps.kx[t] = t < ntaps ? tap_kx[t] : 0; }
    const long long total = (long long)ntaps * rows * cols;
    const unsigned grid = (unsigned)min((long long)num_sms() * 8, (total + 255) / 256);
    conv_prep_weights_kernel<<<grid, 256, 0, stream>>>(w, stride_row, stride_col, stride_ky, stride_kx, rows, cols, ntaps, taps, w_scale, wp, wp_lo);
    SGV_LAUNCH_OK("conv_prep_weights_kernel");
    return SGV_OK;
}

static int pow2_floor(int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; }
static int pow2_ceil(int v) { int p = 1; while (p < v) p *= 2; return p; }

static int conv2d_tf32_dispatch(const sgv_conv_params* p, cudaStream_t stream, sgv_conv_variant* query);

extern "C" int sgv_conv2d_tf32(const sgv_conv_params* p, void* stream_)
{
    return conv2d_tf32_dispatch(p, (cudaStream_t)stream_, nullptr);
}

extern "C" int sgv_conv2d_tf32_variant(const sgv_conv_params* p, sgv_conv_variant* out)
{
    SGV_CHECK_ARG(out != nullptr, "sgv_conv2d_tf32_variant: out is NULL");
    memset(out, 0, sizeof(*out));
    return conv2d_tf32_dispatch(p, nullptr, out);
}

static int conv2d_tf32_dispatch(const sgv_conv_params* p, cudaStream_t stream, sgv_conv_variant* query)
{
    using namespace sgv;
    SGV_CHECK_ARG(p != nullptr, "sgv_conv2d_tf32: params is NULL");
    SGV_CHECK_ARG(p->x && p->wp && p->y, "sgv_conv2d_tf32: x, wp and y must be non-NULL");
    SGV_CHECK_ARG(p->n >= 1 && p->h >= 1 && p->w >= 1 && p->out_h >= 1 && p->out_w >= 1, "extents must be positive");
    SGV_CHECK_ARG(p->cin >= 32 && p->cin % 32 == 0, "cin must be a multiple of 32 (got %d)", p->cin);
    // 16 output channels are rejected: the epilogue reads the accumulator in 32-column TMEM slices (tmem_ld_32x32), so a 16-wide N tile ran
    // zero epilogue iterations and stored nothing (found on the GPU in round 1, profiles/debug_d_layers_cout16_r1.txt).  No reference
    // configuration has such a layer; callers take the library path for it (stylegan_v_b200/native_conv.py::_ok_channels).
    SGV_CHECK_ARG(p->cout % 64 == 0 || p->cout == 32, "cout must be a multiple of 64, or 32 (got %d)", p->cout);
    SGV_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= SGV_CONV_MAX_TAPS, "ntaps must be in [1, %d]", SGV_CONV_MAX_TAPS);
    SGV_CHECK_ARG(p->in_stride == 1 || p->in_stride == 2, "in_stride must be 1 or 2");
    SGV_CHECK_ARG(p->act == 1 || p->act == 3, "act must be 1 (linear) or 3 (lrelu)");
    SGV_CHECK_ARG((reinterpret_cast<uintptr_t>(p->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->wp) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(p->y) & 15) == 0, "x, wp and y must be 16-byte aligned");
    SGV_CHECK_ARG(p->out_stride_n % 4 == 0 && p->out_stride_y % 4 == 0 && p->out_stride_x % 4 == 0, "output strides must be multiples of 4 elements");
    SGV_CHECK_ARG((long long)p->n * p->h * p->w * p->cin <= 0x7fffffffLL, "x is too large");
    SGV_CHECK_ARG(!p->accumulate || (p->o_scale == nullptr && p->bias == nullptr && p->act == 1 && p->gain == 1.0f && p->clamp < 0.f),
                  "accumulate=1 cannot be combined with o_scale / bias / activation / gain / clamp");
    SGV_CHECK_ARG((p->red_out == nullptr) == (p->red_x == nullptr), "red_x and red_out must be given together");
    SGV_CHECK_ARG((p->in_stride_x == 0 && p->in_stride_y == 0 && p->in_stride_n == 0) ||
                  (p->in_stride_x % 4 == 0 && p->in_stride_y % 4 == 0 && p->in_stride_n % 4 == 0 && p->in_stride_x > 0),
                  "input view strides must be positive multiples of 4 elements (or all zero for a dense tensor)");
    const bool x3 = p->wp_lo != nullptr;
    SGV_CHECK_ARG(!x3 || (p->wp_lo >= p->wp && (p->wp_lo - p->wp) % p->cin == 0 && (p->wp_lo - p->wp) / p->cin < 0x40000000LL &&
                          (reinterpret_cast<uintptr_t>(p->wp_lo) & 15) == 0), "wp_lo must follow wp in the same allocation at a multiple of cin elements");
    SGV_CHECK_ARG(!x3 || !p->a_ready, "tf32x3 mode splits the activations inside the kernel: a_ready must be 0");
    SGV_CHECK_ARG(!p->noise || !p->accumulate, "noise cannot be combined with accumulate=1");
    int rc = sgv_device_check();
    if (rc != SGV_OK) return rc;

    // the persistent halo-patch kernel (csrc/conv_tf32_v3.cu) covers planes >= 12x12; SGV_CONV_V1=1 forces the per-tap kernel below (A/B runs)
    static const bool force_v1 = env_int("SGV_CONV_V1", 0) != 0;
    if (!force_v1)
    {
        rc = conv2d_tf32_v3(p, stream, query);
        if (rc != SGV_ERR_UNSUPPORTED) return rc;
    }

    ConvArgs a;
    a.y = p->y; a.a_scale = p->a_scale; a.o_scale = p->o_scale; a.bias = p->bias;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    a.osn = p->out_stride_n; a.osy = p->out_stride_y; a.osx = p->out_stride_x;
    a.in_stride = p->in_stride; a.ntaps = p->ntaps;
    for (int t = 0; t < SGV_CONV_MAX_TAPS; t++) { a.tap_dy[t] = p->tap_dy[t]; a.tap_dx[t] = p->tap_dx[t]; }
    a.act = p->act; a.alpha = p->alpha; a.gain = p->gain; a.clamp = p->clamp;
    a.accumulate = p->accumulate;
    a.red_x = p->red_x; a.red_out = p->red_out;
    a.parts = x3 ? 3 : 1;
    a.lo_row0 = x3 ? (int)((p->wp_lo - p->wp) / p->cin) : 0;
    a.noise = p->noise; a.nsn = p->noise_stride_n; a.nsy = p->noise_stride_y; a.nsx = p->noise_stride_x;

    // activation box of 128 output pixels: as square as the plane allows, spilling into the batch dimension for tiny planes
    int tw = pow2_floor(p->out_w < 16 ? p->out_w : 16);
    if (tw < p->out_w && tw < 16 && pow2_ceil(p->out_w) <= 16) tw = pow2_ceil(p->out_w);
    int th = 128 / tw;
    if (th > pow2_ceil(p->out_h)) th = pow2_ceil(p->out_h);
    int tn = 128 / (tw * th);
    a.tw = tw; a.th = th; a.tn = tn;
    a.tiles_x = ceil_div(p->out_w, tw); a.tiles_y = ceil_div(p->out_h, th); a.tiles_nb = ceil_div(p->n, tn);

    // N tile and split-K factor for the few pixel tiles of the 4x4 ... 16x16 planes (K = 9 * 512 ... 9 * 1024): K is split over a thread-block
    // cluster of KS <= 4 CTAs (>= 16 k-steps each; partial tiles are reduce-scattered through distributed shared memory, see the kernel) and the
    // N tile is chosen so that CTAs x KS fills the SMs in ONE wave as far as possible; ties go to the wider tile (a k-step costs about the same
    // whatever the N tile: its pace is set by the staging pass over the 128-row A tile).  Measured sweep (profiles/bench_conv_small_r2j.jsonl,
    // 32 frames): b4.conv1 (4 pixel tiles) 0.0455 ms at (64, 4), 0.060 at (256, 4), 0.070-0.088 at KS = 8; b8.conv1 (16 tiles) 0.0475 ms at
    // (256, 4) = (128, 2), 0.066 at (128, 4) (two waves), 0.13 at KS = 8 (clusters of 8 schedule poorly and lengthen the reduction).
    static const int split_k = env_int("SGV_CONV_SPLITK", 1);
    const int mtiles_total = a.tiles_x * a.tiles_y * a.tiles_nb;
    const int ksteps_all = (p->cin / kBK) * p->ntaps * a.parts;
    int bn = (p->cout % 256 == 0) ? 256 : (p->cout % 128 == 0) ? 128 : (p->cout % 64 == 0) ? 64 : p->cout;
    int ks = 1;
    {
        int best_bn = bn, best_ks = 1, best_fill = -1;
        for (int cand = bn; cand >= 64 && p->cout % cand == 0; cand /= 2)
        {
            const int ctas = mtiles_total * (p->cout / cand);
            int k = 1;
            while (split_k && k < 4 && ctas * k * 2 <= num_sms() && ksteps_all / (k * 2) >= 16) k *= 2;
            const int fill = ctas * k <= num_sms() ? ctas * k : 0;          // more than one wave: never better than the wider tile
            if (fill > best_fill) { best_fill = fill; best_bn = cand; best_ks = k; }
        }
        bn = best_bn; ks = best_ks;
    }

    // tuning overrides (read once; used by scripts/bench_conv.py sweeps): force the N tile and / or the split factor when they are legal
    static const int force_bn = env_int("SGV_CONV_V1_BN", 0), force_ks = env_int("SGV_CONV_V1_KS", 0);
    if (force_bn > 0 && force_bn <= 256 && p->cout % force_bn == 0 && (force_bn & (force_bn - 1)) == 0 && force_bn >= 32) bn = force_bn;
    if (force_ks > 0 && force_ks <= 8 && (force_ks & (force_ks - 1)) == 0 && ksteps_all / force_ks >= 1) ks = force_ks;

    if (query)
    {
        query->kernel = 1; query->bn = bn; query->mh = 1; query->cluster = ks; query->cta_pair = 0; query->x3 = x3 ? 1 : 0;
        return SGV_OK;
    }
    CUtensorMap tmx, tmw;
    {
        const uint64_t dims[4] = {(uint64_t)p->cin, (uint64_t)p->w, (uint64_t)p->h, (uint64_t)p->n};
        const bool view = p->in_stride_x != 0;
        const uint64_t strides[3] = {(uint64_t)(view ? p->in_stride_x : p->cin) * 4, (uint64_t)(view ? p->in_stride_y : (int64_t)p->w * p->cin) * 4,
                                     (uint64_t)(view ? p->in_stride_n : (int64_t)p->h * p->w * p->cin) * 4};
        const uint32_t box[4] = {(uint32_t)kBK, (uint32_t)(tw * p->in_stride), (uint32_t)(th * p->in_stride), (uint32_t)tn};
        const uint32_t es[4] = {1, (uint32_t)p->in_stride, (uint32_t)p->in_stride, 1};
        SGV_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "activation box too large");
        rc = make_tmap_f32(&tmx, p->x, 4, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)p->cin, (uint64_t)a.lo_row0 + (uint64_t)p->ntaps * p->cout};
        const uint64_t strides[1] = {(uint64_t)p->cin * 4};
        const uint32_t box[2] = {(uint32_t)kBK, (uint32_t)bn};
        const uint32_t es[2] = {1, 1};
        rc = make_tmap_f32(&tmw, p->wp, 2, dims, strides, box, es);
        if (rc != SGV_OK) return rc;
    }
    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * a.tiles_nb), (unsigned)(p->cout / bn), (unsigned)ks);
    switch (bn)
    {
        case 256: return launch_conv<256, 4>(tmx, tmw, a, grid, stream);
        case 128: return launch_conv<128, 6>(tmx, tmw, a, grid, stream);
        case 64:  return launch_conv<64, 8>(tmx, tmw, a, grid, stream);
        case 32:  return launch_conv<32, 8>(tmx, tmw, a, grid, stream);
        default:  return fail(SGV_ERR_UNSUPPORTED, "sgv_conv2d_tf32: no kernel for an N tile of %d columns", bn);
    }
}
