// Weight gradient of a 3x3 stride-1 convolution with 64 OUTPUT channels on tcgen05 tensor cores (sm_100a): "stacked-M" kernel.
//
//   dw[t][o][i] += sum_{n,p} (g[n, p, o] * g_scale[n,o]) * (x[n, p + d_t, i] * x_scale[n,i]),   t over the 9 taps d_t in {-1,0,1}^2
//
// Why a kernel of its own.  The grouped-tap kernel (wgrad_tf32.cu) puts the gradient's channels on the 128 rows of the MMA.  With 64 output
// channels (the 256^2 layers of the 256^2 network, the 512^2 layers of the 1024^2 one) half of every MMA multiplies zero padding, the
// three dy-groups of CTAs each re-read both operands (ncu, profiles/ncu_r2b_summary.txt: 2.48 GB of DRAM reads for 1.07 GB of operands,
// tensor pipe 39 % active, 0.80 ms per launch — the slowest kernel of the step).  Here the other 64 rows carry a SECOND, pixel-shifted copy
// of the same gradient tile: since  sum_p g[p + e] x[p + s] = sum_q g[q] x[q + s - e],  rows 64..127 of an MMA with the input patch shifted
// by s accumulate tap (s - e) while rows 0..63 accumulate tap s.  With two stackings
//     A1 = [ g | g shifted one image row ]       s = (+1, dx), dx = -1, 0, +1    ->  taps (+1, dx) and (0, dx)
//     A2 = [ g shifted one pixel | g ]           s = (-1, +1) and (-1, -1)       ->  taps (-1, 0), (-1, +1) and (-1, -1)
// five MMAs per 8-pixel k-row produce all nine taps (9 of 10 row halves useful instead of 9 of 18), ONE CTA owns all taps of a
// [64 x 64] block (each operand is fetched once: the three gradient copies are three TMA boxes of the same L2-resident lines), and the
// shared-memory operand reads per useful FLOP drop by 1.8x.  The pixel lattice is walked from (-1, -1) so that the shifted copies see
// every gradient pixel; TMA zero-fills outside the image.
//
// Layout per stage (40 KB): [g shifted one pixel | g | g shifted one row] = 3 x [2 channel blocks][32 pixels][32 ch] (MN-major,
// SWIZZLE_128B_ATOM_32B: A1 = 4 blocks starting at the middle copy, A2 = 4 blocks starting at the first), then the input patch
// [NT/32 blocks][6 x 10 pixels][32 ch].  Warp roles as in wgrad_tf32.cu: warp 0 TMA, warp 1 MMA issue + TMEM, warps 2-9 operand
// staging (scale, TF32 rounding or its residual for the tf32x3 passes) and the atomic epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "../../include/sgv_b200_conv.h"

#include <string.h>

namespace sgv {

using namespace ptx;

constexpr int kS64Threads = 64 + 256;
constexpr int kS64GCopy = 2 * 32 * 128;              // one gradient copy: 2 channel blocks x 32 pixels x 128 B = 8 KB
constexpr int kS64XBlock = 6 * 10 * 128;             // one 32-channel block of the 6 x 10 pixel input patch = 7680 B (a multiple of the 512 B swizzle atom)
constexpr int kS64NT = 64;
constexpr int kS64XTile = ((kS64NT / 32) * kS64XBlock + 1023) & ~1023;
constexpr int kS64Stage = 3 * kS64GCopy + kS64XTile;  // 40 KB
constexpr int kS64Stages = 5;
constexpr int kS64BarOffset = kS64Stages * kS64Stage;
constexpr int kS64Total = kS64BarOffset + (3 * kS64Stages + 1) * 8 + 16 + 1024;
constexpr int kS64Accs = 5;

struct S64Args
{
    float* dw; const float* g_scale; const float* x_scale;
    int n, cin, cout, out_h, out_w;
    int tiles_x, tiles_y, ktiles, ksplit;
    int slot[kS64Accs][2];            // dw block of (accumulator j, row half): -1 = the wasted half
    int g_ready, x_ready, g_lo, x_lo;
};

__global__ void __launch_bounds__(kS64Threads, 1)
wgrad_tf32_s64_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x, const S64Args p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kS64BarOffset);
    uint64_t* ready_bar = full_bar + kS64Stages;
    uint64_t* empty_bar = ready_bar + kS64Stages;
    uint64_t* accum_bar = empty_bar + kS64Stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * kS64NT;               // first input channel (columns of dw)
    const int per = (p.ktiles + p.ksplit - 1) / p.ksplit;
    const int kt0 = blockIdx.y * per;
    const int kt1 = min(kt0 + per, p.ktiles);
    const int ksteps = kt1 - kt0;

    if (threadIdx.x == 0)
    {
        prefetch_tmap(&tmap_g);
        prefetch_tmap(&tmap_x);
        for (int s = 0; s < kS64Stages; s++) { mbar_init(full_bar + s, 1); mbar_init(ready_bar + s, 8); mbar_init(empty_bar + s, 1); }
        mbar_init(accum_bar, 1);
        fence_mbar_init();
    }
    constexpr int kCols = 512;                        // 5 accumulators x 64 columns
    if (warp == 1) { tmem_alloc(tmem_slot, kCols); tmem_relinquish(); }
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
                    const int px0 = tx * 8 - 1, py0 = ty * 4 - 1, n = r;      // the lattice starts at (-1, -1): see the header
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    uint8_t* st = smem + stage * kS64Stage;
                    mbar_expect_tx(full_bar + stage, 3 * kS64GCopy + (kS64NT / 32) * kS64XBlock);
                    tma_load_5d(st, &tmap_g, full_bar + stage, 0, px0 + 1, py0, n, 0);                       // g shifted one pixel
                    tma_load_5d(st + kS64GCopy, &tmap_g, full_bar + stage, 0, px0, py0, n, 0);               // g
                    tma_load_5d(st + 2 * kS64GCopy, &tmap_g, full_bar + stage, 0, px0, py0 + 1, n, 0);       // g shifted one image row
                    tma_load_5d(st + 3 * kS64GCopy, &tmap_x, full_bar + stage, 0, px0 - 1, py0 - 1, n, c0 / 32);
                    if (++stage == kS64Stages) { stage = 0; phase ^= 1; }
                }
            }
        }
        else if (warp == 1)
        {
            constexpr uint32_t idesc = umma_idesc_tf32(128, kS64NT, 1, 1);
            if (elect_one())
            {
                int stage = 0; uint32_t phase = 0;
                for (int ks = 0; ks < ksteps; ks++)
                {
                    mbar_wait(ready_bar + stage, phase);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + stage * kS64Stage);
                    const uint32_t a1 = st + kS64GCopy, a2 = st, sx = st + 3 * kS64GCopy;
#pragma unroll
                    for (int k = 0; k < 4; k++)                       // image row k of the 8 x 4 tile = 8 consecutive pixel rows (K = 8)
                    {
                        const uint64_t da1 = umma_desc_mn_sw128_32b(a1 + k * 1024, 32 * 128, 512);
                        const uint64_t da2 = umma_desc_mn_sw128_32b(a2 + k * 1024, 32 * 128, 512);
                        const uint32_t acc_flag = (ks > 0 || k > 0) ? 1u : 0u;
#pragma unroll
                        for (int j = 0; j < 3; j++)                   // s = (+1, j - 1): patch row k + 2, column j
                            mma_tf32(tmem_base + (uint32_t)(j * kS64NT), da1, umma_desc_mn_sw128_32b(sx + (uint32_t)((k + 2) * 10 + j) * 128u, kS64XBlock, 512), idesc, acc_flag);
                        // s = (-1, +1): patch row k, column 2;  s = (-1, -1): patch row k, column 0
                        mma_tf32(tmem_base + 3u * kS64NT, da2, umma_desc_mn_sw128_32b(sx + (uint32_t)(k * 10 + 2) * 128u, kS64XBlock, 512), idesc, acc_flag);
                        mma_tf32(tmem_base + 4u * kS64NT, da2, umma_desc_mn_sw128_32b(sx + (uint32_t)(k * 10 + 0) * 128u, kS64XBlock, 512), idesc, acc_flag);
                    }
                    mma_commit(empty_bar + stage);
                    if (ks == ksteps - 1) mma_commit(accum_bar);
                    if (++stage == kS64Stages) { stage = 0; phase ^= 1; }
                }
            }
            __syncwarp();
        }
        else
        {
            // 8 staging warps.  Unified row space: rows [0, 192) = the three gradient copies (copy = row / 64, channel block = (row / 32) % 2),
            // rows [192, 192 + 120) = the input patch (channel block = (row - 192) / 60).  A thread owns rows tid and tid + 256.
            const int tid = threadIdx.x - 64;
            constexpr int kGRows = 192, kXRows = (kS64NT / 32) * 60;
            float sv[2][32];
            int cur_n = -1;
            const FastDiv div_plane((uint32_t)(p.tiles_x * p.tiles_y));
            int rr[2]; const float* sbase[2]; int sstride[2]; bool act[2]; bool lo[2]; uint32_t roff[2];
#pragma unroll
            for (int s = 0; s < 2; s++)
            {
                rr[s] = tid + s * 256;
                act[s] = rr[s] < kGRows + kXRows;
                sbase[s] = nullptr; sstride[s] = 0; lo[s] = false; roff[s] = 0;
                if (act[s])
                {
                    if (rr[s] < kGRows)
                    {
                        const int ch = ((rr[s] >> 5) & 1) * 32;
                        act[s] = !p.g_ready; lo[s] = p.g_lo != 0;
                        if (p.g_scale) { sbase[s] = p.g_scale + ch; sstride[s] = p.cout; }
                        roff[s] = (uint32_t)rr[s] * 128u;                                   // copies and blocks are contiguous 4 KB pieces
                    }
                    else
                    {
                        const int xr = rr[s] - kGRows, blk = xr / 60, pix = xr - blk * 60;
                        act[s] = !p.x_ready; lo[s] = p.x_lo != 0;
                        if (p.x_scale) { sbase[s] = p.x_scale + c0 + blk * 32; sstride[s] = p.cin; }
                        roff[s] = (uint32_t)(3 * kS64GCopy + blk * kS64XBlock + pix * 128);
                    }
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
                    const uint32_t st = smem_u32(smem + stage * kS64Stage);
#pragma unroll
                    for (int s = 0; s < 2; s++)
                    {
                        if (!act[s]) continue;
                        const uint32_t rowp = st + roff[s];
                        const int row = (int)(roff[s] >> 7);            // 128-byte row index inside the stage: swizzle phase = absolute address bits
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
                    if (lane == 0) mbar_arrive(ready_bar + stage);
                    if (++stage == kS64Stages) { stage = 0; phase ^= 1; }
                }
            }
            // ---- epilogue: rows 0..63 / 64..127 of accumulator j are two different taps of the same 64 output channels ----
            mbar_wait(accum_bar, 0);
            tc_fence_after();
            const int q = warp & 3;                               // TMEM lane quarter
            const int hsel = q >> 1;                              // row half
            const int o = (q & 1) * 32 + lane;                    // output channel
            const int part = (warp - 2) >> 2;                     // two warps per lane quarter split the (accumulator, column chunk) pairs
#pragma unroll 1
            for (int u = part; u < kS64Accs * 2; u += 2)
            {
                const int j = u >> 1, cc = u & 1;
                const int slot = p.slot[j][hsel];
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * kS64NT + cc * 32), v);
                tmem_ld_wait();
                if (slot >= 0)
                {
                    float* drow = p.dw + ((long long)slot * p.cout + o) * p.cin + c0 + cc * 32;
#pragma unroll
                    for (int e = 0; e < 8; e++)
                        atomicAdd(reinterpret_cast<float4*>(drow + e * 4),
                                  make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]), __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3])));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kCols);
}

// Returns SGV_ERR_UNSUPPORTED when the call is not a full 3x3 stride-1 correlation with 64 output channels (caller uses the grouped-tap kernel).
int conv2d_wgrad_tf32_s64(const sgv_wgrad_params* p, int g_lo, int x_lo, cudaStream_t stream, sgv_wgrad_variant* query)
{
    static const int enabled = env_int("SGV_WGRAD_S64", 1);
    if (!enabled) return SGV_ERR_UNSUPPORTED;
    if (p->cout != 64 || p->cin % kS64NT != 0 || p->ntaps != 9 || p->g_stride != 1 || p->x_stride != 1 || p->out_w < 8 || p->out_h < 4) return SGV_ERR_UNSUPPORTED;
    if (p->x_stride_x != 0 || p->gh != p->out_h || p->gw != p->out_w) return SGV_ERR_UNSUPPORTED;
    int slot_of[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) slot_of[a][b] = -1;
    for (int t = 0; t < 9; t++)
    {
        if (p->g_dy[t] != 0 || p->g_dx[t] != 0) return SGV_ERR_UNSUPPORTED;
        const int dy = p->x_dy[t], dx = p->x_dx[t];
        if (dy < -1 || dy > 1 || dx < -1 || dx > 1 || slot_of[dy + 1][dx + 1] >= 0) return SGV_ERR_UNSUPPORTED;
        slot_of[dy + 1][dx + 1] = p->use_dw_slot ? p->dw_slot[t] : t;
    }
    S64Args a;
    memset(&a, 0, sizeof(a));
    a.dw = p->dw; a.g_scale = p->g_scale; a.x_scale = p->x_scale;
    a.n = p->n; a.cin = p->cin; a.cout = p->cout; a.out_h = p->out_h; a.out_w = p->out_w;
    // accumulators 0..2: A1 with s = (+1, dx): rows 0..63 tap (+1, dx), rows 64..127 tap (0, dx)
    for (int j = 0; j < 3; j++) { a.slot[j][0] = slot_of[2][j]; a.slot[j][1] = slot_of[1][j]; }
    // accumulator 3: A2 with s = (-1, +1): rows 0..63 tap (-1, 0), rows 64..127 tap (-1, +1);  accumulator 4: s = (-1, -1): rows 64..127 tap (-1, -1)
    a.slot[3][0] = slot_of[0][1]; a.slot[3][1] = slot_of[0][2];
    a.slot[4][0] = -1;            a.slot[4][1] = slot_of[0][0];
    a.tiles_x = ceil_div(p->out_w + 1, 8); a.tiles_y = ceil_div(p->out_h + 1, 4);
    a.ktiles = a.tiles_x * a.tiles_y * p->n;
    const int ntiles = p->cin / kS64NT;
    int ksplit = max(1, num_sms() / ntiles);                 // one CTA per SM
    while (ksplit > 1 && a.ktiles / ksplit < 32) ksplit--;
    if (ksplit > a.ktiles) ksplit = a.ktiles;
    a.ksplit = ksplit;
    a.g_ready = p->g_ready && !p->g_scale && !p->precision; a.x_ready = p->x_ready && !p->x_scale && !p->precision;
    a.g_lo = g_lo; a.x_lo = x_lo;
    if (query)
    {
        query->kernel = 3; query->nt = kS64NT; query->stages = kS64Stages; query->ksplit = ksplit; query->passes = 1;
        return SGV_OK;
    }
    CUtensorMap tg, tx;
    {
        const uint64_t dims[5] = {32, (uint64_t)p->gw, (uint64_t)p->gh, (uint64_t)p->n, (uint64_t)(p->cout / 32)};
        const uint64_t strides[4] = {(uint64_t)p->cout * 4, (uint64_t)p->gw * p->cout * 4, (uint64_t)p->gh * p->gw * p->cout * 4, 128};
        const uint32_t box[5] = {32, 8, 4, 1, 2};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        int rc = make_tmap_f32(&tg, p->g, 5, dims, strides, box, es, /*atom32=*/true);
        if (rc != SGV_OK) return rc;
    }
    {
        const uint64_t dims[5] = {32, (uint64_t)p->xw, (uint64_t)p->xh, (uint64_t)p->n, (uint64_t)(p->cin / 32)};
        const uint64_t strides[4] = {(uint64_t)p->cin * 4, (uint64_t)p->xw * p->cin * 4, (uint64_t)p->xh * p->xw * p->cin * 4, 128};
        const uint32_t box[5] = {32, 10, 6, 1, (uint32_t)(kS64NT / 32)};
        const uint32_t es[5] = {1, 1, 1, 1, 1};
        int rc = make_tmap_f32(&tx, p->x, 5, dims, strides, box, es, /*atom32=*/true);
        if (rc != SGV_OK) return rc;
    }
    SGV_OPT_IN_SMEM(wgrad_tf32_s64_kernel, kS64Total);
    dim3 grid((unsigned)ntiles, (unsigned)ksplit, 1);
    wgrad_tf32_s64_kernel<<<grid, kS64Threads, kS64Total, stream>>>(tg, tx, a);
    SGV_LAUNCH_OK("wgrad_tf32_s64_kernel");
    return SGV_OK;
}

} // namespace sgv
