// Hardware-semantics probe for tcgen05.mma shared-memory descriptors (test tool, not part of the product library).
//
// Questions answered on a real B200 (results recorded in profiles/umma_probe_r1.txt and DESIGN.md):
//  Q1  K-major SWIZZLE_128B: may the descriptor start at a row that is NOT a multiple of 8 (1024 B)?  Is the XOR
//      pattern taken from absolute smem address bits [7:9] (then any row offset works with base_offset = 0), or
//      relative to the start (then base_offset must carry the phase)?
//  Q2  K-major: may SBO (distance between 8-row groups) be a non-multiple of 1024 B (e.g. 1280 = a 10-pixel pitch)?
//  Q3  MN-major SWIZZLE_128B with LBO = 4096 / SBO = 1024 (the wgrad staging layout): which rows are fetched?
//
// Method: B is an identity-like operand so that D[m][n] = A[m][k=n]; A holds, at logical (row i, col c), either the
// value i ("row probe") or c ("col probe"), both exact in TF32.  The smem image is written with generic stores using the
// absolute-address 128B swizzle (chunk ^= (addr >> 7) & 7), i.e. what TMA would produce in a 1024B-aligned buffer.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/_bin/umma_probe scripts/umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../stylegan_v_b200/csrc/ptx.cuh"

using namespace sgv::ptx;

struct Case
{
    int a_mn, b_mn;              // majorness
    unsigned a_off, b_off;       // byte offsets of the operand starts inside the smem image
    unsigned a_lbo, a_sbo, b_lbo, b_sbo;
    unsigned a_base_offset;      // 3-bit field of the A descriptor
    int n;                       // N of the MMA (multiple of 16, <= 64 here)
    int ksteps;                  // number of K=8 instructions
    unsigned a_kstep, b_kstep;   // byte advance of the start address per K=8 instruction
};

__device__ unsigned g_layout_type = 2;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B

__device__ uint64_t make_desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned base_offset)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)(g_layout_type & 7) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* image, int image_floats, Case c, float* out)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* sf = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < image_floats; i += blockDim.x) sf[i] = image[i];
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_slot, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0)
    {
        if (elect_one())
        {
            const uint32_t idesc = umma_idesc_tf32(128, c.n, c.a_mn, c.b_mn);
            const uint32_t base = smem_u32(smem);
            for (int k = 0; k < c.ksteps; k++)
            {
                const uint64_t da = make_desc(base + c.a_off + k * c.a_kstep, c.a_lbo, c.a_sbo, c.a_base_offset);
                const uint64_t db = make_desc(base + c.b_off + k * c.b_kstep, c.b_lbo, c.b_sbo, 0);
                mma_tf32(tmem, da, db, idesc, k > 0 ? 1u : 0u);
            }
            mma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[32];
    for (int cc = 0; cc < c.n / 32 + (c.n % 32 ? 1 : 0); cc++)
    {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; j++)
            if (cc * 32 + j < c.n) out[(warp * 32 + lane) * 64 + cc * 32 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// ---- host helpers: write logical K-major rows (128 B each) into the image with the absolute-address swizzle ----
static void put_row_kmajor(std::vector<float>& img, unsigned row_byte_addr, const float* row32)
{
    for (int ch = 0; ch < 8; ch++)
    {
        unsigned phys = ch ^ ((row_byte_addr >> 7) & 7);
        memcpy(&img[(row_byte_addr + phys * 16) / 4], row32 + ch * 4, 16);
    }
}

// 128B rows with the 32-byte-atom swizzle (Swizzle<2,5,2>): 32 B chunk index ^= (addr >> 7) & 3  (TMA: SWIZZLE_128B_ATOM_32B)
static void put_row_sw32(std::vector<float>& img, unsigned row_byte_addr, const float* row32)
{
    for (int ch = 0; ch < 4; ch++)
    {
        unsigned phys = ch ^ ((row_byte_addr >> 7) & 3);
        memcpy(&img[(row_byte_addr + phys * 32) / 4], row32 + ch * 8, 32);
    }
}

static bool run(const char* name, const std::vector<float>& img, const Case& c, std::vector<float>& out)
{
    float *dimg, *dout;
    cudaMalloc(&dimg, img.size() * 4); cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(dimg, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, 128 * 64 * 4);
    size_t smem = img.size() * 4 + 2048;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(dimg, (int)img.size(), c, dout);
    cudaError_t e = cudaDeviceSynchronize();
    out.resize(128 * 64);
    cudaMemcpy(out.data(), dout, 128 * 64 * 4, cudaMemcpyDeviceToHost);
    cudaFree(dimg); cudaFree(dout);
    if (e != cudaSuccess) { printf("%-58s CUDA ERROR %s\n", name, cudaGetErrorString(e)); return false; }
    return true;
}

int main()
{
    const unsigned A_BASE = 0, B_BASE = 96 * 1024;       // A region: 96 KB of K-major rows; B after it
    const int IMG = (96 + 16) * 1024 / 4;
    // ---------------- K-major probes ----------------
    for (int probe = 0; probe < 2; probe++)               // 0 = row probe, 1 = col probe
    {
        std::vector<float> img(IMG, -7777.f);
        for (unsigned r = 0; r < 96 * 1024 / 128; r++)
        {
            float row[32];
            for (int c = 0; c < 32; c++) row[c] = probe == 0 ? (float)r : (float)c;
            put_row_kmajor(img, A_BASE + r * 128, row);
        }
        // B: N=32 rows (n), K-major, B[n][k] = delta(n, k) -> D[m][n] = A[m][n]
        for (unsigned n = 0; n < 32; n++)
        {
            float row[32];
            for (int k = 0; k < 32; k++) row[k] = (k == (int)n) ? 1.f : 0.f;
            put_row_kmajor(img, B_BASE + n * 128, row);
        }
        struct { const char* name; unsigned row_off; unsigned sbo; unsigned bo; } tests[] = {
            {"K-major aligned start, SBO=1024 (sanity)", 0, 1024, 0},
            {"K-major start +1 row, SBO=1024, base_offset=0", 1, 1024, 0},
            {"K-major start +1 row, SBO=1024, base_offset=1", 1, 1024, 1},
            {"K-major start +3 rows, SBO=1024, base_offset=0", 3, 1024, 0},
            {"K-major start +3 rows, SBO=1024, base_offset=3", 3, 1024, 3},
            {"K-major start +10 rows, SBO=1024, base_offset=0", 10, 1024, 0},
            {"K-major aligned start, SBO=1280 (10-row pitch), bo=0", 0, 1280, 0},
            {"K-major start +11 rows, SBO=1280, bo=0", 11, 1280, 0},
            {"K-major start +11 rows, SBO=2304 (18-row pitch), bo=0", 11, 2304, 0},
        };
        for (auto& t : tests)
        {
            Case c{};
            c.a_mn = 0; c.b_mn = 0; c.a_off = A_BASE + t.row_off * 128; c.b_off = B_BASE;
            c.a_lbo = 16; c.a_sbo = t.sbo; c.b_lbo = 16; c.b_sbo = 1024; c.a_base_offset = t.bo;
            c.n = 32; c.ksteps = 4; c.a_kstep = 32; c.b_kstep = 32;
            std::vector<float> out;
            if (!run(t.name, img, c, out)) continue;
            // expectation under "absolute address swizzle": A row m = start_row + (m/8)*(sbo/128) + m%8 ; value = row (probe 0) or n (probe 1)
            int ok = 0, bad = 0; float first_bad_got = 0, first_bad_exp = 0; int fm = -1, fn = -1;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 32; n++)
                {
                    float exp = probe == 0 ? (float)(t.row_off + (m / 8) * (t.sbo / 128) + (m % 8)) : (float)n;
                    float got = out[m * 64 + n];
                    if (got == exp) ok++; else { if (!bad) { first_bad_got = got; first_bad_exp = exp; fm = m; fn = n; } bad++; }
                }
            printf("[%s probe] %-58s match=%d mismatch=%d", probe ? "col" : "row", t.name, ok, bad);
            if (bad) printf("  first: D[%d][%d]=%g expected %g | D[1][0..7]= %g %g %g %g %g %g %g %g", fm, fn, first_bad_got, first_bad_exp,
                            out[64 + 0], out[64 + 1], out[64 + 2], out[64 + 3], out[64 + 4], out[64 + 5], out[64 + 6], out[64 + 7]);
            printf("\n");
        }
    }
    // ---------------- MN-major probe (wgrad staging layout) ----------------
    // A image: 4 channel blocks x 32 pixel rows x 128 B ([blk][k=pixel][32 mn]); value at (m = blk*32 + j, k) is m (row probe) or k (col probe)
    for (int probe = 0; probe < 2; probe++)
    {
        std::vector<float> img(IMG, -7777.f);
        for (unsigned blk = 0; blk < 4; blk++)
            for (unsigned k = 0; k < 32; k++)
            {
                float row[32];
                for (int j = 0; j < 32; j++) row[j] = probe == 0 ? (float)(blk * 32 + j) : (float)k;
                put_row_kmajor(img, A_BASE + (blk * 32 + k) * 128, row);      // same physical row format (128 B rows, absolute swizzle)
            }
        // B MN-major: 1 block of 32 n, 32 k rows: B[k][n] = delta(n, k)  -> D[m][n] = A[m][k = n]
        for (unsigned k = 0; k < 32; k++)
        {
            float row[32];
            for (int n = 0; n < 32; n++) row[n] = (n == (int)k) ? 1.f : 0.f;
            put_row_kmajor(img, B_BASE + k * 128, row);
        }
        struct { const char* name; unsigned lbo, sbo; } tests[] = {
            {"MN-major A: LBO=4096 (mn blocks), SBO=1024 (k groups)", 4096, 1024},
            {"MN-major A: LBO=1024, SBO=4096 (fields swapped)", 1024, 4096},
        };
        for (auto& t : tests)
        {
            Case c{};
            c.a_mn = 1; c.b_mn = 1; c.a_off = A_BASE; c.b_off = B_BASE;
            c.a_lbo = t.lbo; c.a_sbo = t.sbo; c.b_lbo = t.lbo; c.b_sbo = t.sbo; c.a_base_offset = 0;
            c.n = 32; c.ksteps = 4; c.a_kstep = 1024; c.b_kstep = 1024;
            std::vector<float> out;
            if (!run(t.name, img, c, out)) continue;
            int ok = 0, bad = 0; float g0 = 0, e0 = 0; int fm = -1, fn = -1;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < 32; n++)
                {
                    float exp = probe == 0 ? (float)m : (float)n;
                    float got = out[m * 64 + n];
                    if (got == exp) ok++; else { if (!bad) { g0 = got; e0 = exp; fm = m; fn = n; } bad++; }
                }
            printf("[%s probe] %-58s match=%d mismatch=%d", probe ? "k  " : "m  ", t.name, ok, bad);
            if (bad) printf("  first: D[%d][%d]=%g expected %g | D[33][0..7]= %g %g %g %g %g %g %g %g", fm, fn, g0, e0,
                            out[33 * 64 + 0], out[33 * 64 + 1], out[33 * 64 + 2], out[33 * 64 + 3], out[33 * 64 + 4], out[33 * 64 + 5], out[33 * 64 + 6], out[33 * 64 + 7]);
            printf("\n");
        }
    }
    // ---------------- MN-major with SWIZZLE_128B_BASE32B (the only MN-major layout for 32-bit operands per CUTLASS) ----------------
    {
        unsigned lt = 1;
        cudaMemcpyToSymbol(g_layout_type, &lt, sizeof(lt));
        for (int probe = 0; probe < 2; probe++)
        {
            std::vector<float> img(IMG, -7777.f);
            for (unsigned blk = 0; blk < 4; blk++)
                for (unsigned k = 0; k < 32; k++)
                {
                    float row[32];
                    for (int j = 0; j < 32; j++) row[j] = probe == 0 ? (float)(blk * 32 + j) : (float)k;
                    put_row_sw32(img, A_BASE + (blk * 32 + k) * 128, row);
                }
            for (unsigned k = 0; k < 32; k++)
            {
                float row[32];
                for (int n = 0; n < 32; n++) row[n] = (n == (int)k) ? 1.f : 0.f;
                put_row_sw32(img, B_BASE + k * 128, row);
            }
            struct { const char* name; unsigned lbo, sbo; } tests[] = {
                {"MN-major BASE32B: LBO=4096 (mn blocks), SBO=512 (4-row k groups)", 4096, 512},
                {"MN-major BASE32B: LBO=512, SBO=4096 (fields swapped)", 512, 4096},
                {"MN-major BASE32B: LBO=4096, SBO=1024", 4096, 1024},
            };
            for (auto& t : tests)
            {
                Case c{};
                c.a_mn = 1; c.b_mn = 1; c.a_off = A_BASE; c.b_off = B_BASE;
                c.a_lbo = t.lbo; c.a_sbo = t.sbo; c.b_lbo = t.lbo; c.b_sbo = t.sbo; c.a_base_offset = 0;
                c.n = 32; c.ksteps = 4; c.a_kstep = 1024; c.b_kstep = 1024;
                std::vector<float> out;
                if (!run(t.name, img, c, out)) continue;
                int ok = 0, bad = 0; float g0 = 0, e0 = 0; int fm = -1, fn = -1;
                for (int m = 0; m < 128; m++)
                    for (int n = 0; n < 32; n++)
                    {
                        float exp = probe == 0 ? (float)m : (float)n;
                        float got = out[m * 64 + n];
                        if (got == exp) ok++; else { if (!bad) { g0 = got; e0 = exp; fm = m; fn = n; } bad++; }
                    }
                printf("[%s probe] %-66s match=%d mismatch=%d", probe ? "k  " : "m  ", t.name, ok, bad);
                if (bad) printf("  first: D[%d][%d]=%g expected %g | D[33][0..7]= %g %g %g %g %g %g %g %g", fm, fn, g0, e0,
                                out[33 * 64 + 0], out[33 * 64 + 1], out[33 * 64 + 2], out[33 * 64 + 3], out[33 * 64 + 4], out[33 * 64 + 5], out[33 * 64 + 6], out[33 * 64 + 7]);
                printf("\n");
            }
        }
    }
    // ---------------- MN-major BASE32B with unaligned starts / SBO != 512 (halo-patch addressing for wgrad) ----------------
    {
        // A image: one channel block, 200 pixel rows (k), value = k (row index); B: identity over n.  One K=8 instruction:
        // D[m][n] = sum_k A[k][m] * B[k][n]; with B[k][n] = delta(k - k0 == n) for the 8 rows used -> D[m][n] = A-row index fetched for k=n.
        for (int variant = 0; variant < 4; variant++)
        {
            const unsigned start_row = (unsigned[]){0, 3, 13, 13}[variant];
            const unsigned sbo = (unsigned[]){512, 512, 512, 768}[variant];      // 768 = 6-row pitch: rows {s..s+3} and {s+6..s+9}
            std::vector<float> img(IMG, -7777.f);
            for (unsigned k = 0; k < 200; k++)
            {
                float row[32];
                for (int j = 0; j < 32; j++) row[j] = (float)k;                     // every m sees the row index
                put_row_sw32(img, A_BASE + k * 128, row);
            }
            // B rows: 8 rows used by one instruction, aligned start; B[kk][n] = (n == kk)
            for (unsigned kk = 0; kk < 8; kk++)
            {
                float row[32];
                for (int n = 0; n < 32; n++) row[n] = (n == (int)kk) ? 1.f : 0.f;
                put_row_sw32(img, B_BASE + kk * 128, row);
            }
            Case c{};
            c.a_mn = 1; c.b_mn = 1; c.a_off = A_BASE + start_row * 128; c.b_off = B_BASE;
            c.a_lbo = 4096; c.a_sbo = sbo; c.b_lbo = 4096; c.b_sbo = 512; c.a_base_offset = 0;
            c.n = 32; c.ksteps = 1; c.a_kstep = 0; c.b_kstep = 0;
            std::vector<float> out;
            char name[128];
            snprintf(name, sizeof(name), "MN-major BASE32B start +%u rows, SBO=%u", start_row, sbo);
            if (!run(name, img, c, out)) continue;
            int ok = 0, bad = 0;
            for (int m = 0; m < 32; m++)
                for (int n = 0; n < 8; n++)
                {
                    float exp = (float)(start_row + (n / 4) * (sbo / 128) + (n % 4));
                    if (out[m * 64 + n] == exp) ok++; else bad++;
                }
            printf("[k-row probe] %-58s match=%d mismatch=%d | D[0][0..7]= %g %g %g %g %g %g %g %g\n", name, ok, bad,
                   out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
        }
    }
    return 0;
}
