// Rate probe for tcgen05.mma kind::tf32 with both operands in shared memory (test tool, not part of the product library).
//
// Question: how many SM cycles does one 128 x N x 8 TF32 MMA (K-major SWIZZLE_128B operands, SS mode) occupy when issued
// back to back, for N = 64 / 128 / 256, alone and while four other warps stream LDS.128 + STS.128 over a separate 42 KB
// region (what the conv kernel's transform warps do)?  This bounds the implicit-GEMM kernels from above.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/_bin/mma_rate_probe scripts/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "../stylegan_v_b200/csrc/ptx.cuh"

using namespace sgv::ptx;

template <int N>
__global__ void __launch_bounds__(192, 1) rate_kernel(int iters, int ls_traffic, int distinct_b, long long* out_cycles)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop;
    // A: 42 KB "patch" at 0, B: up to 4 slabs of N rows at 48 KB.., LDS/STS scratch at 176 KB
    for (int i = threadIdx.x; i < (48 * 1024 + 4 * N * 128) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0)
    {
        if (elect_one())
        {
            constexpr uint32_t idesc = umma_idesc_tf32(128, N);
            const uint32_t base = smem_u32(smem);
            const long long t0 = clock64();
            for (int i = 0; i < iters; i++)
            {
                // one "tap": 2 halves x 4 k-steps, like the conv kernel
                const uint64_t db = umma_desc_k_sw128(base + 48 * 1024 + (distinct_b ? (i & 3) * N * 128 : 0));
#pragma unroll
                for (int h = 0; h < 2; h++)
                {
                    const uint64_t da = umma_desc_k_sw128(base + (uint32_t)((i % 9) * 3 + 8 * h) * 128u);
#pragma unroll
                    for (int k = 0; k < 4; k++) mma_tf32(tmem + (uint32_t)(h * N), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, 1u);
                }
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            out_cycles[blockIdx.x] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    }
    else if (warp >= 2 && ls_traffic)
    {
        const uint32_t scratch = smem_u32(smem) + 176 * 1024 + (uint32_t)(threadIdx.x - 64) * 16u;
        while (!stop)
        {
#pragma unroll
            for (int j = 0; j < 16; j++)
            {
                float4 v = lds128(scratch + (uint32_t)(j * 2048));
                v.x *= 1.0001f; v.y *= 1.0001f;
                sts128(scratch + (uint32_t)(j * 2048), v);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N>
static void run(int iters, int ls, int distinct_b)
{
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int smem = 220 * 1024;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) rate_kernel<N><<<148, 192, smem>>>(iters, ls, distinct_b, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); exit(1); }
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (int i = 0; i < 148; i++) { avg += (double)h[i]; if (h[i] > mx) mx = h[i]; }
    avg /= 148;
    const double per = avg / ((double)iters * 8);
    printf("N=%3d ls_traffic=%d distinct_b=%d : %.1f clk per 128xNx8 MMA (max-SM %.1f)  => %.0f TFLOP/s at 1.92 GHz x 148 SMs\n", N, ls, distinct_b, per,
           (double)mx / ((double)iters * 8), 128.0 * N * 8 * 2 / per * 1.92e9 * 148 / 1e12);
    cudaFree(d);
}

int main()
{
    const int iters = 4096;
    for (int ls = 0; ls < 2; ls++)
        for (int db = 0; db < 2; db++) { run<64>(iters, ls, db); run<128>(iters, ls, db); run<256>(iters, ls, db); }
    return 0;
}
