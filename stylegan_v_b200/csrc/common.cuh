// Shared host/device helpers for libsgv_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/sgv_b200.h"

namespace sgv {

// thread-local error text returned by sgv_last_error()
char* error_buffer();
int fail(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define SGV_CHECK_ARG(cond, ...) do { if (!(cond)) return ::sgv::fail(SGV_ERR_INVALID, __VA_ARGS__); } while (0)
#define SGV_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) \
    return ::sgv::fail(SGV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); } while (0)

// Launch-error check that does not synchronise.
#define SGV_LAUNCH_OK(name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
    return ::sgv::fail(SGV_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e__)); ::sgv::count_launch(); } while (0)

int num_sms();   // SM count of the current device (cached per device)
int current_device_slot();                       // cudaGetDevice() clamped to [0, kMaxDevices); -1 on failure
int env_int(const char* name, int dflt);         // integer tuning switch from the environment (read by callers ONCE, into a static const)
constexpr int kMaxDevices = 64;

// Kernel attributes (cudaFuncSetAttribute) and occupancy answers belong to a device context, and entry points are called from several
// threads (forward thread + autograd workers): cache them per device in zero-initialised atomics.
struct PerDeviceInt
{
    std::atomic<int> v[kMaxDevices];
    int get(int dev) const { return (dev >= 0 && dev < kMaxDevices) ? v[dev].load(std::memory_order_acquire) : 0; }
    void set(int dev, int x) { if (dev >= 0 && dev < kMaxDevices) v[dev].store(x, std::memory_order_release); }
};

// opt a kernel into more than 48 KB of dynamic shared memory, once per device
#define SGV_OPT_IN_SMEM(kern, bytes) do { static ::sgv::PerDeviceInt done__; const int dev__ = ::sgv::current_device_slot(); \
    if (!done__.get(dev__)) { SGV_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); done__.set(dev__, 1); } } while (0)

// Ablation switches (skip staging / epilogue / MMAs: wrong results, timing only) exist only in builds with -DSGV_ABLATION; release
// kernels carry no such branches.
#ifdef SGV_ABLATION
#define SGV_ABL(flags, bit) (((flags) & (bit)) != 0)
#else
#define SGV_ABL(flags, bit) false
#endif

template <class T> struct acc_type            { typedef float  type; };
template <>        struct acc_type<double>    { typedef double type; };

// Same floor division as the reference kernels (upfirdn2d.cu:20-24); exact for b > 0 and any a.
__host__ __device__ __forceinline__ int floor_div(int a, int b)
{
    int t = 1 - a / b;
    return (a + t * b) / b - t;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Division of 0 <= n < 2^31 by a runtime-constant d >= 1 with one mulhi + add + shift (Granlund-Montgomery).
struct FastDiv
{
    uint32_t d, m, s;
    __host__ __device__ FastDiv() : d(1), m(1), s(0) {}
    __host__ __device__ explicit FastDiv(uint32_t div) : d(div)
    {
        s = 0;
        while ((1ull << s) < div) s++;
        m = (uint32_t)(((1ull << 32) * ((1ull << s) - div)) / div + 1);
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return (__umulhi(n, m) + n) >> s; }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const { q = div(n); r = n - q * d; }
};

} // namespace sgv
