// libsgv_b200 housekeeping: error text, device check, launch accounting.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

namespace sgv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

char* error_buffer() { return g_err; }

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device_slot()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}

int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

int num_sms()
{
    static PerDeviceInt cached;
    const int dev = current_device_slot();
    if (dev < 0) return 148;
    int n = cached.get(dev);
    if (n == 0)
    {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached.set(dev, n);
    }
    return n;
}

} // namespace sgv

extern "C" {

int sgv_abi_version(void) { return SGV_ABI_VERSION; }

const char* sgv_last_error(void) { return sgv::error_buffer(); }

int64_t sgv_kernel_launch_count(void) { return (int64_t)sgv::g_launches.load(); }

int sgv_device_check(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
        return sgv::fail(SGV_ERR_NO_DEVICE, "no CUDA device visible; libsgv_b200 has no CPU path");
    int dev = 0, major = 0;
    SGV_CUDA_OK(cudaGetDevice(&dev));
    SGV_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10)
        return sgv::fail(SGV_ERR_NO_DEVICE, "device %d has compute capability %d.x; libsgv_b200 is built for sm_100a only", dev, major);
    return SGV_OK;
}

} // extern "C"
