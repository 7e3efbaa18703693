// Diagnostics entry points of the C ABI.
#include "common.cuh"

#include <atomic>

namespace lens {

static thread_local char g_err[512] = "";

char *err_buf() { return g_err; }

void set_err(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

}  // namespace lens

extern "C" int lens_version(int *major, int *minor)
{
    if (major) *major = LENS_B200_VERSION_MAJOR;
    if (minor) *minor = LENS_B200_VERSION_MINOR;
    return 0;
}

extern "C" const char *lens_last_error(void) { return lens::err_buf(); }

extern "C" int lens_device_sm_count(int *n_sm)
{
    LENS_CHECK_ARG(n_sm != nullptr, "lens_device_sm_count: n_sm is NULL");
    int n = lens::sm_count();
    if (n <= 0) {
        lens::set_err("lens_device_sm_count: no CUDA device");
        return (int)cudaErrorNoDevice;
    }
    *n_sm = n;
    return 0;
}

namespace lens { long long launches(); }
extern "C" int lens_launch_count(int64_t *n_launches)
{
    LENS_CHECK_ARG(n_launches != nullptr, "lens_launch_count: NULL argument");
    *n_launches = lens::launches();
    return 0;
}
