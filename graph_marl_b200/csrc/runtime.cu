// runtime.cu -- error reporting, device check, launch accounting.
#include <stdarg.h>

#include "common.cuh"

namespace gm {
static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gm

extern "C" {
const char* gm_last_error(void) { return gm::g_err; }
int gm_abi_version(void) { return 4; }
int64_t gm_kernel_launch_count(void) { return gm::g_launches.load(); }

int gm_device_check(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        gm::set_error("no CUDA device (%s); libgraphmarl_b200 has no CPU fallback",
                      e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        cudaGetLastError();
        return GM_ERR_NO_DEVICE;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        gm::set_error("device compute capability %d.x, built for sm_100a only", major);
        return GM_ERR_NO_DEVICE;
    }
    return GM_OK;
}
}
