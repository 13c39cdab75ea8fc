// Library plumbing: version, thread-local error text, device checks.
#include "hg_common.cuh"
#include "../../include/hg_api.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace hg {

static thread_local char g_err[512] = {0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return HG_OK;
    set_last_error("CUDA error %d (%s) at: %s", (int)e, cudaGetErrorString(e), what);
    return HG_ERR_CUDA;
}

bool pdl_enabled() {
    static const bool on = getenv("HG_NO_PDL") == nullptr;
    return on;
}

int num_sms() {
    static std::mutex mu;
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 1;
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 64 && cache[dev] > 0) return cache[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 1;
    if (dev < 64) cache[dev] = n;
    return n;
}

}  // namespace hg

extern "C" int hg_api_version(void) { return HG_API_VERSION; }

extern "C" size_t hg_last_error(char* buf, size_t cap) {
    const size_t n = strlen(hg::g_err);
    if (buf != nullptr && cap > 0) {
        const size_t m = n < cap - 1 ? n : cap - 1;
        memcpy(buf, hg::g_err, m);
        buf[m] = 0;
    }
    return n;
}

extern "C" int hg_check_device(void) {
    int dev = 0;
    HG_CUDA_OK(cudaGetDevice(&dev));
    int major = 0;
    HG_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
        hg::set_last_error("device %d has compute capability %d.x; this library is built for sm_100a only", dev, major);
        return HG_ERR_ARCH;
    }
    return HG_OK;
}
