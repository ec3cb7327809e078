// Error plumbing and device queries shared by every entry point of
// libtfb_b200.so (C ABI in include/tfb.h).
#include "tfb_common.cuh"

#include <stdarg.h>
#include <string.h>

namespace tfb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

// 256 bytes of device memory per device for the result words of the set-up functions
// (allocated once, never freed: no cudaMalloc / cudaFree on their paths)
void *small_scratch() {
    static void *cached[64] = {nullptr};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!cached[dev] && cudaMalloc(&cached[dev], 256) != cudaSuccess) cached[dev] = nullptr;
    return cached[dev];
}

}  // namespace tfb

extern "C" int tfb_abi_version(void) { return TFB_ABI_VERSION; }

extern "C" const char *tfb_last_error(void) { return tfb::g_err; }

extern "C" int tfb_stream_synchronize(void *stream) {
    TFB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    return TFB_OK;
}

extern "C" int tfb_device_info(char *name, int name_len, int *sm_count, int *cc_major,
                               int *cc_minor) {
    int dev = 0, n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        tfb::set_error("no CUDA device visible");
        return TFB_ERR_NO_DEVICE;
    }
    TFB_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    TFB_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (name && name_len > 0) {
        strncpy(name, p.name, (size_t)name_len - 1);
        name[name_len - 1] = '\0';
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return TFB_OK;
}

extern "C" int64_t tfb_sizeof_channels_step_args(void) {
    return (int64_t)sizeof(tfb_channels_step_args);
}
extern "C" int64_t tfb_sizeof_channels_scalars(void) {
    return (int64_t)sizeof(tfb_channels_scalars);
}
