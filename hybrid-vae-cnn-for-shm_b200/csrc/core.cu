// Error strings, device checks, small shared host helpers of libshmfast.
#include "common.cuh"

namespace shm {

static thread_local char g_cuda_err[512] = "";

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}

int device_sm_count(int device) {
    static int cached[64] = {0};
    if (device >= 0 && device < 64 && cached[device]) return cached[device];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n <= 0) n = 148;
    if (device >= 0 && device < 64) cached[device] = n;
    return n;
}

int check_device(int device) {
    static int cached[64] = {0};      // 1 = ok, 2 = bad
    if (device >= 0 && device < 64 && cached[device]) return cached[device] == 1 ? SHM_OK : SHM_ERR_DEVICE;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return SHM_ERR_DEVICE;
    }
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device) != cudaSuccess) return SHM_ERR_DEVICE;
    const bool ok = (major == 10);
    if (device < 64) cached[device] = ok ? 1 : 2;
    return ok ? SHM_OK : SHM_ERR_DEVICE;
}

}  // namespace shm

extern "C" const char* shm_strerror(int code) {
    switch (code) {
        case SHM_OK: return "ok";
        case SHM_ERR_ARG: return "invalid argument (null pointer, negative size or inconsistent shapes)";
        case SHM_ERR_UNSUPPORTED: return "unsupported configuration (see shm_vae_cfg in shmfast.h)";
        case SHM_ERR_CUDA: return "CUDA runtime error (see shm_last_cuda_error)";
        case SHM_ERR_DEVICE: return "no CUDA device of compute capability 10.x (B200); libshmfast has no CPU fallback";
        case SHM_ERR_NOMEM: return "out of memory";
        default: return "unknown error";
    }
}

extern "C" const char* shm_last_cuda_error(void) { return shm::g_cuda_err; }
extern "C" int shm_version(void) { return 100; }
extern "C" int shm_device_check(int device) { return shm::check_device(device); }
