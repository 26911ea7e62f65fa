// Exact device-side percentile with NumPy's default ("linear") interpolation, reproducing
// `np.percentile(scores_fp32, q)` bit for bit: threshold calibration of
// 4DOF/Scripts/04_vae_thresholding.py:283 (P99) and openLAB 05_validate_vae.py:253 (P95).
//
// NumPy (2.x) evaluates everything in the array's dtype: qf = fp32(q)/fp32(100),
// v = fp32(N-1)*qf, lo = floor(v), g = v - lo, r = a + (b-a)*g, and for g >= 0.5 r = b - (b-a)*(1-g),
// with a, b the lo-th and (lo+1)-th order statistics.  Those two are found by a most-significant-digit
// radix select (4 passes of 8 bits over an order-preserving key), both ranks in the same pass, so the
// scores never leave the device and are never fully sorted.
#include "common.cuh"

namespace shm {

struct PctState {
    unsigned int hist[2][256];
    unsigned int prefix[2];
    unsigned int rank[2];
    unsigned int lo_is_last;
    float g;
};

__device__ __forceinline__ unsigned int f2key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
    const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__global__ void pct_init_kernel(PctState* s, long long N, float q) {
    if (threadIdx.x < 256) { s->hist[0][threadIdx.x] = 0; s->hist[1][threadIdx.x] = 0; }
    if (threadIdx.x == 0) {
        const float qf = __fdiv_rn(q, 100.0f);
        const float v = __fmul_rn((float)(N - 1), qf);
        long long lo;
        float g;
        if (v >= (float)(N - 1)) { lo = N - 1; g = 0.f; s->lo_is_last = 1; }
        else { lo = (long long)floorf(v); g = __fsub_rn(v, (float)lo); s->lo_is_last = 0; }
        s->rank[0] = (unsigned int)lo;
        s->rank[1] = (unsigned int)min(lo + 1, N - 1);
        s->prefix[0] = 0; s->prefix[1] = 0;
        s->g = g;
    }
}

__global__ void __launch_bounds__(256)
pct_hist_kernel(const float* __restrict__ x, long long N, PctState* s, int pass) {
    __shared__ unsigned int h[2][256];
    h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0;
    __syncthreads();
    const int shift = 24 - 8 * pass;
    const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    const unsigned int p0 = s->prefix[0], p1 = s->prefix[1];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int k = f2key(x[i]);
        const unsigned int d = (k >> shift) & 255u;
        if (((k ^ p0) & himask) == 0) atomicAdd(&h[0][d], 1u);
        if (((k ^ p1) & himask) == 0) atomicAdd(&h[1][d], 1u);
    }
    __syncthreads();
    if (h[0][threadIdx.x]) atomicAdd(&s->hist[0][threadIdx.x], h[0][threadIdx.x]);
    if (h[1][threadIdx.x]) atomicAdd(&s->hist[1][threadIdx.x], h[1][threadIdx.x]);
}

__global__ void pct_pick_kernel(PctState* s, int pass, double* result) {
    const int shift = 24 - 8 * pass;
    const int r = threadIdx.x;            // 2 threads
    if (r < 2) {
        unsigned int rank = s->rank[r], cum = 0;
        int d = 0;
        for (; d < 256; ++d) {
            const unsigned int c = s->hist[r][d];
            if (rank < cum + c) break;
            cum += c;
        }
        if (d > 255) d = 255;
        s->prefix[r] |= (unsigned int)d << shift;
        s->rank[r] = rank - cum;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&s->hist[0][0])[i] = 0;
    if (pass == 3 && threadIdx.x == 0) {
        const float a = key2f(s->prefix[0]), b = key2f(s->prefix[1]);
        const float g = s->g;
        const float diff = __fsub_rn(b, a);
        float res = __fadd_rn(a, __fmul_rn(diff, g));
        if (g >= 0.5f) res = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
        *result = (double)res;
    }
}

}  // namespace shm

extern "C" int64_t shm_percentile_workspace_bytes(int64_t) { return (int64_t)sizeof(shm::PctState); }

extern "C" int shm_percentile(const float* scores, int64_t N, double q, double* result, void* workspace, void* stream) {
    using namespace shm;
    if (!scores || !result || !workspace || N <= 0 || N > 0xffffffffLL || !(q >= 0.0 && q <= 100.0)) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PctState* s = static_cast<PctState*>(workspace);
    pct_init_kernel<<<1, 256, 0, st>>>(s, N, (float)q);
    SHM_LAUNCH_CHECK();
    const int grid = (int)min((long long)device_sm_count(dev) * 8, (long long)((N + 255) / 256));
    for (int pass = 0; pass < 4; ++pass) {
        pct_hist_kernel<<<grid, 256, 0, st>>>(scores, N, s, pass);
        SHM_LAUNCH_CHECK();
        pct_pick_kernel<<<1, 64, 0, st>>>(s, pass, result);
        SHM_LAUNCH_CHECK();
    }
    return SHM_OK;
}
