// Window gather + normalise: one coalesced strided-gather kernel replacing the reference's
// Python-loop make_windows/np.stack + NumPy normalize passes
// (4DOF/Scripts/06_test_full_pipeline.py:106-126; openLAB feature_utils.py:130-152 and
// 10_test_hybrid_pipeline.py:233-237,351; 1_DOF/Scripts/datasets.py:17-35).
//
// HBM-bound: writes 4*T*D bytes per window; at stride 1 the source rows are re-read T times but
// from L1/L2 (unique read traffic is ~4*D*stride bytes per window).  Output-driven mapping: each
// thread produces one 16-byte chunk of the contiguous [N,T,D] output (streaming st.global.cs), the
// (t,d) of its first element comes from one division, the other three advance incrementally.
#include "common.cuh"

namespace shm {

constexpr int WIN_THREADS = 256;
constexpr int WIN_PER_CTA = 8;

template <bool VEC4>
__global__ void __launch_bounds__(WIN_THREADS)
window_normalize_kernel(WinSrc src, const int* __restrict__ idx, long long N, float* __restrict__ out) {
    const int TD = src.T * src.D;
    const int chunks = VEC4 ? (TD >> 2) : TD;
    const long long groups = (N + WIN_PER_CTA - 1) / WIN_PER_CTA;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long n0 = g * WIN_PER_CTA;
        const int nw = (int)min((long long)WIN_PER_CTA, N - n0);
        for (int i = threadIdx.x; i < nw * chunks; i += WIN_THREADS) {
            const int wl = i / chunks;
            const int c = i - wl * chunks;
            const long long n = n0 + wl;
            const long long win = idx ? (long long)idx[n] : n;
            const float* wbase = src.base + win * src.win_stride;
            if (VEC4) {
                const int e0 = c << 2;
                int t = e0 / src.D;
                int d = e0 - t * src.D;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = win_transform(src, __ldg(wbase + (long long)t * src.row_stride + src.chan[d]), d);
                    if (++d == src.D) { d = 0; ++t; }
                }
                __stcs(reinterpret_cast<float4*>(out + n * TD) + c, make_float4(v[0], v[1], v[2], v[3]));
            } else {
                const int t = c / src.D;
                const int d = c - t * src.D;
                __stcs(out + n * TD + c, win_transform(src, __ldg(wbase + (long long)t * src.row_stride + src.chan[d]), d));
            }
        }
    }
}

}  // namespace shm

extern "C" int shm_window_normalize(const shm_window_src* src_host, const int32_t* idx, int64_t N, float* out,
                                    void* stream) {
    using namespace shm;
    if (N < 0 || (N > 0 && !out)) return SHM_ERR_ARG;
    WinSrc w;
    int rc = make_winsrc(src_host, &w);
    if (rc != SHM_OK) return rc;
    if (N == 0) return SHM_OK;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    if ((rc = check_device(dev)) != SHM_OK) return rc;
    const long long groups = (N + WIN_PER_CTA - 1) / WIN_PER_CTA;
    const int sms = device_sm_count(dev);
    const int grid = (int)min(groups, (long long)sms * 8 * 4);     // multiple of the SM count, 8 CTAs/SM resident
    const bool vec4 = ((w.T * w.D) % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec4) window_normalize_kernel<true><<<grid, WIN_THREADS, 0, st>>>(w, idx, N, out);
    else window_normalize_kernel<false><<<grid, WIN_THREADS, 0, st>>>(w, idx, N, out);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
