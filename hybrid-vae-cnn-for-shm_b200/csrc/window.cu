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

// Tile kernel for sources whose windows are whole rows apart (a series [rows, d_all] with an integer row stride, or
// materialised windows): a CTA stages the rows its group of windows touches into shared memory ONCE -- channel select,
// normalise, clip, NaN policy applied once per source element instead of once per output element (T/stride times fewer
// IEEE divisions at stride 1) -- and every window then is a plain shared -> global copy (one LDS.128 + one streaming
// STG.128 per 16 output bytes, one warp per window, no index divisions).  Same arithmetic (win_transform), hence the
// same bits, as the generic kernel.
constexpr int WT_THREADS = 256;
constexpr int WT_SMEM_FLOATS = 11 * 1024;          // largest tile: 44 KB of dynamic shared memory (the launch asks for what its group needs)
constexpr int WT_TARGET_FLOATS = 6 * 1024;         // series groups are sized to <= 24 KB when that still holds >= 32 windows: 8 CTAs per SM

__global__ void __launch_bounds__(WT_THREADS)
window_tile_kernel(WinSrc src, const int* __restrict__ idx, long long N, float* __restrict__ out, int group, int step_rows) {
    extern __shared__ __align__(16) float tile[];
    __shared__ int s_chan[SHM_MAX_D];
    const int T = src.T, D = src.D, TD = T * D;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < D) s_chan[tid] = src.chan[tid];
    const long long groups = (N + group - 1) / group;
    const int slab = (idx == nullptr) ? step_rows * D : TD;         // floats between consecutive windows in the tile
    const bool vec = (slab % 4 == 0) && (TD % 4 == 0);
    const int r_t = tid / D, d_t = tid - r_t * D;                    // staging: element tid + k * WT_THREADS = (row, channel)
    const int dq_s = WT_THREADS / D, dr_s = WT_THREADS - dq_s * D;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long n0 = g * group;
        const int nw = (int)min((long long)group, N - n0);
        __syncthreads();
        if (idx == nullptr) {
            // rows [n0*step, n0*step + (nw-1)*step + T) of the source, shared by the windows of the group
            const int rows = (nw - 1) * step_rows + T;
            const float* base = src.base + n0 * src.win_stride;
            // (row, channel) stepped incrementally: at stride 20 the staged rows are 11 % of the output, and an integer division
            // per staged element made this phase 2/3 of the kernel's instructions (openLAB gather: 68 % of the copy bandwidth)
            int r = r_t, d = d_t;
            for (int i = tid; i < rows * D; i += WT_THREADS) {
                tile[i] = win_transform(src, __ldg(base + (long long)r * src.row_stride + s_chan[d]), d);
                r += dq_s; d += dr_s;
                if (d >= D) { d -= D; ++r; }
            }
        } else {
            for (int w = 0; w < nw; ++w) {                           // gathered windows: one private slab per window
                const float* base = src.base + (long long)idx[n0 + w] * src.win_stride;
                int r = r_t, d = d_t;
                for (int e = tid; e < TD; e += WT_THREADS) {
                    tile[w * TD + e] = win_transform(src, __ldg(base + (long long)r * src.row_stride + s_chan[d]), d);
                    r += dq_s; d += dr_s;
                    if (d >= D) { d -= D; ++r; }
                }
            }
        }
        __syncthreads();
        if (vec) {
            // the group's windows are one contiguous run of the output: walk it with all threads (no idle lanes on the last
            // 16-byte chunk of a window), stepping (window, chunk) incrementally instead of dividing
            const int c4 = TD >> 2, slab4 = slab >> 2;
            const float4* s4 = reinterpret_cast<const float4*>(tile);
            float4* o4 = reinterpret_cast<float4*>(out + n0 * TD);
            const int dq = WT_THREADS / c4, dr = WT_THREADS - dq * c4;
            int w = tid / c4, c = tid - w * c4;
            for (int i = tid; i < nw * c4; i += WT_THREADS) {
                __stcs(o4 + i, s4[w * slab4 + c]);
                w += dq; c += dr;
                if (c >= c4) { c -= c4; ++w; }
            }
        } else {
            for (int w = warp; w < nw; w += WT_THREADS / 32)
                for (int e = lane; e < TD; e += 32) __stcs(out + (n0 + w) * TD + e, tile[w * slab + e]);
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
    const int sms = device_sm_count(dev);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool out16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    // tile path: windows a whole number of rows apart (series or materialised windows) whose T*D slab fits the tile
    const int TD = w.T * w.D;
    if (out16 && TD <= WT_SMEM_FLOATS && w.row_stride > 0 && w.win_stride > 0 && w.win_stride % w.row_stride == 0 &&
        w.win_stride / w.row_stride < (1 << 20)) {
        const int step_rows = (int)(w.win_stride / w.row_stride);
        int group, tile_floats;
        if (idx == nullptr) {                                       // consecutive windows share the staged rows
            auto fit = [&](int floats) { const int g = (floats / w.D - w.T) / step_rows + 1; return g > 128 ? 128 : g; };
            group = fit(WT_SMEM_FLOATS);
            // a smaller tile = more resident CTAs, so one CTA's staging phase hides under the others' copy phase (openLAB,
            // stride 20: 128-window groups in 44 KB tiles ran 5 CTAs per SM at 68 % of the copy bandwidth)
            if (TD <= WT_TARGET_FLOATS && fit(WT_TARGET_FLOATS) >= 32) group = fit(WT_TARGET_FLOATS);
            tile_floats = ((group - 1) * step_rows + w.T) * w.D;
        } else {                                                    // gathered windows: one slab each
            group = WT_SMEM_FLOATS / TD;
            group = group > 16 ? 16 : group;
            tile_floats = group * TD;
        }
        const long long tgroups = (N + group - 1) / group;
        const int tgrid = (int)min(tgroups, (long long)sms * 16);
        window_tile_kernel<<<tgrid, WT_THREADS, (size_t)tile_floats * sizeof(float), st>>>(w, idx, N, out, group, step_rows);
        SHM_LAUNCH_CHECK();
        return SHM_OK;
    }
    const long long groups = (N + WIN_PER_CTA - 1) / WIN_PER_CTA;
    const int grid = (int)min(groups, (long long)sms * 8 * 4);     // multiple of the SM count, 8 CTAs/SM resident
    const bool vec4 = ((w.T * w.D) % 4 == 0) && out16;
    if (vec4) window_normalize_kernel<true><<<grid, WIN_THREADS, 0, st>>>(w, idx, N, out);
    else window_normalize_kernel<false><<<grid, WIN_THREADS, 0, st>>>(w, idx, N, out);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
