// openLAB extraction front-end on the device: one run's parsed catman columns ->
// obstruction sentinel, provider outlier masks, AND-rule cleaning + interpolation + moving average, finite-DMS row
// compaction, and the per-window metadata / rule labels of
// 20250506_openLAB_tests/Codes/01_extract_windows_and_labels.py:104-236 (helpers :58-83; feature_utils.py:49-99,130-177).
// The kept series stay on the device ([rows_kept, 4] float32, clean and raw): the scorer and the CNN read their windows
// from them as strided views (shm_window_src), so nothing between the text parser and the labels touches the host.
//
// The reference's cleaning loop is sequential in form only: once a sample is removed x2[i] is NaN, which removes sample
// i+1 ("else" branch, feature_utils.py:90-92), so everything from the FIRST invalid sample or AND-rule hit onwards goes.
// That first index is an atomicMin; pandas' interpolate(limit_direction="both") then holds the last valid value, and
// np.convolve(mode="same") with the flat kernel is a zero-padded centred mean accumulated in ascending order (fp64).
// All outputs are bit-identical to the reference on its own data (tests/golden/openlab_frontend.npz).
#include "common.cuh"

namespace shm {

struct ExtractWs {
    int* i0;            // [3] first removed sample per displacement channel
    int* keep_idx;      // [R] kept row ids (ascending)
    int* keep_count;    // [1]
    float* masks;       // [3][R] per-row (outlier, invalid, removed) masks before compaction
    float* mk;          // [3][R] the same after compaction
    float* keepf;       // [R] 1.0 where DMS is finite
    void* compact_ws;
};

__device__ __forceinline__ float sentinel_f(float v, float sentinel) { return (v <= sentinel) ? __int_as_float(0x7fc00000) : v; }

// pass 1: first trigger of the cleaning rule per channel; provider outlier / invalid masks; keep flags
__global__ void extract_scan_kernel(const float* __restrict__ raw, int R, shm_openlab_extract_cfg cfg, int* __restrict__ i0,
                                    float* __restrict__ masks, float* __restrict__ keepf) {
    const float sent = (float)cfg.obstruction_sentinel;
    const float dth = (float)cfg.raw_diff_th, ath = (float)cfg.raw_abs_th;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R; i += gridDim.x * blockDim.x) {
        float out_any = 0.f, inv_any = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float u = sentinel_f(raw[(size_t)i * 4 + 1 + c], sent);
            const bool fin = isfinite(u);
            bool outl = !fin;                                                        // 01:74
            bool trig = !fin;                                                        // feature_utils.py:78-80
            if (i > 0) {
                const float up = sentinel_f(raw[(size_t)(i - 1) * 4 + 1 + c], sent);
                // provider mask: float32 arithmetic, >= thresholds (01:77-79)
                const float du = fabsf(__fsub_rn(u, up));
                outl = outl || (du >= dth && fabsf(u) >= ath);
                // cleaning rule: fp64 arithmetic, strict thresholds, both samples finite (feature_utils.py:86-89)
                if (fin && isfinite(up)) trig = fabs((double)u - (double)up) > cfg.clean_max_jump && fabs((double)u) > cfg.clean_max_abs;
            }
            if (trig) atomicMin(i0 + c, i);
            out_any = fmaxf(out_any, outl ? 1.f : 0.f);
            inv_any = fmaxf(inv_any, fin ? 0.f : 1.f);
        }
        masks[i] = out_any;
        masks[(size_t)R + i] = inv_any;
        keepf[i] = isfinite(raw[(size_t)i * 4]) ? 1.f : 0.f;                         // 01:151
    }
}

// cleaned value of channel c at row i: hold-last-valid interpolation + zero-padded centred moving average, fp64 -> fp32
__device__ __forceinline__ float cleaned_at(const float* __restrict__ raw, int R, int c, int i, int i0, float sent, int w) {
    const double hold = (i0 > 0) ? (double)sentinel_f(raw[(size_t)(i0 - 1) * 4 + 1 + c], sent) : (double)__int_as_float(0x7fc00000);
    auto xi = [&](int j) -> double {
        if (j < 0 || j >= R) return 0.0;
        return (j < i0) ? (double)sentinel_f(raw[(size_t)j * 4 + 1 + c], sent) : hold;
    };
    if (w <= 1) return (float)xi(i);
    const int half = w / 2;
    const double k = 1.0 / (double)w;
    double acc = 0.0;
    for (int j = 0; j < w; ++j) acc = __dadd_rn(acc, __dmul_rn(xi(i - half + j), k));   // ascending order, multiply then add
    return (float)acc;
}

// pass 2: gather the kept rows: A_clean / A_raw [rows_kept, 4] and the three row masks
__global__ void extract_gather_kernel(const float* __restrict__ raw, int R, shm_openlab_extract_cfg cfg, const int* __restrict__ i0,
                                      const int* __restrict__ keep_idx, const int* __restrict__ keep_count,
                                      const float* __restrict__ masks, float* __restrict__ a_clean, float* __restrict__ a_raw,
                                      float* __restrict__ mk) {
    const float sent = (float)cfg.obstruction_sentinel;
    const int n = *keep_count;
    const int f0 = min(i0[0], R), f1 = min(i0[1], R), f2 = min(i0[2], R);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = keep_idx[j];
        const float dms = raw[(size_t)i * 4];
        const float4 rw = make_float4(dms, sentinel_f(raw[(size_t)i * 4 + 1], sent), sentinel_f(raw[(size_t)i * 4 + 2], sent),
                                      sentinel_f(raw[(size_t)i * 4 + 3], sent));
        const float4 cl = make_float4(dms, cleaned_at(raw, R, 0, i, f0, sent, cfg.ma_window), cleaned_at(raw, R, 1, i, f1, sent, cfg.ma_window),
                                      cleaned_at(raw, R, 2, i, f2, sent, cfg.ma_window));
        reinterpret_cast<float4*>(a_raw)[j] = rw;
        reinterpret_cast<float4*>(a_clean)[j] = cl;
        mk[j] = masks[i];                                                            // outlier
        mk[(size_t)R + j] = masks[(size_t)R + i];                                    // invalid
        mk[(size_t)2 * R + j] = (i >= f0 || i >= f1 || i >= f2) ? 1.f : 0.f;         // removed by the cleaning rule
    }
}

// pass 3: one warp per window: mask ratios, structural envelope, flatline proxy, labels (01:170-214)
__global__ void extract_window_kernel(const float* __restrict__ a_clean, const float* __restrict__ mk, int R,
                                      const int* __restrict__ keep_count, shm_openlab_extract_cfg cfg, int* __restrict__ n_windows,
                                      int* __restrict__ label, float* __restrict__ u_min, float* __restrict__ u_max,
                                      float* __restrict__ dms_range, float* __restrict__ inv_ratio, float* __restrict__ out_ratio,
                                      float* __restrict__ rem_ratio, int* __restrict__ flatline, int* __restrict__ all_nan) {
    const int n = *keep_count, T = cfg.T, S = cfg.stride;
    const int nW = (n < T) ? 0 : (n - T) / S + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) *n_windows = nW;
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    const float NANF = __int_as_float(0x7fc00000);
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nW; w += warps) {
        const int r0 = w * S;
        float s_out = 0.f, s_inv = 0.f, s_rem = 0.f;
        float mn = INFINITY, mx = -INFINITY, dmn = INFINITY, dmx = -INFINITY;
        double sum = 0.0;
        int cnt = 0;
        for (int t = lane; t < T; t += 32) {
            const int r = r0 + t;
            s_out += mk[r]; s_inv += mk[(size_t)R + r]; s_rem += mk[(size_t)2 * R + r];
            const float4 c = reinterpret_cast<const float4*>(a_clean)[r];
            if (!isnan(c.x)) { dmn = fminf(dmn, c.x); dmx = fmaxf(dmx, c.x); }
            const float uv[3] = {c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if ((cfg.struct_channel_mask >> k) & 1) {
                    const float v = uv[k];
                    if (!isnan(v)) { mn = fminf(mn, v); mx = fmaxf(mx, v); sum += (double)v; ++cnt; }
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_out += __shfl_xor_sync(0xffffffffu, s_out, o); s_inv += __shfl_xor_sync(0xffffffffu, s_inv, o);
            s_rem += __shfl_xor_sync(0xffffffffu, s_rem, o);
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            dmn = fminf(dmn, __shfl_xor_sync(0xffffffffu, dmn, o)); dmx = fmaxf(dmx, __shfl_xor_sync(0xffffffffu, dmx, o));
            sum += __shfl_xor_sync(0xffffffffu, sum, o); cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        }
        // variance over the finite structural samples (np.nanvar), fp64
        const double mean = cnt ? sum / (double)cnt : 0.0;
        double ss = 0.0;
        for (int t = lane; t < T; t += 32) {
            const float4 c = reinterpret_cast<const float4*>(a_clean)[r0 + t];
            const float uv[3] = {c.y, c.z, c.w};
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if ((cfg.struct_channel_mask >> k) & 1) {
                    const float v = uv[k];
                    if (!isnan(v)) { const double d = (double)v - mean; ss += d * d; }
                }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        if (lane == 0) {
            const float fT = (float)T;
            const float r_out = __fdiv_rn(s_out, fT), r_inv = __fdiv_rn(s_inv, fT), r_rem = __fdiv_rn(s_rem, fT);   // float32 mean of 0/1 masks
            const float umin = cnt ? mn : NANF, umax = cnt ? mx : NANF;       // np.nanmin / nanmax of an all-NaN slice is NaN
            const bool alln = !isfinite(umin) || !isfinite(umax);
            const float drng = (dmx >= dmn) ? __fsub_rn(dmx, dmn) : NANF;
            const float uvar = cnt ? (float)(ss / (double)cnt) : NANF;
            const int flat = (uvar < cfg.flat_var_eps && drng > cfg.force_range_for_flatline) ? 1 : 0;
            const bool sensor = (r_inv >= cfg.raw_invalid_ratio_fault) || (r_out > 0.f) || (r_rem > 0.f) || flat || alln;
            const bool structural = umax > cfg.allow_max;
            label[w] = sensor ? 1 : (structural ? 2 : 0);                     // strict precedence SF > ST > Normal (01:212-214)
            u_min[w] = umin; u_max[w] = umax; dms_range[w] = drng;
            inv_ratio[w] = r_inv; out_ratio[w] = r_out; rem_ratio[w] = r_rem;
            flatline[w] = flat; all_nan[w] = alln ? 1 : 0;
        }
    }
}

__global__ void extract_init_kernel(int* i0, int R) { if (threadIdx.x < 3) i0[threadIdx.x] = R; }

}  // namespace shm

using namespace shm;

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" int64_t shm_openlab_extract_workspace_bytes(int64_t R) {
    if (R < 0) return 0;
    const size_t r = (size_t)(R > 0 ? R : 1);
    return (int64_t)(256 + align256(r * 4) + 256 + 2 * align256(3 * r * 4) + align256(r * 4) +
                     align256((size_t)shm_compact_workspace_bytes((int64_t)r)) + 256);
}

extern "C" int shm_openlab_extract(const float* raw, int64_t R, const shm_openlab_extract_cfg* cfg_host, float* a_clean,
                                   float* a_raw, int32_t* rows_kept, int32_t* n_windows, int32_t* label, float* u_min,
                                   float* u_max, float* dms_range, float* raw_invalid_ratio, float* raw_outlier_ratio,
                                   float* removed_ratio, int32_t* flatline_loadaware, int32_t* all_nan_struct, void* workspace,
                                   void* stream) {
    if (!raw || !cfg_host || !a_clean || !a_raw || !rows_kept || !n_windows || !label || !u_min || !u_max || !dms_range ||
        !raw_invalid_ratio || !raw_outlier_ratio || !removed_ratio || !flatline_loadaware || !all_nan_struct || !workspace)
        return SHM_ERR_ARG;
    if (R < 1 || R > 0x3fffffffLL || cfg_host->T < 1 || cfg_host->stride < 1 || cfg_host->ma_window < 0 || cfg_host->ma_window > 64 ||
        (cfg_host->ma_window > 1 && cfg_host->ma_window % 2 == 0))
        return SHM_ERR_UNSUPPORTED;          // the reference's centred window is odd (config.py:45)
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t r = (size_t)R;
    char* p = static_cast<char*>(workspace);
    ExtractWs ws;
    ws.i0 = reinterpret_cast<int*>(p); p += 256;
    ws.keep_idx = reinterpret_cast<int*>(p); p += align256(r * 4);
    ws.keep_count = rows_kept;
    p += 256;
    ws.masks = reinterpret_cast<float*>(p); p += align256(3 * r * 4);
    ws.mk = reinterpret_cast<float*>(p); p += align256(3 * r * 4);
    ws.keepf = reinterpret_cast<float*>(p); p += align256(r * 4);
    ws.compact_ws = p;
    const shm_openlab_extract_cfg cfg = *cfg_host;
    const int sms = device_sm_count(dev);
    const int grid = (int)min((long long)((R + 255) / 256), (long long)sms * 8);
    extract_init_kernel<<<1, 32, 0, st>>>(ws.i0, (int)R);
    SHM_LAUNCH_CHECK();
    extract_scan_kernel<<<grid, 256, 0, st>>>(raw, (int)R, cfg, ws.i0, ws.masks, ws.keepf);
    SHM_LAUNCH_CHECK();
    // rows with a finite DMS, ascending (np boolean indexing, 01:151-156) = the threshold/compaction primitive on 0/1 flags
    rc = shm_compact(ws.keepf, 0.5f, R, nullptr, ws.keep_idx, ws.keep_count, ws.compact_ws, stream);
    if (rc != SHM_OK) return rc;
    extract_gather_kernel<<<grid, 256, 0, st>>>(raw, (int)R, cfg, ws.i0, ws.keep_idx, ws.keep_count, ws.masks, a_clean, a_raw, ws.mk);
    SHM_LAUNCH_CHECK();
    const long long maxW = (R < cfg.T) ? 1 : (R - cfg.T) / cfg.stride + 1;
    const int wgrid = (int)min((maxW * 32 + 255) / 256, (long long)sms * 8);
    extract_window_kernel<<<wgrid, 256, 0, st>>>(a_clean, ws.mk, (int)R, ws.keep_count, cfg, n_windows, label, u_min, u_max, dms_range,
                                                 raw_invalid_ratio, raw_outlier_ratio, removed_ratio, flatline_loadaware, all_nan_struct);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
