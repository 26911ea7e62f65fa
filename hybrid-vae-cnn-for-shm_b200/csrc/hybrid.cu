// Fused hybrid entry points: the whole script-level loop of the reference as ONE C call on one stream, with no host
// synchronisation (the flagged count never leaves the device) and bounded scratch:
//   shm_hybrid4dof_score   == eval_group, 4DOF/Scripts/06_test_full_pipeline.py:327-383
//   shm_hybridol_score     == 10_test_hybrid_pipeline.py:351-367 + stage2_predict_cnn (:265-302) + the label scatter (:387-401)
//   shm_scatter_flagged_*  == `y_pred[sel] = cls + 1; hyb_score_full[sel] = p_struct` (06:368-372)
#include "common.cuh"

namespace shm {

constexpr long long HY_CHUNK = 75776;     // flagged windows per second-pass chunk = 4 whole waves of 128-window tiles on 148 SMs (65,536 = 3.46 waves left the
                                          // scorer a 13 % tail per chunk); bounds cnn_in at 727 MB for the 4DOF shape

struct HyLayout {
    size_t mu, lv, cnn_in, label, p, logits, cnt, compact, total;
    long long chunk, n_chunks;
};

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

static HyLayout hy_layout4(long long n, long long cap, int Z, int T, int D) {
    HyLayout L;
    L.chunk = cap < HY_CHUNK ? (cap > 0 ? cap : 1) : HY_CHUNK;
    L.n_chunks = cap > 0 ? (cap + L.chunk - 1) / L.chunk : 0;
    size_t o = 0;
    auto take = [&](size_t b) { size_t r = o; o += al256(b); return r; };
    L.mu = take((size_t)n * Z * 4); L.lv = take((size_t)n * Z * 4);
    L.cnn_in = take((size_t)L.chunk * 2 * T * D * 4);
    L.label = take((size_t)L.chunk * 8); L.p = take((size_t)L.chunk * 4); L.logits = take((size_t)L.chunk * 8);
    L.cnt = take((size_t)(L.n_chunks + 2) * 4);
    L.compact = take((size_t)shm_compact_workspace_bytes(n));
    L.total = o;
    return L;
}

// cnt[k] = number of flagged windows chunk k works on; status = {count, count > cap}
__global__ void hy_chunk_counts_kernel(const int* __restrict__ count, long long cap, long long chunk, int n_chunks,
                                       int* __restrict__ cnt, int* __restrict__ status) {
    const long long c = *count;
    const long long eff = c < cap ? c : cap;
    for (int k = threadIdx.x; k < n_chunks; k += blockDim.x) {
        long long v = eff - (long long)k * chunk;
        cnt[k] = (int)(v < 0 ? 0 : (v > chunk ? chunk : v));
    }
    if (threadIdx.x == 0 && status) { status[0] = (int)c; status[1] = c > cap ? 1 : 0; }
}

__global__ void hy_scatter4_kernel(const int* __restrict__ idx, const int* __restrict__ cnt, long long cap,
                                   const long long* __restrict__ label, const float* __restrict__ p,
                                   long long* __restrict__ y_pred, float* __restrict__ p_full) {
    long long n = cap;
    if (cnt) n = min(n, (long long)__ldg(cnt));
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const int w = __ldg(idx + j);
        if (y_pred) y_pred[w] = label[j];
        if (p_full) p_full[w] = p[j];
    }
}

// openLAB: prob_st (fp64) >= thr -> pred_bin; dense label 1 = sensor fault (pred 0), 2 = structural (pred 1), 0 = not flagged
__global__ void hy_scatter_ol_kernel(const int* __restrict__ idx, const int* __restrict__ cnt, long long cap,
                                     const double* __restrict__ prob, double thr, long long* __restrict__ pred_bin,
                                     long long* __restrict__ y_pred, double* __restrict__ prob_full) {
    long long n = cap;
    if (cnt) n = min(n, (long long)__ldg(cnt));
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const double pr = prob[j];
        const long long b = pr >= thr ? 1 : 0;
        if (pred_bin) pred_bin[j] = b;
        if (idx) {
            const int w = __ldg(idx + j);
            if (y_pred) y_pred[w] = 1 + b;
            if (prob_full) prob_full[w] = pr;
        }
    }
}

static inline int scatter_grid(long long n, int device) {
    const long long want = (n + 255) / 256;
    const long long cap = (long long)device_sm_count(device) * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace shm

using namespace shm;

static int cur_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

extern "C" int shm_scatter_flagged_4dof(const int32_t* idx, const int32_t* count, int64_t cap, const int64_t* label,
                                        const float* p_struct, int64_t n, int64_t* y_pred, float* p_full, void* stream) {
    if (cap < 0 || n < 0 || (cap > 0 && !idx) || (y_pred && !label) || (p_full && !p_struct)) return SHM_ERR_ARG;
    const int dev = cur_device();
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_pred && n) SHM_CUDA(cudaMemsetAsync(y_pred, 0, (size_t)n * 8, st));
    if (p_full && n) SHM_CUDA(cudaMemsetAsync(p_full, 0, (size_t)n * 4, st));
    if (cap == 0) return SHM_OK;
    hy_scatter4_kernel<<<scatter_grid(cap, dev), 256, 0, st>>>(idx, count, cap, reinterpret_cast<const long long*>(label), p_struct,
                                                               reinterpret_cast<long long*>(y_pred), p_full);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

extern "C" int shm_scatter_flagged_openlab(const int32_t* idx, const int32_t* count, int64_t cap, const double* prob, double cnn_thr,
                                           int64_t n, int64_t* pred_bin, int64_t* y_pred, double* prob_full, void* stream) {
    if (cap < 0 || n < 0 || (cap > 0 && !prob) || ((y_pred || prob_full) && cap > 0 && !idx)) return SHM_ERR_ARG;
    const int dev = cur_device();
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (y_pred && n) SHM_CUDA(cudaMemsetAsync(y_pred, 0, (size_t)n * 8, st));
    if (prob_full && n) SHM_CUDA(cudaMemsetAsync(prob_full, 0, (size_t)n * 8, st));
    if (cap == 0) return SHM_OK;
    hy_scatter_ol_kernel<<<scatter_grid(cap, dev), 256, 0, st>>>(idx, count, cap, prob, cnn_thr, reinterpret_cast<long long*>(pred_bin),
                                                                 reinterpret_cast<long long*>(y_pred), prob_full);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

extern "C" int64_t shm_hybrid4dof_workspace_bytes(const shm_vae* vae, int64_t n, int64_t max_flagged) {
    shm_vae_cfg cfg;
    if (n < 0 || max_flagged < 0 || shm_vae_get_cfg(vae, &cfg) != SHM_OK) return SHM_ERR_ARG;
    return (int64_t)hy_layout4(n, max_flagged < n ? max_flagged : n, cfg.Z, 100, cfg.D).total;
}

extern "C" int shm_hybrid4dof_score(shm_vae* vae, shm_cnn4dof* cnn, const shm_window_src* src, int64_t n,
                                    const float* eps1, const float* eps2, float thr, int64_t max_flagged, float* score,
                                    uint8_t* flag, int32_t* idx, int32_t* status, float* logits, int64_t* label, float* p_struct,
                                    int64_t* y_pred, float* p_full, void* workspace, int64_t workspace_bytes, void* stream) {
    shm_vae_cfg cfg_v;
    if (!vae || !cnn || !src || n < 0 || max_flagged < 0 || !score || !idx || !status || !workspace) return SHM_ERR_ARG;
    if (shm_vae_get_cfg(vae, &cfg_v) != SHM_OK) return SHM_ERR_ARG;
    const shm_vae_cfg* cfg = &cfg_v;
    if (src->T != 100 || src->D != 12 || cfg->D != 12) return SHM_ERR_ARG;          // the 4DOF CNN is (2,100,12) only
    const long long cap = max_flagged < n ? max_flagged : n;
    const HyLayout L = hy_layout4(n, cap, cfg->Z, src->T, src->D);
    if ((size_t)workspace_bytes < L.total) return SHM_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    float* mu = reinterpret_cast<float*>(ws + L.mu);
    float* lv = reinterpret_cast<float*>(ws + L.lv);
    int* cnt = reinterpret_cast<int*>(ws + L.cnt);
    if (y_pred && n) SHM_CUDA(cudaMemsetAsync(y_pred, 0, (size_t)n * 8, st));
    if (p_full && n) SHM_CUDA(cudaMemsetAsync(p_full, 0, (size_t)n * 4, st));
    if (n == 0) { SHM_CUDA(cudaMemsetAsync(status, 0, 8, st)); return SHM_OK; }
    int rc = shm_vae_score(vae, src, nullptr, nullptr, eps1, n, score, mu, lv, nullptr, nullptr, stream);
    if (rc != SHM_OK) return rc;
    rc = shm_compact(score, thr, n, flag, idx, status, ws + L.compact, stream);
    if (rc != SHM_OK) return rc;
    hy_chunk_counts_kernel<<<1, 128, 0, st>>>(status, cap, L.chunk, (int)L.n_chunks, cnt, status);
    SHM_LAUNCH_CHECK();
    const int dev = cur_device();
    for (long long k = 0; k < L.n_chunks; ++k) {
        const long long base = k * L.chunk;
        const long long nk = (cap - base) < L.chunk ? (cap - base) : L.chunk;
        float* cin = reinterpret_cast<float*>(ws + L.cnn_in);
        const float* e2 = eps2 ? eps2 + (size_t)base * cfg->Z : nullptr;
        // second pass with fresh noise on flagged windows [base, base+nk): encoder outputs of the first pass are reused
        rc = shm_vae_rescore(vae, src, idx + base, cnt + k, mu, lv, e2, nk, nullptr, nullptr, cin, stream);
        if (rc == SHM_ERR_UNSUPPORTED)      // engine without the re-score entry: the full forward gives the same result
            rc = shm_vae_score(vae, src, idx + base, cnt + k, e2, nk, nullptr, nullptr, nullptr, nullptr, cin, stream);
        if (rc != SHM_OK) return rc;
        float* lg = logits ? logits + (size_t)base * 2 : reinterpret_cast<float*>(ws + L.logits);
        int64_t* lb = label ? label + base : reinterpret_cast<int64_t*>(ws + L.label);
        float* pp = p_struct ? p_struct + base : reinterpret_cast<float*>(ws + L.p);
        rc = shm_cnn4dof_forward(cnn, cin, cnt + k, nk, lg, lb, pp, stream);
        if (rc != SHM_OK) return rc;
        if (y_pred || p_full) {
            hy_scatter4_kernel<<<scatter_grid(nk, dev), 256, 0, st>>>(idx + base, cnt + k, nk, reinterpret_cast<const long long*>(lb), pp,
                                                                      reinterpret_cast<long long*>(y_pred), p_full);
            SHM_LAUNCH_CHECK();
        }
    }
    return SHM_OK;
}

extern "C" int64_t shm_hybridol_workspace_bytes(int64_t n, int64_t max_flagged) {
    if (n < 0 || max_flagged < 0) return SHM_ERR_ARG;
    const long long cap = max_flagged < n ? max_flagged : n;
    return (int64_t)(al256((size_t)shm_compact_workspace_bytes(n)) + al256((size_t)cap * 8) + al256((size_t)cap * 8) + 256);
}

extern "C" int shm_hybridol_score(shm_vae* vae, shm_cnnol* cnn, const shm_window_src* src_gate, const shm_window_src* src_raw,
                                  int64_t n, const float* eps, float vae_thr, double cnn_thr, int64_t max_flagged, float* score,
                                  uint8_t* flag, int32_t* idx, int32_t* status, float* logits, double* prob, int64_t* pred_bin,
                                  int64_t* y_pred, double* prob_full, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!vae || !cnn || !src_gate || !src_raw || n < 0 || max_flagged < 0 || !score || !idx || !status || !workspace) return SHM_ERR_ARG;
    const long long cap = max_flagged < n ? max_flagged : n;
    if (workspace_bytes < shm_hybridol_workspace_bytes(n, max_flagged)) return SHM_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    size_t o = al256((size_t)shm_compact_workspace_bytes(n));
    double* prob_ws = reinterpret_cast<double*>(ws + o); o += al256((size_t)cap * 8);
    float* logit_ws = reinterpret_cast<float*>(ws + o); o += al256((size_t)cap * 8);
    int* cnt = reinterpret_cast<int*>(ws + o);
    if (y_pred && n) SHM_CUDA(cudaMemsetAsync(y_pred, 0, (size_t)n * 8, st));
    if (prob_full && n) SHM_CUDA(cudaMemsetAsync(prob_full, 0, (size_t)n * 8, st));
    if (n == 0) { SHM_CUDA(cudaMemsetAsync(status, 0, 8, st)); return SHM_OK; }
    int rc = shm_vae_score(vae, src_gate, nullptr, nullptr, eps, n, score, nullptr, nullptr, nullptr, nullptr, stream);
    if (rc != SHM_OK) return rc;
    rc = shm_compact(score, vae_thr, n, flag, idx, status, ws, stream);
    if (rc != SHM_OK) return rc;
    hy_chunk_counts_kernel<<<1, 32, 0, st>>>(status, cap, cap > 0 ? cap : 1, cap > 0 ? 1 : 0, cnt, status);
    SHM_LAUNCH_CHECK();
    if (cap == 0) return SHM_OK;
    double* pr = prob ? prob : prob_ws;
    rc = shm_cnnol_forward(cnn, src_raw, idx, cnt, cap, logits ? logits : logit_ws, pr, stream);
    if (rc != SHM_OK) return rc;
    hy_scatter_ol_kernel<<<scatter_grid(cap, cur_device()), 256, 0, st>>>(idx, cnt, cap, pr, cnn_thr, reinterpret_cast<long long*>(pred_bin),
                                                                          reinterpret_cast<long long*>(y_pred), prob_full);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
