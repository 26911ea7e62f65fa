// openLAB attribution CNN, eval mode: 20250506_openLAB_tests/Codes/Models/cnn_model.py:16-43,54-57.
//   x [n,1,200,4] -> [Conv(7x3,1->32)+GN8+SiLU] -> MaxPool(2,1) -> [Conv(5x3,32->64)+GN8+SiLU] -> MaxPool(2,1)
//     -> [Conv(5x3,64->128)+GN8+SiLU] -> MaxPool(2,1) -> [Conv(3x3,128->256)+GN8+SiLU] -> GAP
//     -> Linear 256->128 + SiLU (+Dropout = identity) -> Linear 128->2 ; prob = softmax[:,1] as fp64
// with the stage-2 input path of 10_test_hybrid_pipeline.py:272-278 (X_raw[mask] gather, standardise
// with the CNN's own mu/sd, clip 10, NaN->0) fused into the load of the first plane.
//
// One CTA per window: every activation plane lives in shared memory (pooled input plane 50 KB +
// pre-pool convolution plane 100 KB), GroupNorm statistics are per sample so nothing crosses CTAs.
// Each of the 320 threads owns a 4-channel x 5-row x 4-column register tile (80 accumulators); all
// four blocks factor into exactly 320 such tiles.  Weights are repacked to [ci][co/4][kt][kf][4] so a
// thread fetches 4 output channels per 16-byte read-only load (L1/L2 resident, shared by all CTAs).
#include <new>
#include "common.cuh"
#include "cnnol_tc.cuh"

struct shm_cnnol {
    int device;
    int engine;                 // SHM_ENGINE_FP32 or SHM_ENGINE_TC_BF16X3 (tensor cores, 3-pass fp16 split)
    shm::CnnOlTc tc;
    const float* raw_w[4];      // staged reference-layout conv weights inside `raw`
    float* buf;
    float* raw;
    size_t o_w[4], o_b[4], o_gw[4], o_gb[4], o_fc1t, o_fc1b, o_fc2w, o_fc2b, total, raw_total;
    float gn_eps;
};

namespace shm {

constexpr int OL_T = 200, OL_F = 4, OL_THREADS = 320;
constexpr int OL_CIN[4] = {1, 32, 64, 128};
constexpr int OL_COUT[4] = {32, 64, 128, 256};
constexpr int OL_KT[4] = {7, 5, 5, 3};
// plane heights per block: 200, 100, 50, 25
constexpr int OL_PLANE_IN = 12800, OL_PLANE_OUT = 25600;

struct CnnOlDev {
    const float* w[4];      // repacked conv weights
    const float* b[4];
    const float* gw[4];
    const float* gb[4];
    const float *fc1t, *fc1b, *fc2w, *fc2b;
    float gn_eps;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + expf(-x)); }

template <int CIN, int COUT, int KT, int HH>
__device__ __forceinline__ void conv_block(const float* __restrict__ wr, const float* __restrict__ bias,
                                           const float* __restrict__ in, float* __restrict__ out) {
    constexpr int PT = KT / 2;
    constexpr int RT = HH / 5;              // row tiles of 5
    constexpr int R = 5 + KT - 1;
    static_assert((COUT / 4) * RT == OL_THREADS, "block must factor into 320 thread tiles");
    const int tid = threadIdx.x;
    const int cg = tid / RT;
    const int rt = tid - cg * RT;
    const int r0 = rt * 5;
    float acc[4][5][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[c][r][q] = 0.f;

    const float4* wbase = reinterpret_cast<const float4*>(wr) + (size_t)cg * (KT * 3);
    for (int ci = 0; ci < CIN; ++ci) {
        float x[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int h = r0 + r - PT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (h >= 0 && h < HH) v = *reinterpret_cast<const float4*>(in + (ci * HH + h) * 4);
            x[r][0] = v.x; x[r][1] = v.y; x[r][2] = v.z; x[r][3] = v.w;
        }
        const float4* wc = wbase + (size_t)ci * (COUT / 4) * (KT * 3);
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
            for (int kf = 0; kf < 3; ++kf) {
                const float4 w4 = __ldg(wc + kt * 3 + kf);
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int qs = q + kf - 1;
                    if (qs < 0 || qs > 3) continue;
#pragma unroll
                    for (int r = 0; r < 5; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[c][r][q] = fmaf(wv[c], x[r + kt][qs], acc[c][r][q]);
                }
            }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float b = __ldg(bias + cg * 4 + c);
#pragma unroll
        for (int r = 0; r < 5; ++r)
            *reinterpret_cast<float4*>(out + ((cg * 4 + c) * HH + r0 + r) * 4) =
                make_float4(acc[c][r][0] + b, acc[c][r][1] + b, acc[c][r][2] + b, acc[c][r][3] + b);
    }
}

// GroupNorm(8) statistics of the 25600-element plane: group g = 3200 contiguous floats.
__device__ __forceinline__ void gn_stats(const float* __restrict__ plane, float eps, float* s_mean, float* s_rstd) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 8) {
        const float* g = plane + warp * 3200;
        float s = 0.f;
        for (int i = lane; i < 3200; i += 32) s += g[i];
        s = warp_sum(s);
        const float m = s / 3200.f;
        float v = 0.f;
        for (int i = lane; i < 3200; i += 32) { const float d = g[i] - m; v = fmaf(d, d, v); }
        v = warp_sum(v) / 3200.f;
        if (lane == 0) { s_mean[warp] = m; s_rstd[warp] = 1.0f / sqrtf(v + eps); }
    }
}

template <int COUT, int HH>
__device__ __forceinline__ void gn_silu_pool(const float* __restrict__ plane, const float* __restrict__ gw,
                                             const float* __restrict__ gb, const float* s_mean, const float* s_rstd,
                                             float* __restrict__ next) {
    constexpr int CPG = COUT / 8;
    constexpr int H2 = HH / 2;
    for (int i = threadIdx.x; i < COUT * H2; i += OL_THREADS) {
        const int c = i / H2, h2 = i - c * H2;
        const int g = c / CPG;
        const float m = s_mean[g], rs = s_rstd[g], w = __ldg(gw + c), b = __ldg(gb + c);
        const float4 a = *reinterpret_cast<const float4*>(plane + (c * HH + 2 * h2) * 4);
        const float4 d = *reinterpret_cast<const float4*>(plane + (c * HH + 2 * h2 + 1) * 4);
        float4 o;
        o.x = fmaxf(silu_f(fmaf((a.x - m) * rs, w, b)), silu_f(fmaf((d.x - m) * rs, w, b)));
        o.y = fmaxf(silu_f(fmaf((a.y - m) * rs, w, b)), silu_f(fmaf((d.y - m) * rs, w, b)));
        o.z = fmaxf(silu_f(fmaf((a.z - m) * rs, w, b)), silu_f(fmaf((d.z - m) * rs, w, b)));
        o.w = fmaxf(silu_f(fmaf((a.w - m) * rs, w, b)), silu_f(fmaf((d.w - m) * rs, w, b)));
        *reinterpret_cast<float4*>(next + (c * H2 + h2) * 4) = o;
    }
}

__global__ void __launch_bounds__(OL_THREADS, 1)
cnnol_kernel(CnnOlDev P, WinSrc src, const int* __restrict__ idx, const int* __restrict__ n_dev, long long n,
             float* __restrict__ logits, double* __restrict__ prob) {
    extern __shared__ __align__(16) float sm[];
    float* pin = sm;                               // pooled input plane of the current block
    float* pout = sm + OL_PLANE_IN;                // pre-pool convolution plane
    __shared__ float s_mean[8], s_rstd[8], s_gap[256], s_h[128];
    const int tid = threadIdx.x;
    long long n_eff = n;
    if (n_dev) n_eff = min(n_eff, (long long)__ldg(n_dev));

    for (long long w = blockIdx.x; w < n_eff; w += gridDim.x) {
        const long long win = idx ? (long long)idx[w] : w;
        __syncthreads();
        for (int i = tid; i < OL_T * OL_F; i += OL_THREADS) {
            const int t = i >> 2, d = i & 3;
            pin[i] = win_fetch(src, win, t, d);
        }
        __syncthreads();
        conv_block<1, 32, 7, 200>(P.w[0], P.b[0], pin, pout);
        __syncthreads();
        gn_stats(pout, P.gn_eps, s_mean, s_rstd);
        __syncthreads();
        gn_silu_pool<32, 200>(pout, P.gw[0], P.gb[0], s_mean, s_rstd, pin);
        __syncthreads();
        conv_block<32, 64, 5, 100>(P.w[1], P.b[1], pin, pout);
        __syncthreads();
        gn_stats(pout, P.gn_eps, s_mean, s_rstd);
        __syncthreads();
        gn_silu_pool<64, 100>(pout, P.gw[1], P.gb[1], s_mean, s_rstd, pin);
        __syncthreads();
        conv_block<64, 128, 5, 50>(P.w[2], P.b[2], pin, pout);
        __syncthreads();
        gn_stats(pout, P.gn_eps, s_mean, s_rstd);
        __syncthreads();
        gn_silu_pool<128, 50>(pout, P.gw[2], P.gb[2], s_mean, s_rstd, pin);
        __syncthreads();
        conv_block<128, 256, 3, 25>(P.w[3], P.b[3], pin, pout);
        __syncthreads();
        gn_stats(pout, P.gn_eps, s_mean, s_rstd);
        __syncthreads();
        // GroupNorm + SiLU + global average pool over 25x4
        if (tid < 256) {
            const int g = tid >> 5;
            const float m = s_mean[g], rs = s_rstd[g], gw = __ldg(P.gw[3] + tid), gb = __ldg(P.gb[3] + tid);
            float s = 0.f;
            for (int i = 0; i < 100; ++i) s += silu_f(fmaf((pout[tid * 100 + i] - m) * rs, gw, gb));
            s_gap[tid] = s / 100.f;
        }
        __syncthreads();
        if (tid < 128) {
            float y = __ldg(P.fc1b + tid);
            for (int k = 0; k < 256; ++k) y = fmaf(s_gap[k], __ldg(P.fc1t + k * 128 + tid), y);
            s_h[tid] = silu_f(y);
        }
        __syncthreads();
        if (tid < 32) {
            float l0 = 0.f, l1 = 0.f;
            for (int k = tid; k < 128; k += 32) {
                l0 = fmaf(s_h[k], __ldg(P.fc2w + k), l0);
                l1 = fmaf(s_h[k], __ldg(P.fc2w + 128 + k), l1);
            }
            l0 = warp_sum(l0) + __ldg(P.fc2b);
            l1 = warp_sum(l1) + __ldg(P.fc2b + 1);
            if (tid == 0) {
                logits[w * 2] = l0;
                logits[w * 2 + 1] = l1;
                if (prob) {
                    const float m = fmaxf(l0, l1);
                    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
                    prob[w] = (double)(e1 / (e0 + e1));
                }
            }
        }
    }
}

// [co][ci][kt][kf] -> [ci][co/4][kt][kf][co%4]
__global__ void cnnol_pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int kt, float* __restrict__ out) {
    const int total = cout * cin * kt * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int kf = r % 3; r /= 3;
        const int k = r % kt; r /= kt;
        const int ci = r % cin;
        const int co = r / cin;
        out[((((size_t)ci * (cout / 4) + co / 4) * kt + k) * 3 + kf) * 4 + (co & 3)] = w[i];
    }
}
__global__ void cnnol_transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        out[(size_t)c * rows + r] = in[i];
    }
}

}  // namespace shm

using namespace shm;

static int cnnol_upload(shm_cnnol* h, const shm_cnnol_weights* w, cudaStream_t st) {
    size_t off = 0;
    auto stage = [&](const float* src, size_t n, float** dst) -> int {
        if (!src) return SHM_ERR_ARG;
        *dst = h->raw + off;
        off += (n + 3) / 4 * 4;
        SHM_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDefault, st));
        return SHM_OK;
    };
    auto direct = [&](size_t o, const float* src, size_t n) -> int {
        if (!src) return SHM_ERR_ARG;
        SHM_CUDA(cudaMemcpyAsync(h->buf + o, src, n * sizeof(float), cudaMemcpyDefault, st));
        return SHM_OK;
    };
    int rc;
    for (int b = 0; b < 4; ++b) {
        float* d;
        const size_t nw = (size_t)OL_COUT[b] * OL_CIN[b] * OL_KT[b] * 3;
        if ((rc = stage(w->conv_w[b], nw, &d))) return rc;
        h->raw_w[b] = d;
        cnnol_pack_conv_kernel<<<148, 256, 0, st>>>(d, OL_COUT[b], OL_CIN[b], OL_KT[b], h->buf + h->o_w[b]);
        SHM_LAUNCH_CHECK();
        if ((rc = direct(h->o_b[b], w->conv_b[b], OL_COUT[b]))) return rc;
        if ((rc = direct(h->o_gw[b], w->gn_w[b], OL_COUT[b]))) return rc;
        if ((rc = direct(h->o_gb[b], w->gn_b[b], OL_COUT[b]))) return rc;
    }
    float* d;
    if ((rc = stage(w->fc1_w, 128 * 256, &d))) return rc;
    cnnol_transpose_kernel<<<64, 256, 0, st>>>(d, 128, 256, h->buf + h->o_fc1t);
    SHM_LAUNCH_CHECK();
    if ((rc = direct(h->o_fc1b, w->fc1_b, 128))) return rc;
    if ((rc = direct(h->o_fc2w, w->fc2_w, 256))) return rc;
    if ((rc = direct(h->o_fc2b, w->fc2_b, 2))) return rc;
    h->gn_eps = w->gn_eps > 0.f ? w->gn_eps : 1e-5f;
    return cnnol_tc_pack(&h->tc, h->raw_w, st);
}

static constexpr size_t kCnnOlSmem = (size_t)(OL_PLANE_IN + OL_PLANE_OUT) * sizeof(float);

extern "C" int shm_cnnol_create(shm_cnnol** out, const shm_cnnol_weights* w, int device) {
    if (!out || !w) return SHM_ERR_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != SHM_OK) return rc;
    int prev = 0;
    SHM_CUDA(cudaGetDevice(&prev));
    SHM_CUDA(cudaSetDevice(device));
    shm_cnnol* h = new (std::nothrow) shm_cnnol();
    if (!h) return SHM_ERR_NOMEM;
    memset(h, 0, sizeof(*h));
    h->device = device;
    size_t off = 0, roff = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return o; };
    for (int b = 0; b < 4; ++b) {
        const size_t nw = (size_t)OL_COUT[b] * OL_CIN[b] * OL_KT[b] * 3;
        h->o_w[b] = take(nw); h->o_b[b] = take(OL_COUT[b]); h->o_gw[b] = take(OL_COUT[b]); h->o_gb[b] = take(OL_COUT[b]);
        roff += (nw + 3) / 4 * 4;
    }
    h->o_fc1t = take(128 * 256); h->o_fc1b = take(128); h->o_fc2w = take(256); h->o_fc2b = take(4);
    roff += 128 * 256;
    h->total = off; h->raw_total = roff;
    if (cudaMalloc(&h->buf, h->total * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&h->raw, h->raw_total * sizeof(float)) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaMalloc(cnnol)");
        shm_cnnol_destroy(h);
        cudaSetDevice(prev);
        return SHM_ERR_NOMEM;
    }
    h->engine = SHM_ENGINE_TC_BF16X3;
    rc = cnnol_tc_init(&h->tc, device);
    if (rc == SHM_OK) rc = cnnol_upload(h, w, 0);
    if (rc == SHM_OK && cudaStreamSynchronize(0) != cudaSuccess) { set_cuda_error(cudaGetLastError(), "cnnol create"); rc = SHM_ERR_CUDA; }
    if (rc == SHM_OK && cudaFuncSetAttribute(cnnol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCnnOlSmem) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaFuncSetAttribute(cnnol)");
        rc = SHM_ERR_CUDA;
    }
    cudaSetDevice(prev);
    if (rc != SHM_OK) { shm_cnnol_destroy(h); return rc; }
    *out = h;
    return SHM_OK;
}

extern "C" int shm_cnnol_update_weights(shm_cnnol* h, const shm_cnnol_weights* w, void* stream) {
    if (!h || !w) return SHM_ERR_ARG;
    return cnnol_upload(h, w, static_cast<cudaStream_t>(stream));
}

extern "C" int shm_cnnol_destroy(shm_cnnol* h) {
    if (!h) return SHM_OK;
    if (h->buf) cudaFree(h->buf);
    if (h->raw) cudaFree(h->raw);
    cnnol_tc_free(&h->tc);
    delete h;
    return SHM_OK;
}

extern "C" int shm_cnnol_forward(shm_cnnol* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev,
                                 int64_t n, float* logits, double* prob, void* stream) {
    if (!h || !src_host || n < 0 || (n > 0 && !logits)) return SHM_ERR_ARG;
    WinSrc src;
    int rc = make_winsrc(src_host, &src);
    if (rc != SHM_OK) return rc;
    if (src.T != OL_T || src.D != OL_F) return SHM_ERR_ARG;
    if (n == 0) return SHM_OK;
    CnnOlDev P;
    for (int b = 0; b < 4; ++b) {
        P.w[b] = h->buf + h->o_w[b]; P.b[b] = h->buf + h->o_b[b]; P.gw[b] = h->buf + h->o_gw[b]; P.gb[b] = h->buf + h->o_gb[b];
    }
    P.fc1t = h->buf + h->o_fc1t; P.fc1b = h->buf + h->o_fc1b; P.fc2w = h->buf + h->o_fc2w; P.fc2b = h->buf + h->o_fc2b;
    P.gn_eps = h->gn_eps;
    if (h->engine == SHM_ENGINE_TC_BF16X3)
        return cnnol_tc_forward(&h->tc, h->raw_w[0], P.b, P.gw, P.gb, P.fc1t, P.fc1b, P.fc2w, P.fc2b, h->gn_eps, src, idx, n_dev, n,
                                logits, prob, static_cast<cudaStream_t>(stream));
    const int sms = device_sm_count(h->device);
    const int grid = (int)min((long long)n, (long long)sms * 8);
    cnnol_kernel<<<grid, OL_THREADS, kCnnOlSmem, static_cast<cudaStream_t>(stream)>>>(P, src, idx, n_dev, n, logits, prob);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

extern "C" int shm_cnnol_set_engine(shm_cnnol* h, int engine) {
    if (!h) return SHM_ERR_ARG;
    if (engine == SHM_ENGINE_AUTO) engine = SHM_ENGINE_TC_BF16X3;
    if (engine != SHM_ENGINE_FP32 && engine != SHM_ENGINE_TC_BF16X3) return SHM_ERR_UNSUPPORTED;
    h->engine = engine;
    return SHM_OK;
}

extern "C" int shm_cnnol_engine(const shm_cnnol* h) { return h ? h->engine : SHM_ERR_ARG; }
