// CNN training steps for libshmfast (sm_100a): forward in train() mode with saved activations + full backward for
//   * the 4DOF CNN (4DOF/Scripts/Models/cnn_model.py:16-51): 2 x [Conv3x3 + BatchNorm(batch statistics, running-stat
//     update) + ReLU + MaxPool2] + Linear(2400,128) + ReLU + Dropout + Linear(128,2) -- inner loop of
//     4DOF/Scripts/05_train_cnn.py:266-281 (CrossEntropyLoss, Adam lr 1e-4 wd 5e-5);
//   * the openLAB CNN (20250506_openLAB_tests/Codes/Models/cnn_model.py:16-57): 4 x [Conv(kt x 3) + GroupNorm(8) + SiLU]
//     with MaxPool(2,1) x 3 and a global average pool, Linear(256,128) + SiLU + Dropout + Linear(128,2) -- inner loop of
//     Codes/06_train_cnn.py:410-421 (weighted focal loss gamma 2, clip_grad_norm_ 2.0, AdamW lr 3e-4 wd 1e-4).
//
// All tensors NCHW fp32 (the reference's layout).  Convolutions (all stride 1, "same" padding) are implicit GEMMs on the
// fp32 FMA pipe: forward / backward-data share one kernel over re-packed weights, backward-weight is a split-K
// contraction over (sample, position).  Normalisation + activation + pooling are fused per block in both directions;
// the pooling argmax is recomputed from the saved convolution output instead of being stored.
#include <new>
#include "common.cuh"
#include "gemm.cuh"

namespace shm {

constexpr int CT_BM = 64, CT_BN = 64, CT_BK = 16;

struct ConvG {            // stride-1 convolution with "same" padding, NCHW
    int B, Cin, Cout, H, W, KH, KW, PH, PW;
};

// W[co][ci][khw] -> Wf[(khw*Cin + ci)][co] (forward operand) and Wb[(khw*Cout + co)][ci] (backward-data operand): TAP-major K, so
// that a 16-wide K tile of the implicit GEMM is 16 channels at ONE filter tap (no per-element index arithmetic in the loader)
__global__ void conv_repack_kernel(const float* __restrict__ W, int Cout, int Cin, int KHW, float* __restrict__ Wf,
                                   float* __restrict__ Wb) {
    const int n = Cout * Cin * KHW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int co = i / (Cin * KHW);
        const int r = i - co * (Cin * KHW);
        const int ci = r / KHW, khw = r - ci * KHW;
        const float v = W[i];
        Wf[(size_t)(khw * Cin + ci) * Cout + co] = v;
        Wb[(size_t)(khw * Cout + co) * Cin + ci] = v;
    }
}

// MODE 0 (forward):       dst[b,n,h,w] = bias[n] + sum_{kh,kw,ci} src[b,ci,h+kh-PH,w+kw-PW] * Wm[(kh,kw,ci)][n],  n < Cout
// MODE 1 (backward data): dst[b,n,h,w] =           sum_{kh,kw,co} src[b,co,h-kh+PH,w-kw+PW] * Wm[(kh,kw,co)][n],  n < Cin
// GEMM view: M = B*H*W rows (b,h,w), K = KH*KW*Csrc (tap-major: ncu showed the channel-major loader spending 40 % of all issue slots
// on ALU index arithmetic, FMA 38 % -- profiles/r02_cnn_train_conv_raw.csv), N columns; 128 x 64 x 16 tiles, 256 threads, 8 x 4 per thread (two groups of 4
// consecutive rows: LDS.128 for both operands, 3 loads per 32 FMAs, float4 stores along w -- H*W is a multiple of 4 in every
// block, so a row quad never straddles a sample).  r02 first version: 64 x 64 tiles, 4 x 4 per thread, 5 loads per 16 FMAs.
constexpr int CI_BM = 128, CI_BN = 64;
template <int MODE>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvG s, const float* __restrict__ src, const float* __restrict__ Wm,
                                                         const float* __restrict__ bias, float* __restrict__ dst) {
    const int HW = s.H * s.W, KHW = s.KH * s.KW;
    const int Csrc = MODE == 0 ? s.Cin : s.Cout;
    const int N = MODE == 0 ? s.Cout : s.Cin;
    const int M = s.B * HW, K = Csrc * KHW;
    __shared__ __align__(16) float As[CT_BK][CI_BM + 4];
    __shared__ __align__(16) float Bs[CT_BK][CI_BN + 4];
    const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
    const int m0 = blockIdx.x * CI_BM, n0 = blockIdx.y * CI_BN;
    // loader coordinates: A: row m_l, k = (tid >> 7) + 2e (8 per thread);  B: column n_l, k = (tid >> 6) + 4e (4 per thread)
    const int m_l = tid & 127, kqa = tid >> 7;
    const int gm = m0 + m_l;
    const bool m_ok = gm < M;
    const int b_l = m_ok ? gm / HW : 0;
    const int hw_l = m_ok ? gm - b_l * HW : 0;
    const int h_l = hw_l / s.W, w_l = hw_l - h_l * s.W;
    const float* src_b = src + (size_t)b_l * Csrc * HW;
    const int n_l = tid & 63, kqb = tid >> 6;
    const int gn_l = n0 + n_l;
    const bool n_ok = gn_l < N;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float ra[8], rb[4];
    const bool fast = (Csrc % CT_BK) == 0;                 // a K tile never straddles a filter tap
    auto fetch = [&](int k0) {
        if (fast) {
            const int tap = k0 / Csrc, c0 = k0 - tap * Csrc;
            const int kh = tap / s.KW, kw = tap - kh * s.KW;
            const int hh = MODE == 0 ? h_l + kh - s.PH : h_l - kh + s.PH;
            const int ww = MODE == 0 ? w_l + kw - s.PW : w_l - kw + s.PW;
            const bool ok = m_ok && hh >= 0 && hh < s.H && ww >= 0 && ww < s.W;
            const float* base = src_b + (size_t)(c0 + kqa) * HW + hh * s.W + ww;
#pragma unroll
            for (int e = 0; e < 8; ++e) ra[e] = ok ? __ldg(base + (size_t)(2 * e) * HW) : 0.f;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = k0 + kqa + 2 * e;
                float a = 0.f;
                if (k < K && m_ok) {
                    const int tap = k / Csrc;
                    const int c = k - tap * Csrc;
                    const int kh = tap / s.KW, kw = tap - kh * s.KW;
                    const int hh = MODE == 0 ? h_l + kh - s.PH : h_l - kh + s.PH;
                    const int ww = MODE == 0 ? w_l + kw - s.PW : w_l - kw + s.PW;
                    if (hh >= 0 && hh < s.H && ww >= 0 && ww < s.W) a = __ldg(src_b + (size_t)c * HW + hh * s.W + ww);
                }
                ra[e] = a;
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + kqb + 4 * e;
            rb[e] = (k < K && n_ok) ? __ldg(Wm + (size_t)k * N + gn_l) : 0.f;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += CT_BK) {
#pragma unroll
        for (int e = 0; e < 8; ++e) As[kqa + 2 * e][m_l] = ra[e];
#pragma unroll
        for (int e = 0; e < 4; ++e) Bs[kqb + 4 * e][n_l] = rb[e];
        __syncthreads();
        if (k0 + CT_BK < K) fetch(k0 + CT_BK);
#pragma unroll
        for (int k = 0; k < CT_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + tm * 4]);
            const float4 bq = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int m = m0 + half * 64 + tm * 4;            // 4 consecutive rows = 4 consecutive positions of one sample
        if (m >= M) continue;
        const int b = m / HW, hw = m - b * HW;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n >= N) continue;
            const float bv = (MODE == 0 && bias) ? __ldg(bias + n) : 0.f;
            float* d = dst + ((size_t)b * N + n) * HW + hw;
            if (m + 3 < M) {
                *reinterpret_cast<float4*>(d) = make_float4(acc[half * 4][j] + bv, acc[half * 4 + 1][j] + bv, acc[half * 4 + 2][j] + bv,
                                                            acc[half * 4 + 3][j] + bv);
            } else {
                for (int i = 0; i < 4 && m + i < M; ++i) d[i] = acc[half * 4 + i][j] + bv;
            }
        }
    }
}

// Backward weight: dW[co][(ci,kh,kw)] += sum_{b,h,w} dy[b,co,h,w] * x[b,ci,h+kh-PH,w+kw-PW].  GEMM view: M = Cout,
// N = Cin*KH*KW, K = (b, hw) in tiles of 16 positions that never straddle a sample; blockIdx.z splits the samples;
// atomicAdd epilogue into the zeroed gradient.
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const ConvG s, const float* __restrict__ dy, const float* __restrict__ x,
                                                         float* __restrict__ dW, int b_per_split) {
    const int HW = s.H * s.W, KHW = s.KH * s.KW;
    const int M = s.Cout, N = s.Cin * KHW;
    __shared__ __align__(16) float As[CT_BK][CT_BM + 4];
    __shared__ __align__(16) float Bs[CT_BK][CT_BN + 4];
    const int tid = threadIdx.x, tm = tid & 15, tn = tid >> 4;
    const int m0 = blockIdx.x * CT_BM, n0 = blockIdx.y * CT_BN;
    const int b_beg = blockIdx.z * b_per_split, b_end = min(s.B, b_beg + b_per_split);
    const int kk = tid & 15, rq = tid >> 4;            // loader: position kk of the tile, rows / columns rq + 16e
    const int tpi = (HW + CT_BK - 1) / CT_BK;          // k tiles per image
    // the 4 columns n this thread loads: (ci, dh, dw)
    int coff[4], dh[4], dw[4];
    bool nok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int n = n0 + rq + 16 * e;
        nok[e] = n < N;
        const int ci = nok[e] ? n / KHW : 0;
        const int r = nok[e] ? n - ci * KHW : 0;
        const int kh = r / s.KW, kw = r - kh * s.KW;
        coff[e] = ci * HW; dh[e] = kh - s.PH; dw[e] = kw - s.PW;
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int n_tiles = (b_end - b_beg) * tpi;
    float ra[4], rb[4];
    auto fetch = [&](int t) {
        const int b = b_beg + t / tpi;
        const int hw = (t - (t / tpi) * tpi) * CT_BK + kk;
        const bool ok = hw < HW;
        const int h = hw / s.W, w = hw - h * s.W;
        const float* dyb = dy + (size_t)b * s.Cout * HW;
        const float* xb = x + (size_t)b * s.Cin * HW;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int m = m0 + rq + 16 * e;
            ra[e] = (ok && m < M) ? __ldg(dyb + (size_t)m * HW + hw) : 0.f;
            const int hh = h + dh[e], ww = w + dw[e];
            rb[e] = (ok && nok[e] && hh >= 0 && hh < s.H && ww >= 0 && ww < s.W) ? __ldg(xb + coff[e] + hh * s.W + ww) : 0.f;
        }
    };
    if (n_tiles > 0) fetch(0);
    for (int t = 0; t < n_tiles; ++t) {
#pragma unroll
        for (int e = 0; e < 4; ++e) { As[kk][rq + 16 * e] = ra[e]; Bs[kk][rq + 16 * e] = rb[e]; }
        __syncthreads();
        if (t + 1 < n_tiles) fetch(t + 1);
#pragma unroll
        for (int k = 0; k < CT_BK; ++k) {
            const float4 aq = *reinterpret_cast<const float4*>(&As[k][tm * 4]);
            const float4 bq = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
            const float a[4] = {aq.x, aq.y, aq.z, aq.w};
            const float bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + tm * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tn * 4 + j;
            if (n < N) atomicAdd(dW + (size_t)m * N + n, acc[i][j]);
        }
    }
}

// part[(c*S + s)*2 + {0,1}] = sum / sum of squares of src[b,c,hw] over the samples of split s (blockIdx.y), fp64, fixed order;
// used for the convolution bias gradient (sum only) and the BatchNorm batch statistics.  r02: 16-32 CTAs walking 120 k
// elements each took 150-170 us per call (profiles/r02_launches_cnn_train_4dof.csv): the batch is now split over gridDim.y CTAs.
__global__ void __launch_bounds__(256) chan_partial_kernel(const float* __restrict__ src, int B, int C, int HW, int b_per_split,
                                                           double* __restrict__ part) {
    const int c = blockIdx.x, S = gridDim.y, sp = blockIdx.y;
    const int b_beg = sp * b_per_split, b_end = min(B, b_beg + b_per_split);
    double s = 0.0, q = 0.0;
    const int n = max(0, b_end - b_beg) * HW;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int b = b_beg + i / HW, hw = i % HW;
        const double v = src[((size_t)b * C + c) * HW + hw];
        s += v; q += v * v;
    }
    __shared__ double rs_[256], rq_[256];
    rs_[threadIdx.x] = s; rq_[threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { rs_[threadIdx.x] += rs_[threadIdx.x + o]; rq_[threadIdx.x] += rq_[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part[((size_t)c * S + sp) * 2] = rs_[0]; part[((size_t)c * S + sp) * 2 + 1] = rq_[0]; }
}

// out[c] = sum over the S partials (fixed order)
__global__ void chan_sum_final_kernel(const double* __restrict__ part, int C, int S, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s = 0.0;
    for (int i = 0; i < S; ++i) s += part[((size_t)c * S + i) * 2];
    out[c] = (float)s;
}

// ------------------------------------------------------------------------------------------------------------
// normalisation + activation + pooling block.  norm 0 = BatchNorm2d in train() mode (statistics per channel over
// (b,h,w), biased variance; running statistics updated with the unbiased one), norm 1 = GroupNorm(8) (statistics per
// (sample, group)).  act 0 = ReLU, 1 = SiLU.  pool (ph,pw) in {(1,1),(2,1),(2,2)} with floor semantics, or gap = 1
// (AdaptiveAvgPool2d(1)).  mu / rs are indexed by stat_idx(n, c).
// ------------------------------------------------------------------------------------------------------------
struct NormBlk {
    int B, C, H, W;
    int norm, act, ph, pw, gap;
    const float* gamma; const float* beta;
    float* mu; float* rs;
    float eps;
};

__device__ __forceinline__ int stat_idx(const NormBlk& k, int n, int c) { return k.norm == 0 ? c : n * 8 + c / (k.C / 8); }
__device__ __forceinline__ float act_fwd(int act, float z) { return act == 0 ? fmaxf(z, 0.f) : z / (1.f + __expf(-z)); }
__device__ __forceinline__ float act_bwd(int act, float z) {
    if (act == 0) return z > 0.f ? 1.f : 0.f;
    const float sg = 1.f / (1.f + __expf(-z));
    return sg * (1.f + z * (1.f - sg));
}

// statistics: one CTA per channel (BN) or per (sample, group) (GN); fp64 accumulation in a fixed order
__global__ void __launch_bounds__(256) norm_stats_kernel(const NormBlk k, const float* __restrict__ y, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float momentum) {
    const int HW = k.H * k.W;
    double s = 0.0, q = 0.0;
    long long count;
    if (k.norm == 0) {
        const int c = blockIdx.x;
        count = (long long)k.B * HW;
        for (long long i = threadIdx.x; i < count; i += 256) {
            const int b = (int)(i / HW), hw = (int)(i - (long long)b * HW);
            const double v = y[((size_t)b * k.C + c) * HW + hw];
            s += v; q += v * v;
        }
    } else {
        const int n = blockIdx.x >> 3, g = blockIdx.x & 7, cg = k.C / 8;
        count = (long long)cg * HW;
        const float* base = y + ((size_t)n * k.C + (size_t)g * cg) * HW;
        for (long long i = threadIdx.x; i < count; i += 256) {
            const double v = base[i];
            s += v; q += v * v;
        }
    }
    __shared__ double rs_[256], rq_[256];
    rs_[threadIdx.x] = s; rq_[threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { rs_[threadIdx.x] += rs_[threadIdx.x + o]; rq_[threadIdx.x] += rq_[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mean = rs_[0] / (double)count;
        double var = rq_[0] / (double)count - mean * mean;
        if (var < 0.0) var = 0.0;
        k.mu[blockIdx.x] = (float)mean;
        k.rs[blockIdx.x] = (float)(1.0 / sqrt(var + (double)k.eps));
        if (k.norm == 0 && running_mean && running_var) {        // nn.BatchNorm2d: running = (1-m)*running + m*stat, unbiased variance
            const double unb = count > 1 ? var * (double)count / (double)(count - 1) : var;
            running_mean[blockIdx.x] = (float)((1.0 - momentum) * running_mean[blockIdx.x] + momentum * mean);
            running_var[blockIdx.x] = (float)((1.0 - momentum) * running_var[blockIdx.x] + momentum * unb);
        }
    }
}

// BatchNorm statistics from the (C, S) partials of chan_partial_kernel: mean / rstd per channel + running-stat update
__global__ void bn_stats_final_kernel(const NormBlk k, const double* __restrict__ part, int S, float* __restrict__ running_mean,
                                      float* __restrict__ running_var, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k.C) return;
    double s = 0.0, q = 0.0;
    for (int i = 0; i < S; ++i) { s += part[((size_t)c * S + i) * 2]; q += part[((size_t)c * S + i) * 2 + 1]; }
    const double count = (double)k.B * k.H * k.W;
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    k.mu[c] = (float)mean;
    k.rs[c] = (float)(1.0 / sqrt(var + (double)k.eps));
    if (running_mean && running_var) {           // nn.BatchNorm2d: running = (1-m)*running + m*stat, unbiased variance
        const double unb = count > 1 ? var * count / (count - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
}

__device__ __forceinline__ float norm_z(const NormBlk& k, const float* __restrict__ y, int n, int c, int h, int w) {
    const int si = stat_idx(k, n, c);
    const float v = y[(((size_t)n * k.C + c) * k.H + h) * k.W + w];
    return fmaf(__ldg(k.gamma + c) * k.rs[si], v - k.mu[si], __ldg(k.beta + c));
}

// out[n,c,ho,wo] = max over the pool window of act(z)   (or the mean over (h,w) when gap)
__global__ void __launch_bounds__(256) norm_act_pool_fwd_kernel(const NormBlk k, const float* __restrict__ y, float* __restrict__ out) {
    if (k.gap) {
        // one warp per (n, c)
        const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (warp >= k.B * k.C) return;
        const int n = warp / k.C, c = warp - n * k.C;
        const int HW = k.H * k.W;
        float s = 0.f;
        for (int i = lane; i < HW; i += 32) s += act_fwd(k.act, norm_z(k, y, n, c, i / k.W, i % k.W));
        s = warp_sum(s);
        if (lane == 0) out[warp] = s / (float)HW;
        return;
    }
    const int Ho = k.H / k.ph, Wo = k.W / k.pw;
    const long long total = (long long)k.B * k.C * Ho * Wo;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int wo = (int)(i % Wo);
        long long r = i / Wo;
        const int ho = (int)(r % Ho); r /= Ho;
        const int c = (int)(r % k.C), n = (int)(r / k.C);
        float best = -INFINITY;
        for (int a = 0; a < k.ph; ++a)
            for (int b = 0; b < k.pw; ++b) best = fmaxf(best, act_fwd(k.act, norm_z(k, y, n, c, ho * k.ph + a, wo * k.pw + b)));
        out[i] = best;
    }
}

// dz[n,c,h,w] = upstream routed through the pool (first maximum of the window, as MaxPool2d) times act'(z), written into dy;
// pb[n,c] = sum_hw dz, pg[n,c] = sum_hw dz * xhat.  One warp per (n, c).
__global__ void __launch_bounds__(256) norm_bwd_reduce_kernel(const NormBlk k, const float* __restrict__ y, const float* __restrict__ up,
                                                              float* __restrict__ dy, float* __restrict__ pb, float* __restrict__ pg) {
    const int warp = (blockIdx.x * 256 + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= k.B * k.C) return;
    const int n = warp / k.C, c = warp - n * k.C;
    const int HW = k.H * k.W;
    const int si = stat_idx(k, n, c);
    const float mu = k.mu[si], rs = k.rs[si], ga = __ldg(k.gamma + c), be = __ldg(k.beta + c);
    const float* yc = y + ((size_t)n * k.C + c) * HW;
    float* dc = dy + ((size_t)n * k.C + c) * HW;
    const int Ho = k.gap ? 1 : k.H / k.ph, Wo = k.gap ? 1 : k.W / k.pw;
    const float* upc = up + ((size_t)n * k.C + c) * Ho * Wo;
    float sb = 0.f, sg = 0.f;
    for (int i = lane; i < HW; i += 32) {
        const int h = i / k.W, w = i - h * k.W;
        const float xh = (yc[i] - mu) * rs;
        const float z = fmaf(ga, xh, be);
        float u;
        if (k.gap) {
            u = upc[0] / (float)HW;
        } else {
            const int ho = h / k.ph, wo = w / k.pw;
            u = 0.f;
            if (ho < Ho && wo < Wo) {
                // is (h,w) the first maximum of its window?
                const float mine = act_fwd(k.act, z);
                bool is_max = true;
                for (int a = 0; a < k.ph; ++a)
                    for (int b = 0; b < k.pw; ++b) {
                        const int hh = ho * k.ph + a, ww = wo * k.pw + b;
                        if (hh == h && ww == w) continue;
                        const float other = act_fwd(k.act, fmaf(ga, (yc[hh * k.W + ww] - mu) * rs, be));
                        const bool before = (hh < h) || (hh == h && ww < w);
                        if (other > mine || (before && other == mine)) is_max = false;
                    }
                if (is_max) u = upc[ho * Wo + wo];
            }
        }
        const float dz = u * act_bwd(k.act, z);
        dc[i] = dz;
        sb += dz; sg += dz * xh;
    }
    sb = warp_sum(sb); sg = warp_sum(sg);
    if (lane == 0) { pb[warp] = sb; pg[warp] = sg; }
}

// dgamma[c] = sum_n pg[n,c], dbeta[c] = sum_n pb[n,c]; m1 / m2 = the two means the input gradient needs:
// BN: per channel  m1[c] = gamma*dbeta/(B*HW), m2[c] = gamma*dgamma/(B*HW);
// GN: per (n,g)    m1 = sum_{c in g} gamma_c pb[n,c] / (cg*HW), m2 likewise with pg.
__global__ void __launch_bounds__(256) norm_bwd_finalize_kernel(const NormBlk k, const float* __restrict__ pb, const float* __restrict__ pg,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                float* __restrict__ m1, float* __restrict__ m2) {
    const int HW = k.H * k.W;
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t < k.C) {
        double sg = 0.0, sb = 0.0;
        for (int n = 0; n < k.B; ++n) { sg += pg[(size_t)n * k.C + t]; sb += pb[(size_t)n * k.C + t]; }
        dgamma[t] = (float)sg; dbeta[t] = (float)sb;
        if (k.norm == 0) {
            const double inv = 1.0 / ((double)k.B * HW);
            m1[t] = (float)(k.gamma[t] * sb * inv);
            m2[t] = (float)(k.gamma[t] * sg * inv);
        }
    }
    if (k.norm == 1 && t < k.B * 8) {
        const int n = t >> 3, g = t & 7, cg = k.C / 8;
        double a = 0.0, b = 0.0;
        for (int c = g * cg; c < (g + 1) * cg; ++c) {
            a += (double)k.gamma[c] * pb[(size_t)n * k.C + c];
            b += (double)k.gamma[c] * pg[(size_t)n * k.C + c];
        }
        const double inv = 1.0 / ((double)cg * HW);
        m1[t] = (float)(a * inv); m2[t] = (float)(b * inv);
    }
}

// in place: dy = rs * (gamma_c * dz - m1 - xhat * m2)
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const NormBlk k, const float* __restrict__ y, float* __restrict__ dy,
                                                             const float* __restrict__ m1, const float* __restrict__ m2) {
    const int HW = k.H * k.W;
    const long long total = (long long)k.B * k.C * HW;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const long long r = i / HW;
        const int c = (int)(r % k.C), n = (int)(r / k.C);
        const int si = stat_idx(k, n, c);
        const float rs = k.rs[si];
        const float xh = (y[i] - k.mu[si]) * rs;
        dy[i] = rs * (__ldg(k.gamma + c) * dy[i] - m1[si] - xh * m2[si]);
    }
}

// dense head: v = dropout(act(u)), keep-mask uint8 (1 = keep), kept values scaled by 1/(1-p)
__global__ void head_act_drop_fwd_kernel(const float* __restrict__ u, const uint8_t* __restrict__ mask, float scale, int act, int n,
                                         float* __restrict__ v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float a = act_fwd(act, u[i]);
        v[i] = mask ? (mask[i] ? a * scale : 0.f) : a;
    }
}
__global__ void head_act_drop_bwd_kernel(const float* __restrict__ dv, const float* __restrict__ u, const uint8_t* __restrict__ mask,
                                         float scale, int act, int n, float* __restrict__ du) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float g = mask ? (mask[i] ? dv[i] * scale : 0.f) : dv[i];
        du[i] = g * act_bwd(act, u[i]);
    }
}
// out[n] = sum_m src[m][n], fixed order
__global__ void col_sum_kernel(const float* __restrict__ src, int M, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double s = 0.0;
    for (int m = 0; m < M; ++m) s += src[(size_t)m * N + n];
    out[n] = (float)s;
}

// Loss and d loss / d logits for [B,2] logits.  alpha == NULL && gamma == 0: nn.CrossEntropyLoss (mean), 05_train_cnn.py:257.
// Otherwise WeightedFocalLoss (06_train_cnn.py:195-207): ce_i = CE(logits_i, t_i); pt = exp(-ce); loss = mean(alpha[t] (1-pt)^gamma ce).
__global__ void __launch_bounds__(256) cnn_loss_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int B,
                                                            const float* __restrict__ alpha, float gamma, float* __restrict__ dlogits,
                                                            float* __restrict__ loss) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) {
        const float l0 = logits[2 * i], l1 = logits[2 * i + 1];
        const int t = (int)target[i];
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        const float lse = mx + logf(e0 + e1);
        const float p0 = e0 / (e0 + e1), p1 = e1 / (e0 + e1);
        const float ce = lse - (t ? l1 : l0);
        float li = ce, dce = 1.f;
        if (alpha || gamma != 0.f) {
            const float at = alpha ? alpha[t] : 1.f;
            const float pt = expf(-ce);
            const float om = 1.f - pt;
            const float f = powf(fmaxf(om, 0.f), gamma);
            li = at * f * ce;
            const float df = gamma == 0.f ? 0.f : gamma * powf(fmaxf(om, 0.f), gamma - 1.f) * pt;     // d (1-pt)^gamma / d ce
            dce = at * (f + df * ce);
        }
        acc += (double)li;
        if (dlogits) {
            const float s = dce / (float)B;
            dlogits[2 * i] = s * (p0 - (t == 0 ? 1.f : 0.f));
            dlogits[2 * i + 1] = s * (p1 - (t == 1 ? 1.f : 0.f));
        }
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && loss) loss[0] = (float)(red[0] / (double)B);
}

// ------------------------------------------------------------------------------------------------------------
// architecture tables
// ------------------------------------------------------------------------------------------------------------
struct BlkSpec { int Cin, Cout, H, W, KH, KW, PH, PW, norm, act, ph, pw, gap; };
struct ArchSpec {
    int n_blk;
    BlkSpec blk[4];
    int in_C, in_H, in_W;
    int F, HID, NCLS, head_act;
};

static ArchSpec arch_spec(int arch) {
    ArchSpec a;
    memset(&a, 0, sizeof(a));
    if (arch == SHM_CNN_4DOF) {               // cnn_model.py:16-34
        a.n_blk = 2; a.in_C = 2; a.in_H = 100; a.in_W = 12;
        a.blk[0] = {2, 16, 100, 12, 3, 3, 1, 1, 0, 0, 2, 2, 0};
        a.blk[1] = {16, 32, 50, 6, 3, 3, 1, 1, 0, 0, 2, 2, 0};
        a.F = 32 * 25 * 3; a.HID = 128; a.NCLS = 2; a.head_act = 0;
    } else {                                  // openLAB Models/cnn_model.py:16-43
        a.n_blk = 4; a.in_C = 1; a.in_H = 200; a.in_W = 4;
        a.blk[0] = {1, 32, 200, 4, 7, 3, 3, 1, 1, 1, 2, 1, 0};
        a.blk[1] = {32, 64, 100, 4, 5, 3, 2, 1, 1, 1, 2, 1, 0};
        a.blk[2] = {64, 128, 50, 4, 5, 3, 2, 1, 1, 1, 2, 1, 0};
        a.blk[3] = {128, 256, 25, 4, 3, 3, 1, 1, 1, 1, 1, 1, 1};
        a.F = 256; a.HID = 128; a.NCLS = 2; a.head_act = 1;
    }
    return a;
}

struct CnnLayout {        // list(model.parameters()) order: per block conv.weight, conv.bias, norm.weight, norm.bias; fc1 w,b; fc2 w,b
    size_t cw[4], cb[4], nw[4], nb[4], f1w, f1b, f2w, f2b, total;
};

static CnnLayout cnn_layout(const ArchSpec& a) {
    CnnLayout L;
    memset(&L, 0, sizeof(L));
    size_t o = 0;
    for (int i = 0; i < a.n_blk; ++i) {
        const BlkSpec& b = a.blk[i];
        L.cw[i] = o; o += (size_t)b.Cout * b.Cin * b.KH * b.KW;
        L.cb[i] = o; o += b.Cout;
        L.nw[i] = o; o += b.Cout;
        L.nb[i] = o; o += b.Cout;
    }
    L.f1w = o; o += (size_t)a.HID * a.F; L.f1b = o; o += a.HID;
    L.f2w = o; o += (size_t)a.NCLS * a.HID; L.f2b = o; o += a.NCLS;
    L.total = o;
    return L;
}

}  // namespace shm

using namespace shm;

struct shm_cnn_trainer {
    int arch, Bmax, device, nsm;
    ArchSpec a;
    CnnLayout L;
    float* ws;
    size_t ws_floats;
    // workspace offsets (floats)
    size_t y[4], act[4], mu[4], rs[4], wf[4], wb[4];
    size_t u, vd, dlg_u, dvd, dfeat, dy, da, pb, pg, m1, m2, part;
    uint8_t* mask;
    int B, have_fwd, use_mask;
    float drop_scale;
    const float* x_in;       // the caller's input of the last forward (must stay alive until backward)
};

extern "C" int64_t shm_cnn_param_count(int arch) {
    if (arch != SHM_CNN_4DOF && arch != SHM_CNN_OPENLAB) return SHM_ERR_ARG;
    return (int64_t)cnn_layout(arch_spec(arch)).total;
}

extern "C" int shm_cnn_trainer_create(shm_cnn_trainer** out, int arch, int32_t max_batch, int device) {
    if (!out || max_batch < 1 || (arch != SHM_CNN_4DOF && arch != SHM_CNN_OPENLAB)) return SHM_ERR_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != SHM_OK) return rc;
    int prev = 0;
    SHM_CUDA(cudaGetDevice(&prev));
    SHM_CUDA(cudaSetDevice(device));
    shm_cnn_trainer* h = new (std::nothrow) shm_cnn_trainer();
    if (!h) { cudaSetDevice(prev); return SHM_ERR_NOMEM; }
    memset(h, 0, sizeof(*h));
    h->arch = arch; h->Bmax = max_batch; h->device = device; h->nsm = device_sm_count(device);
    h->a = arch_spec(arch); h->L = cnn_layout(h->a);
    const size_t B = max_batch;
    size_t o = 0, max_y = 0, max_a = 0, max_c = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };
    for (int i = 0; i < h->a.n_blk; ++i) {
        const BlkSpec& b = h->a.blk[i];
        const size_t ysz = B * b.Cout * b.H * b.W;
        const size_t asz = b.gap ? B * b.Cout : B * b.Cout * (b.H / b.ph) * (b.W / b.pw);
        h->y[i] = take(ysz); h->act[i] = take(asz);
        const size_t ns = b.norm == 0 ? b.Cout : B * 8;
        h->mu[i] = take(ns); h->rs[i] = take(ns);
        const size_t wsz = (size_t)b.Cout * b.Cin * b.KH * b.KW;
        h->wf[i] = take(wsz); h->wb[i] = take(wsz);
        max_y = ysz > max_y ? ysz : max_y; max_a = asz > max_a ? asz : max_a;
        max_c = (size_t)b.Cout > max_c ? b.Cout : max_c;
    }
    const size_t in_sz = B * h->a.in_C * h->a.in_H * h->a.in_W;
    max_a = in_sz > max_a ? in_sz : max_a;
    h->u = take(B * h->a.HID); h->vd = take(B * h->a.HID); h->dlg_u = take(B * h->a.HID); h->dvd = take(B * h->a.HID);
    h->dfeat = take(B * h->a.F);
    h->dy = take(max_y); h->da = take(max_a);
    h->pb = take(B * max_c); h->pg = take(B * max_c);
    h->m1 = take(B * 8 > max_c ? B * 8 : max_c); h->m2 = take(B * 8 > max_c ? B * 8 : max_c);
    h->part = take((size_t)256 * 32 * 2 * 2);          // fp64 [C <= 256][S <= 32][2] partial sums
    h->ws_floats = o;
    if (cudaMalloc(&h->ws, o * sizeof(float)) != cudaSuccess || cudaMalloc(&h->mask, B * h->a.HID) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaMalloc(cnn trainer)");
        if (h->ws) cudaFree(h->ws);
        delete h;
        cudaSetDevice(prev);
        return SHM_ERR_NOMEM;
    }
    cudaSetDevice(prev);
    *out = h;
    return SHM_OK;
}

extern "C" int shm_cnn_trainer_destroy(shm_cnn_trainer* h) {
    if (!h) return SHM_OK;
    int prev = 0;
    const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
    cudaSetDevice(h->device);
    cudaFree(h->ws);
    cudaFree(h->mask);
    delete h;
    if (have_prev) cudaSetDevice(prev);
    return SHM_OK;
}

static NormBlk norm_blk(const shm_cnn_trainer* h, const float* params, int i, int B) {
    const BlkSpec& b = h->a.blk[i];
    NormBlk k;
    k.B = B; k.C = b.Cout; k.H = b.H; k.W = b.W; k.norm = b.norm; k.act = b.act; k.ph = b.ph; k.pw = b.pw; k.gap = b.gap;
    k.gamma = params + h->L.nw[i]; k.beta = params + h->L.nb[i];
    k.mu = h->ws + h->mu[i]; k.rs = h->ws + h->rs[i];
    k.eps = 1e-5f;
    return k;
}

static ConvG conv_g(const BlkSpec& b, int B) { return ConvG{B, b.Cin, b.Cout, b.H, b.W, b.KH, b.KW, b.PH, b.PW}; }

static inline int ew_grid(long long n, int nsm) {
    const long long want = (n + 255) / 256;
    const long long cap = (long long)nsm * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// batch splits for the per-channel reductions: enough CTAs to fill the GPU, at most 32 per channel
static inline void chan_splits(int B, int C, int nsm, int* S, int* bps) {
    int s = (2 * nsm + C - 1) / C;
    if (s > 32) s = 32;
    if (s > B) s = B;
    if (s < 1) s = 1;
    *bps = (B + s - 1) / s;
    *S = (B + *bps - 1) / *bps;
}

extern "C" int shm_cnn_train_forward(shm_cnn_trainer* h, const float* params, const float* x, int32_t B, float* bn_running,
                                     float bn_momentum, const uint8_t* drop_mask, float drop_p, float* logits, void* stream) {
    if (!h || !params || !x || !logits || B < 1 || B > h->Bmax || drop_p < 0.f || drop_p >= 1.f) return SHM_ERR_ARG;
    if (h->arch == SHM_CNN_4DOF && B < 2) return SHM_ERR_ARG;        // BatchNorm in train mode needs more than one value per channel
    cudaStream_t st = (cudaStream_t)stream;
    const ArchSpec& a = h->a;
    float* ws = h->ws;
    const float* in = x;
    size_t run_off = 0;
    for (int i = 0; i < a.n_blk; ++i) {
        const BlkSpec& b = a.blk[i];
        const int KHW = b.KH * b.KW;
        conv_repack_kernel<<<ew_grid((long long)b.Cout * b.Cin * KHW, h->nsm), 256, 0, st>>>(params + h->L.cw[i], b.Cout, b.Cin, KHW,
                                                                                           ws + h->wf[i], ws + h->wb[i]);
        SHM_LAUNCH_CHECK();
        const ConvG g = conv_g(b, B);
        dim3 grid((unsigned)(((long long)B * b.H * b.W + CI_BM - 1) / CI_BM), (unsigned)((b.Cout + CI_BN - 1) / CI_BN));
        conv_igemm_kernel<0><<<grid, 256, 0, st>>>(g, in, ws + h->wf[i], params + h->L.cb[i], ws + h->y[i]);
        SHM_LAUNCH_CHECK();
        const NormBlk k = norm_blk(h, params, i, B);
        float* rm = nullptr; float* rv = nullptr;
        if (b.norm == 0 && bn_running) { rm = bn_running + run_off; rv = bn_running + run_off + b.Cout; run_off += 2 * (size_t)b.Cout; }
        if (b.norm == 0) {
            int S, bps;
            chan_splits(B, b.Cout, h->nsm, &S, &bps);
            double* part = reinterpret_cast<double*>(ws + h->part);
            chan_partial_kernel<<<dim3(b.Cout, S), 256, 0, st>>>(ws + h->y[i], B, b.Cout, b.H * b.W, bps, part);
            SHM_LAUNCH_CHECK();
            bn_stats_final_kernel<<<(b.Cout + 63) / 64, 64, 0, st>>>(k, part, S, rm, rv, bn_momentum);
        } else {
            norm_stats_kernel<<<B * 8, 256, 0, st>>>(k, ws + h->y[i], rm, rv, bn_momentum);
        }
        SHM_LAUNCH_CHECK();
        const long long n_out = b.gap ? (long long)B * b.Cout * 32 : (long long)B * b.Cout * (b.H / b.ph) * (b.W / b.pw);
        norm_act_pool_fwd_kernel<<<b.gap ? (unsigned)((n_out + 255) / 256) : ew_grid(n_out, h->nsm), 256, 0, st>>>(k, ws + h->y[i], ws + h->act[i]);
        SHM_LAUNCH_CHECK();
        in = ws + h->act[i];
    }
    // dense head: u = feat W1^T + b1; vd = dropout(act(u)); logits = vd W2^T + b2
    int rc = sgemm(st, in, a.F, 1, params + h->L.f1w, 1, a.F, ws + h->u, a.HID, B, a.HID, a.F, params + h->L.f1b, nullptr, false);
    if (rc != SHM_OK) return rc;
    h->use_mask = drop_mask != nullptr && drop_p > 0.f;
    h->drop_scale = h->use_mask ? 1.f / (1.f - drop_p) : 1.f;
    if (h->use_mask) SHM_CUDA(cudaMemcpyAsync(h->mask, drop_mask, (size_t)B * a.HID, cudaMemcpyDeviceToDevice, st));
    head_act_drop_fwd_kernel<<<ew_grid((long long)B * a.HID, h->nsm), 256, 0, st>>>(ws + h->u, h->use_mask ? h->mask : nullptr, h->drop_scale,
                                                                                    a.head_act, B * a.HID, ws + h->vd);
    SHM_LAUNCH_CHECK();
    rc = sgemm(st, ws + h->vd, a.HID, 1, params + h->L.f2w, 1, a.HID, logits, a.NCLS, B, a.NCLS, a.HID, params + h->L.f2b, nullptr, false);
    if (rc != SHM_OK) return rc;
    h->B = B; h->have_fwd = 1; h->x_in = x;
    return SHM_OK;
}

extern "C" int shm_cnn_train_backward(shm_cnn_trainer* h, const float* params, const float* d_logits, float* grads, void* stream) {
    if (!h || !params || !d_logits || !grads) return SHM_ERR_ARG;
    if (!h->have_fwd) return SHM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const ArchSpec& a = h->a;
    const CnnLayout& L = h->L;
    float* ws = h->ws;
    const int B = h->B;
    SHM_CUDA(cudaMemsetAsync(grads, 0, L.total * sizeof(float), st));
    const float* feat = ws + h->act[a.n_blk - 1];
    int rc;
    // fc2: dW2 = dlogits^T vd, db2 = colsum(dlogits), dvd = dlogits W2
    if ((rc = sgemm(st, d_logits, 1, a.NCLS, ws + h->vd, a.HID, 1, grads + L.f2w, a.HID, a.NCLS, a.HID, B, nullptr, nullptr, false))) return rc;
    col_sum_kernel<<<1, 32, 0, st>>>(d_logits, B, a.NCLS, grads + L.f2b);
    SHM_LAUNCH_CHECK();
    if ((rc = sgemm(st, d_logits, a.NCLS, 1, params + L.f2w, a.HID, 1, ws + h->dvd, a.HID, B, a.HID, a.NCLS, nullptr, nullptr, false))) return rc;
    head_act_drop_bwd_kernel<<<ew_grid((long long)B * a.HID, h->nsm), 256, 0, st>>>(ws + h->dvd, ws + h->u, h->use_mask ? h->mask : nullptr,
                                                                                    h->drop_scale, a.head_act, B * a.HID, ws + h->dlg_u);
    SHM_LAUNCH_CHECK();
    // fc1: dW1 = du^T feat, db1 = colsum(du), dfeat = du W1
    if ((rc = sgemm(st, ws + h->dlg_u, 1, a.HID, feat, a.F, 1, grads + L.f1w, a.F, a.HID, a.F, B, nullptr, nullptr, false))) return rc;
    col_sum_kernel<<<(a.HID + 127) / 128, 128, 0, st>>>(ws + h->dlg_u, B, a.HID, grads + L.f1b);
    SHM_LAUNCH_CHECK();
    if ((rc = sgemm(st, ws + h->dlg_u, a.HID, 1, params + L.f1w, a.F, 1, ws + h->dfeat, a.F, B, a.F, a.HID, nullptr, nullptr, false))) return rc;
    const float* up = ws + h->dfeat;
    for (int i = a.n_blk - 1; i >= 0; --i) {
        const BlkSpec& b = a.blk[i];
        const NormBlk k = norm_blk(h, params, i, B);
        const long long nw = (long long)B * b.Cout * 32;
        norm_bwd_reduce_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(k, ws + h->y[i], up, ws + h->dy, ws + h->pb, ws + h->pg);
        SHM_LAUNCH_CHECK();
        const int nf = b.norm == 0 ? b.Cout : (B * 8 > b.Cout ? B * 8 : b.Cout);
        norm_bwd_finalize_kernel<<<(nf + 255) / 256, 256, 0, st>>>(k, ws + h->pb, ws + h->pg, grads + L.nw[i], grads + L.nb[i], ws + h->m1,
                                                                  ws + h->m2);
        SHM_LAUNCH_CHECK();
        norm_bwd_apply_kernel<<<ew_grid((long long)B * b.Cout * b.H * b.W, h->nsm), 256, 0, st>>>(k, ws + h->y[i], ws + h->dy, ws + h->m1, ws + h->m2);
        SHM_LAUNCH_CHECK();
        // convolution: bias, weight, (data)
        {
            int S, bps;
            chan_splits(B, b.Cout, h->nsm, &S, &bps);
            double* part = reinterpret_cast<double*>(ws + h->part);
            chan_partial_kernel<<<dim3(b.Cout, S), 256, 0, st>>>(ws + h->dy, B, b.Cout, b.H * b.W, bps, part);
            SHM_LAUNCH_CHECK();
            chan_sum_final_kernel<<<(b.Cout + 63) / 64, 64, 0, st>>>(part, b.Cout, S, grads + L.cb[i]);
            SHM_LAUNCH_CHECK();
        }
        const ConvG g = conv_g(b, B);
        const float* xin = i == 0 ? h->x_in : ws + h->act[i - 1];
        const int Nw = b.Cin * b.KH * b.KW;
        const int tiles = ((b.Cout + CT_BM - 1) / CT_BM) * ((Nw + CT_BN - 1) / CT_BN);
        int splits = (2 * h->nsm + tiles - 1) / tiles;
        if (splits > B) splits = B;
        if (splits < 1) splits = 1;
        const int bps = (B + splits - 1) / splits;
        splits = (B + bps - 1) / bps;
        dim3 gw((unsigned)((b.Cout + CT_BM - 1) / CT_BM), (unsigned)((Nw + CT_BN - 1) / CT_BN), (unsigned)splits);
        conv_wgrad_kernel<<<gw, 256, 0, st>>>(g, ws + h->dy, xin, grads + L.cw[i], bps);
        SHM_LAUNCH_CHECK();
        if (i > 0) {
            dim3 gd((unsigned)(((long long)B * b.H * b.W + CI_BM - 1) / CI_BM), (unsigned)((b.Cin + CI_BN - 1) / CI_BN));
            conv_igemm_kernel<1><<<gd, 256, 0, st>>>(g, ws + h->dy, ws + h->wb[i], nullptr, ws + h->da);
            SHM_LAUNCH_CHECK();
            up = ws + h->da;
        }
    }
    h->have_fwd = 0;
    return SHM_OK;
}

extern "C" int shm_cnn_loss_grad(const float* logits, const int64_t* targets, int64_t B, const float* alpha, float gamma,
                                 float* d_logits, float* loss, void* stream) {
    if (!logits || !targets || B < 1 || B > (1 << 24) || gamma < 0.f) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cnn_loss_grad_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, reinterpret_cast<const long long*>(targets), (int)B, alpha, gamma,
                                                             d_logits, loss);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
