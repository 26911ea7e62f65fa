// Fused LSTM-VAE scorer, fp32 FMA engine (SHM_ENGINE_FP32).
//
// One CTA owns a tile of BM windows and carries them through the WHOLE model without touching HBM
// in between: encoder LSTM stack over T steps -> LayerNorm -> mu/logvar heads -> reparameterise ->
// tanh(latent->hidden) -> decoder LSTM stack -> output Linear -> squared error accumulated per window.
// Replaces TemporalVAE.forward (4DOF/Scripts/Models/temporal_vae.py:51-77 and the 1_DOF / openLAB
// twins) plus the scoring loops (04_vae_thresholding.py:113-124, 10_test_hybrid_pipeline.py:240-251).
//
// Data layout on chip (all fp32, shared memory, "k-major": row = feature, column = window):
//   hT[l][H][BM], cT[l][H][BM]  recurrent state, uT[H][BM] decoder input tanh(fc(z)),
//   xT[TCH][Dpad][BM] a chunk of TCH timesteps of the (gathered + normalised) windows.
// Per (layer, step) the gate pre-activations are a [BM x K] x [K x 4H] product with
// K = Kin_pad + H ([input | recurrent] concatenated).  Weights are repacked once on the device into
// Wp[k][j*H + ug*4 + gate] (unit u = 4*ug + j) and streamed L2 -> smem in KT-row tiles with a
// double-buffered cp.async pipeline; each thread keeps a 4-window x 4-unit x 4-gate register tile so
// the LSTM cell update is thread-local.  Warp = 8 window-groups x 4 unit-groups: the A read is one
// 128-byte wavefront, each B read a 64-byte broadcast wavefront (5 LDS.128 per 64 FFMA).
#pragma once
#include "common.cuh"

namespace shm {

constexpr int VAE_KT = 8;      // weight rows per streamed tile
constexpr int VAE_TCH = 4;     // timesteps of x staged per chunk
constexpr int VAE_MAX_Z = 16;

struct VaeDev {
    int D, H, Z, L, has_ln, Dpad;
    float ln_eps;
    // packed [K][4H] per phase, layers concatenated row-wise; K_l = Kin_pad_l + H
    const float* encW; const float* decW;
    const float* encB[SHM_MAX_L]; const float* decB[SHM_MAX_L];     // packed biases b_ih + b_hh
    int encK[SHM_MAX_L], decK[SHM_MAX_L];
    int encKin[SHM_MAX_L], decKin[SHM_MAX_L];                        // padded input widths
    const float *ln_w, *ln_b, *mu_w, *mu_b, *lv_w, *lv_b, *l2h_w, *l2h_b, *out_w, *out_b;
};

struct VaeIO {
    const int* idx; const int* n_dev; const float* eps;
    const float* z_in;          // decode-only entry (TemporalVAE.decode): latent supplied, encoder skipped
    const float* mu_in;         // re-score entry: mu / logvar of an earlier pass, [all windows, Z], indexed like the windows (idx);
    const float* logvar_in;     //   the (deterministic) encoder is skipped, z = mu + eps * exp(0.5 logvar) with this call's eps
    long long n;
    float *score, *mu, *logvar, *recon, *cnn_in;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return 1.0f - 2.0f / (__expf(2.0f * x) + 1.0f); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int H, int BM>
struct VaeSmem {
    static constexpr int NW = (BM / 32) * (H / 16);
    static constexpr int NT = NW * 32;
    static constexpr int WTILE = VAE_KT * 4 * H;                    // floats per weight tile
    static constexpr int off_h = 0;
    static constexpr int off_c = off_h + SHM_MAX_L * H * BM;
    static constexpr int off_w = off_c + SHM_MAX_L * H * BM;
    static constexpr int off_x = off_w + 2 * WTILE;                 // xT, aliased with head scratch
    static constexpr int XT = VAE_TCH * 16 * BM;
    static constexpr int SCR = (2 * VAE_MAX_Z + VAE_MAX_Z) * BM;    // muS/lvS + zS
    static constexpr int off_u = off_x + (XT > SCR ? XT : SCR);
    static constexpr int off_ow = off_u + H * BM;
    static constexpr int total = off_ow + SHM_MAX_D * H + SHM_MAX_D;
    static constexpr size_t bytes = (size_t)total * sizeof(float);
    static constexpr int MAXI = (BM * SHM_MAX_D + NT - 1) / NT;     // (window, channel) items per thread
};

// One LSTM phase (encoder or decoder stack) over T steps for the CTA's tile.
//   IS_DEC = false: layer-0 input = x_t (from xT);  true: layer-0 input = uT, and the top layer's
//   h_t goes through the output Linear and the squared-error accumulation.
template <int H, int BM, bool IS_DEC>
__device__ __forceinline__ void lstm_phase(const VaeDev& P, const WinSrc& src, const VaeIO& io, float* sm,
                                           long long n0, int nvalid, int T, float* sse) {
    using S = VaeSmem<H, BM>;
    constexpr int NT = S::NT;
    constexpr int G4 = 4 * H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int warp_w = warp % (BM / 32), warp_u = warp / (BM / 32);
    const int w0 = (warp_w * 8 + (lane & 7)) * 4;            // first of this thread's 4 windows
    const int ug = warp_u * 4 + (lane >> 3);                 // unit group: units 4*ug .. 4*ug+3

    float* hT = sm + S::off_h;
    float* cT = sm + S::off_c;
    float* wbuf = sm + S::off_w;
    float* xT = sm + S::off_x;
    float* uT = sm + S::off_u;
    float* outW = sm + S::off_ow;
    float* outB = outW + SHM_MAX_D * H;

    const float* gW = IS_DEC ? P.decW : P.encW;
    int Ktot = 0;
    for (int l = 0; l < P.L; ++l) Ktot += IS_DEC ? P.decK[l] : P.encK[l];
    const int S_tiles = Ktot / VAE_KT;

    // zero recurrent state (both LSTMs start from zeros: nn.LSTM default, temporal_vae.py:53,68)
    for (int i = tid; i < 2 * SHM_MAX_L * H * BM; i += NT) sm[S::off_h + i] = 0.f;

    auto prefetch = [&](int s, int buf) {
        const float* g = gW + (size_t)s * S::WTILE;
        float* d = wbuf + buf * S::WTILE;
#pragma unroll
        for (int i = tid * 4; i < S::WTILE; i += NT * 4) cp_async16(d + i, g + i);
        cp_async_commit();
    };

    int cur = 0, s_next = 0;
    prefetch(0, 0);
    s_next = 1 % S_tiles;
    __syncthreads();

    for (int t = 0; t < T; ++t) {
        const int tl = t % VAE_TCH;
        if (tl == 0) {
            __syncthreads();        // everyone is done with the previous chunk (decoder error pass reads it)
            // stage TCH timesteps of the windows: gather + normalise fused here, nothing materialised
            const int nt = min(VAE_TCH, T - t);
            const int per_w = nt * P.D;
            for (int i = tid; i < BM * per_w; i += NT) {
                const int w = i / per_w;
                const int r = i - w * per_w;
                const int tt = r / P.D;
                const int d = r - tt * P.D;
                float v = 0.f;
                if (w < nvalid && src.base) {
                    const long long n = n0 + w;
                    const long long win = io.idx ? (long long)io.idx[n] : n;
                    v = win_fetch(src, win, t + tt, d);
                }
                xT[(tt * 16 + d) * BM + w] = v;
            }
            // visibility is ordered by the __syncthreads of the first weight tile below
        }

        for (int l = 0; l < P.L; ++l) {
            const int K = IS_DEC ? P.decK[l] : P.encK[l];
            const int Kin = IS_DEC ? P.decKin[l] : P.encKin[l];
            const float* bias = IS_DEC ? P.decB[l] : P.encB[l];
            const float* inT = (l == 0) ? (IS_DEC ? uT : xT + tl * 16 * BM) : hT + (l - 1) * H * BM;
            float* hl = hT + l * H * BM;
            float* cl = cT + l * H * BM;

            float acc[4][4][4];                                   // [window][unit j][gate]
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias + j * H + ug * 4));
#pragma unroll
                for (int w = 0; w < 4; ++w) { acc[w][j][0] = b.x; acc[w][j][1] = b.y; acc[w][j][2] = b.z; acc[w][j][3] = b.w; }
            }

            for (int k0 = 0; k0 < K; k0 += VAE_KT) {
                cp_async_wait_all();
                __syncthreads();
                prefetch(s_next, cur ^ 1);
                s_next = (s_next + 1 == S_tiles) ? 0 : s_next + 1;
                const float* wt = wbuf + cur * S::WTILE;
                const float* aT = (k0 < Kin) ? inT + k0 * BM : hl + (k0 - Kin) * BM;
#pragma unroll
                for (int kk = 0; kk < VAE_KT; ++kk) {
                    const float4 a = *reinterpret_cast<const float4*>(aT + kk * BM + w0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b = *reinterpret_cast<const float4*>(wt + kk * G4 + j * H + ug * 4);
                        acc[0][j][0] = fmaf(a.x, b.x, acc[0][j][0]); acc[0][j][1] = fmaf(a.x, b.y, acc[0][j][1]);
                        acc[0][j][2] = fmaf(a.x, b.z, acc[0][j][2]); acc[0][j][3] = fmaf(a.x, b.w, acc[0][j][3]);
                        acc[1][j][0] = fmaf(a.y, b.x, acc[1][j][0]); acc[1][j][1] = fmaf(a.y, b.y, acc[1][j][1]);
                        acc[1][j][2] = fmaf(a.y, b.z, acc[1][j][2]); acc[1][j][3] = fmaf(a.y, b.w, acc[1][j][3]);
                        acc[2][j][0] = fmaf(a.z, b.x, acc[2][j][0]); acc[2][j][1] = fmaf(a.z, b.y, acc[2][j][1]);
                        acc[2][j][2] = fmaf(a.z, b.z, acc[2][j][2]); acc[2][j][3] = fmaf(a.z, b.w, acc[2][j][3]);
                        acc[3][j][0] = fmaf(a.w, b.x, acc[3][j][0]); acc[3][j][1] = fmaf(a.w, b.y, acc[3][j][1]);
                        acc[3][j][2] = fmaf(a.w, b.z, acc[3][j][2]); acc[3][j][3] = fmaf(a.w, b.w, acc[3][j][3]);
                    }
                }
                cur ^= 1;
            }
            __syncthreads();        // every thread is done reading hl before it is overwritten

            // LSTM cell, gate order i,f,g,o; c = f*c + i*g; h = o*tanh(c)   (thread-local)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = ug * 4 + j;
                float4 c4 = *reinterpret_cast<float4*>(cl + u * BM + w0);
                float cc[4] = {c4.x, c4.y, c4.z, c4.w};
                float hh[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float ig = sigmoid_f(acc[w][j][0]);
                    const float fg = sigmoid_f(acc[w][j][1]);
                    const float gg = tanh_f(acc[w][j][2]);
                    const float og = sigmoid_f(acc[w][j][3]);
                    cc[w] = fmaf(fg, cc[w], ig * gg);
                    hh[w] = og * tanh_f(cc[w]);
                }
                *reinterpret_cast<float4*>(cl + u * BM + w0) = make_float4(cc[0], cc[1], cc[2], cc[3]);
                *reinterpret_cast<float4*>(hl + u * BM + w0) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            }
        }

        if (IS_DEC) {
            __syncthreads();        // top-layer h_t visible
            const float* htop = hT + (P.L - 1) * H * BM;
            const float* xrow = xT + tl * 16 * BM;
#pragma unroll
            for (int it = 0; it < S::MAXI; ++it) {
                const int item = tid + it * NT;
                if (item >= BM * P.D) break;
                const int d = item / BM;
                const int w = item - d * BM;
                float y = outB[d];
                const float* wo = outW + d * H;
#pragma unroll 8
                for (int k = 0; k < H; ++k) y = fmaf(htop[k * BM + w], wo[k], y);
                const float x = xrow[d * BM + w];
                const float e = x - y;
                sse[it] = fmaf(e, e, sse[it]);
                if (w < nvalid) {
                    const long long n = n0 + w;
                    if (io.recon) io.recon[(n * T + t) * P.D + d] = y;
                    if (io.cnn_in) {
                        io.cnn_in[((n * 2 + 0) * T + t) * P.D + d] = x;
                        io.cnn_in[((n * 2 + 1) * T + t) * P.D + d] = e * e;
                    }
                }
            }
            // the next step's first-tile __syncthreads orders these reads before hT is rewritten
        }
    }
    cp_async_wait_all();
    __syncthreads();
}

template <int H, int BM>
__global__ void __launch_bounds__(VaeSmem<H, BM>::NT, 1)
vae_score_fp32_kernel(VaeDev P, WinSrc src, VaeIO io) {
    using S = VaeSmem<H, BM>;
    constexpr int NT = S::NT;
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x;
    long long n_eff = io.n;
    if (io.n_dev) n_eff = min(n_eff, (long long)__ldg(io.n_dev));
    const long long n0 = (long long)blockIdx.x * BM;
    if (n0 >= n_eff) return;
    const int nvalid = (int)min((long long)BM, n_eff - n0);
    const int T = src.T;

    float* hT = sm + S::off_h;
    float* xT = sm + S::off_x;
    float* uT = sm + S::off_u;
    float* outW = sm + S::off_ow;
    float* outB = outW + SHM_MAX_D * H;

    for (int i = tid; i < S::XT; i += NT) xT[i] = 0.f;       // rows d >= D stay zero (K padding)
    for (int i = tid; i < P.D * H; i += NT) outW[i] = __ldg(P.out_w + i);
    if (tid < P.D) outB[tid] = __ldg(P.out_b + tid);
    __syncthreads();

    float sse[S::MAXI];
#pragma unroll
    for (int i = 0; i < S::MAXI; ++i) sse[i] = 0.f;

    float* hlast = hT + (P.L - 1) * H * BM;
    float* muS = xT;                          // [2Z][BM]  (xT is free between the phases)
    float* zS = xT + 2 * VAE_MAX_Z * BM;      // [Z][BM]
    if (io.z_in) {
        for (int item = tid; item < P.Z * BM; item += NT) {
            const int zi = item / BM, w = item - zi * BM;
            zS[item] = (w < nvalid) ? __ldg(io.z_in + (n0 + w) * P.Z + zi) : 0.f;
        }
        __syncthreads();
    } else {
    // ---------------- encoder ----------------
    lstm_phase<H, BM, false>(P, src, io, sm, n0, nvalid, T, sse);

    // ---------------- LayerNorm, heads, reparameterisation, latent->hidden ----------------
    if (P.has_ln) {
        if (tid < BM) {
            float m = 0.f;
            for (int k = 0; k < H; ++k) m += hlast[k * BM + tid];
            m /= (float)H;
            float v = 0.f;
            for (int k = 0; k < H; ++k) { const float d = hlast[k * BM + tid] - m; v = fmaf(d, d, v); }
            v /= (float)H;
            const float rstd = 1.0f / sqrtf(v + P.ln_eps);
            for (int k = 0; k < H; ++k)
                hlast[k * BM + tid] = fmaf((hlast[k * BM + tid] - m) * rstd, __ldg(P.ln_w + k), __ldg(P.ln_b + k));
        }
        __syncthreads();
    }
    for (int item = tid; item < 2 * P.Z * BM; item += NT) {
        const int o = item / BM, w = item - o * BM;
        const bool is_lv = o >= P.Z;
        const int zi = is_lv ? o - P.Z : o;
        const float* wr = (is_lv ? P.lv_w : P.mu_w) + zi * H;
        float y = __ldg((is_lv ? P.lv_b : P.mu_b) + zi);
        for (int k = 0; k < H; ++k) y = fmaf(hlast[k * BM + w], __ldg(wr + k), y);
        muS[o * BM + w] = y;
        if (w < nvalid) {
            float* dst = is_lv ? io.logvar : io.mu;
            if (dst) dst[(n0 + w) * P.Z + zi] = y;
        }
    }
    __syncthreads();
    for (int item = tid; item < P.Z * BM; item += NT) {
        const int zi = item / BM, w = item - zi * BM;
        const float m = muS[zi * BM + w], lv = muS[(P.Z + zi) * BM + w];
        float z = m;
        if (io.eps && w < nvalid) z = fmaf(__ldg(io.eps + (n0 + w) * P.Z + zi), expf(0.5f * lv), m);
        zS[zi * BM + w] = z;
    }
    __syncthreads();
    if (!io.score && !io.recon && !io.cnn_in) return;     // encode-only call (TemporalVAE.encode)
    }
    for (int item = tid; item < H * BM; item += NT) {
        const int k = item / BM, w = item - k * BM;
        float y = __ldg(P.l2h_b + k);
        const float* wr = P.l2h_w + k * P.Z;
        for (int zi = 0; zi < P.Z; ++zi) y = fmaf(zS[zi * BM + w], __ldg(wr + zi), y);
        uT[item] = tanhf(y);
    }
    __syncthreads();
    for (int i = tid; i < S::XT; i += NT) xT[i] = 0.f;       // scratch back to a clean x staging buffer
    __syncthreads();

    // ---------------- decoder + output Linear + squared error ----------------
    lstm_phase<H, BM, true>(P, src, io, sm, n0, nvalid, T, sse);

    // deterministic per-window reduction over channels
    float* sseS = xT;                          // [D][BM]
#pragma unroll
    for (int it = 0; it < S::MAXI; ++it) {
        const int item = tid + it * NT;
        if (item < BM * P.D) sseS[item] = sse[it];
    }
    __syncthreads();
    if (tid < nvalid && io.score) {
        float s = 0.f;
        for (int d = 0; d < P.D; ++d) s += sseS[d * BM + tid];
        io.score[n0 + tid] = s / (float)(T * P.D);
    }
}

}  // namespace shm
