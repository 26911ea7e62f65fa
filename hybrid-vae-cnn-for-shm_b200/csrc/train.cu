// LSTM-VAE training step for libshmfast (sm_100a): forward in train mode with saved activations, full BPTT,
// ELBO loss + upstream gradients, global-norm clip + Adam.  Replaces the inner loop of
// 4DOF/Scripts/03_train_vae.py:260-271 (forward -> mse + kl_w*KL -> backward -> clip_grad_norm_(2.0) -> Adam).
//
// Layout: every sequence tensor in the workspace is TIME-MAJOR [T][B][F], so one timestep of the whole batch is
// contiguous and every batched contraction over (t,b) has a single row stride.
//   * input projections / dX / dW are dense fp32 contractions over all timesteps (sgemm_kernel, split-K for dW);
//   * the recurrence is a persistent kernel per (layer, direction): the batch is split across CTAs (windows are
//     independent, so there is no inter-CTA exchange), W_hh is RESIDENT for all T steps -- 3/4 of it in shared
//     memory (192 KB at H=128) and the rest in registers -- and the 4 gates of a unit sit in 4 adjacent lanes so
//     the cell update is a warp-shuffle exchange;
//   * heads (LayerNorm, mu/logvar, reparameterisation, latent->hidden) are one small kernel per direction.
#include <cooperative_groups.h>
#include <new>
#include "common.cuh"
#include "gemm.cuh"
#include "tcgen05.cuh"

namespace shm {

// ------------------------------------------------------------------------------------------------------------
// parameter layout: list(model.parameters()) order of TemporalVAE (temporal_vae.py:29-49)
// ------------------------------------------------------------------------------------------------------------
struct ParamLayout {
    size_t enc_wih[SHM_MAX_L], enc_whh[SHM_MAX_L], enc_bih[SHM_MAX_L], enc_bhh[SHM_MAX_L];
    size_t ln_w, ln_b, mu_w, mu_b, lv_w, lv_b, l2h_w, l2h_b;
    size_t dec_wih[SHM_MAX_L], dec_whh[SHM_MAX_L], dec_bih[SHM_MAX_L], dec_bhh[SHM_MAX_L];
    size_t out_w, out_b, total;
};

static ParamLayout param_layout(const shm_vae_cfg& c) {
    ParamLayout p;
    memset(&p, 0, sizeof(p));
    size_t o = 0;
    const size_t H = c.H, D = c.D, Z = c.Z;
    for (int l = 0; l < c.L; ++l) {
        const size_t in = l == 0 ? D : H;
        p.enc_wih[l] = o; o += 4 * H * in;
        p.enc_whh[l] = o; o += 4 * H * H;
        p.enc_bih[l] = o; o += 4 * H;
        p.enc_bhh[l] = o; o += 4 * H;
    }
    if (c.has_ln) { p.ln_w = o; o += H; p.ln_b = o; o += H; }
    p.mu_w = o; o += Z * H; p.mu_b = o; o += Z;
    p.lv_w = o; o += Z * H; p.lv_b = o; o += Z;
    p.l2h_w = o; o += H * Z; p.l2h_b = o; o += H;
    for (int l = 0; l < c.L; ++l) {
        p.dec_wih[l] = o; o += 4 * H * H;
        p.dec_whh[l] = o; o += 4 * H * H;
        p.dec_bih[l] = o; o += 4 * H;
        p.dec_bhh[l] = o; o += 4 * H;
    }
    p.out_w = o; o += D * H; p.out_b = o; o += D;
    p.total = o;
    return p;
}

// ------------------------------------------------------------------------------------------------------------
// fp32 contraction  C[m][n] (+)= sum_k A(m,k) * B(k,n) (+ bias1[n] + bias2[n]),  A(m,k) = A[m*a_ms + k*a_ks],
// B(k,n) = B[k*b_ks + n*b_ns].  256 threads, register tile TM x TN per thread, split-K through blockIdx.z with an
// atomicAdd epilogue (the output must be zeroed by the caller in that case).
// ------------------------------------------------------------------------------------------------------------
struct GemmArgs {
    const float* A; long long a_ms, a_ks;
    const float* B; long long b_ks, b_ns;
    float* C; long long ldc;
    int M, N, K;
    const float* bias1; const float* bias2;
    int kchunk, atomic;
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256, 2) sgemm_kernel(const GemmArgs g) {
    constexpr int BK = 16;
    constexpr int NI = TM / 4, NJ = TN / 4;
    constexpr int TX = BN / TN;                  // thread columns
    static_assert((BM / TM) * TX == 256, "256 threads");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = min(g.K, kbeg + g.kchunk);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    // global -> registers -> shared, one tile ahead: the loads of tile k+1 are in flight while tile k is multiplied
    constexpr int NA = BM * BK / 256, NB = BN * BK / 256;
    float ra[NA], rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int e = 0; e < NA; ++e) {
            const int i = tid + e * 256;
            int m, k;
            if (g.a_ks == 1) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }     // fastest index along the contiguous dim
            const int gm = m0 + m, gk = k0 + k;
            ra[e] = (gm < g.M && gk < kend) ? __ldg(g.A + (long long)gm * g.a_ms + (long long)gk * g.a_ks) : 0.f;
        }
#pragma unroll
        for (int e = 0; e < NB; ++e) {
            const int i = tid + e * 256;
            int n, k;
            if (g.b_ks == 1) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
            const int gn = n0 + n, gk = k0 + k;
            rb[e] = (gn < g.N && gk < kend) ? __ldg(g.B + (long long)gk * g.b_ks + (long long)gn * g.b_ns) : 0.f;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int e = 0; e < NA; ++e) {
            const int i = tid + e * 256;
            int m, k;
            if (g.a_ks == 1) { k = i % BK; m = i / BK; } else { m = i % BM; k = i / BM; }
            As[k][m] = ra[e];
        }
#pragma unroll
        for (int e = 0; e < NB; ++e) {
            const int i = tid + e * 256;
            int n, k;
            if (g.b_ks == 1) { k = i % BK; n = i / BK; } else { n = i % BN; k = i / BN; }
            Bs[k][n] = rb[e];
        }
    };
    if (kbeg < kend) fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        commit();
        __syncthreads();
        if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(&As[k][i * (BM / NI) + ty * 4]);
                a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const float4 v = *reinterpret_cast<const float4*>(&Bs[k][j * (BN / NJ) + tx * 4]);
                b[4 * j] = v.x; b[4 * j + 1] = v.y; b[4 * j + 2] = v.z; b[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool add_bias = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / NI) + ty * 4 + (i & 3);
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + (j / 4) * (BN / NJ) + tx * 4 + (j & 3);
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (add_bias) {
                if (g.bias1) v += __ldg(g.bias1 + n);
                if (g.bias2) v += __ldg(g.bias2 + n);
            }
            float* c = g.C + (long long)m * g.ldc + n;
            if (g.atomic) atomicAdd(c, v); else *c = v;
        }
    }
}


// Forward contractions (activations x weights, O(1) magnitudes) take the fp16 split, backward ones (gradients) the bf16 split;
// small or oddly shaped ones stay on the fp32 FMA kernel below (sgemm_mode decides).
static int g_train_tc = 1;
static int g_rec_cluster = 0;       // H = 128 recurrence on clusters of two CTAs (shm_train_set_tensor_cores bit 1 clear: single-CTA kernels)          // shm_vae_trainer_set_engine: 0 = all contractions on the fp32 FMA pipe
static inline int gemm(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns,
                       float* C, long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk,
                       int mode = SHM_GEMM_SIMT) {
    return sgemm_mode(st, A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, splitk, g_train_tc ? mode : SHM_GEMM_SIMT);
}

int sgemm_simt(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
               long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk) {
    if (M <= 0 || N <= 0 || K <= 0) return SHM_OK;
    GemmArgs g{A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, K, splitk ? 1 : 0};
    const bool narrow = N <= 16;
    const bool small = !narrow && (((M + 127) / 128) * ((N + 127) / 128) < 74);      // fewer than half a wave of 128x128 tiles
    const int BM = narrow ? 256 : (small ? 64 : 128), BN = narrow ? 16 : (small ? 64 : 128);
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    int splits = 1;
    if (splitk) {
        splits = max(1, min((K + 63) / 64, 296 / max(1, tiles)));
        g.kchunk = ((K + splits - 1) / splits + 15) / 16 * 16;
        splits = (K + g.kchunk - 1) / g.kchunk;
    }
    dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, splits);
    if (narrow) sgemm_kernel<256, 16, 4, 4><<<grid, 256, 0, st>>>(g);
    else if (small) sgemm_kernel<64, 64, 4, 4><<<grid, 256, 0, st>>>(g);
    else sgemm_kernel<128, 128, 8, 8><<<grid, 256, 0, st>>>(g);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

// out1[n] (and out2[n]) += sum_m A[m][n] * (mul ? mul[m][n] : 1); outputs zeroed by the caller.
__global__ void colsum_kernel(const float* __restrict__ A, const float* __restrict__ mul, int M, int N, int rows_per_block,
                              float* __restrict__ out1, float* __restrict__ out2) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n = blockIdx.x * 32 + tx;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
    float s = 0.f;
    if (n < N)
        for (int r = r0 + ty; r < r1; r += 8) {
            const float v = A[(size_t)r * N + n];
            s += mul ? v * mul[(size_t)r * N + n] : v;
        }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        atomicAdd(out1 + n, t);
        if (out2) atomicAdd(out2 + n, t);
    }
}

static int colsum(cudaStream_t st, const float* A, const float* mul, int M, int N, float* out1, float* out2) {
    const int rpb = max(64, (M + 63) / 64);
    dim3 grid((N + 31) / 32, (M + rpb - 1) / rpb);
    colsum_kernel<<<grid, dim3(32, 8), 0, st>>>(A, mul, M, N, rpb, out1, out2);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

// [B][T][F] <-> [T][B][F]
__global__ void swap01_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int Bd, int F) {
    const size_t n = (size_t)A * Bd * F;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const size_t r = i / F;
        const int b = (int)(r % Bd), a = (int)(r / Bd);           // out index (b-major? no): out[b][a][f]
        out[((size_t)b * A + a) * F + f] = in[i];
    }
}

// ------------------------------------------------------------------------------------------------------------
// Recurrence, forward.  grid = ceil(B / NW), block = 2H threads.  Thread j <-> (unit u = j>>1, pair p = j&1) owns TWO
// gate rows of unit u -- p=0: (i, f), p=1: (g, o) -- so every broadcast load of h_{t-1} feeds two dot products and the
// cell update needs one lane exchange (shfl.xor 1).  W_hh stays resident for all T steps: k < KS in shared memory
// ([k/4][row] float4, conflict-free), k >= KS in registers (H=128: 128 KB smem + 128 registers per thread; H<=64: all
// of it in registers).
// ------------------------------------------------------------------------------------------------------------
template <int H> struct RecCfg {
    static constexpr int KS = (H == 128) ? 64 : 0;     // forward: k range of W_hh kept in shared memory
    static constexpr int KR = H - KS;                  // forward: k range kept in registers (x2 rows per thread)
    static constexpr int MS = (H == 128) ? 128 : 0;    // backward: rows (of the 2H a thread reduces over) kept in shared memory
    static constexpr int MR = 2 * H - MS;              // backward: rows kept in registers
};

// ex2.approx + rcp.approx (2 ulp each; the accurate expf / IEEE division / tanhf each carry a slow-path call, visible as
// CALL.REL + BSSY in the step loop's SASS), far inside the 1e-4 gradient tolerance
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_acc(float x) { return fmaf(2.f, sigmoidf_acc(2.f * x), -1.f); }

template <int H, int NW>
__global__ void __launch_bounds__(2 * H)
lstm_rec_fwd_kernel(const float* __restrict__ w_hh, const float* gx, long long gx_tstride, float* act,
                    float* __restrict__ hs, float* __restrict__ cs, float* __restrict__ hd, const uint8_t* __restrict__ mask,
                    float scale, int T, int B) {
    constexpr int G4 = 4 * H, KS = RecCfg<H>::KS, KR = RecCfg<H>::KR;
    extern __shared__ float4 smem4[];
    float4* Wt = smem4;                                         // [KS/4][4H]: slot j = row r0 of thread j, slot 2H+j = row r1
    float* hbuf = reinterpret_cast<float*>(Wt + (KS / 4) * G4); // [2][NW][H]
    const int j = threadIdx.x, p = j & 1, u = j >> 1;
    const int r0 = (2 * p) * H + u, r1 = (2 * p + 1) * H + u;   // reference row order: gate*H + unit, gates i,f,g,o
    const float* w0 = w_hh + (size_t)r0 * H;
    const float* w1 = w_hh + (size_t)r1 * H;
    for (int k4 = 0; k4 < KS / 4; ++k4) {      // scalar loads: flat parameter offsets are not 16-byte aligned in general
        Wt[k4 * G4 + j] = make_float4(w0[4 * k4], w0[4 * k4 + 1], w0[4 * k4 + 2], w0[4 * k4 + 3]);
        Wt[k4 * G4 + 2 * H + j] = make_float4(w1[4 * k4], w1[4 * k4 + 1], w1[4 * k4 + 2], w1[4 * k4 + 3]);
    }
    float wr0[KR], wr1[KR];
#pragma unroll
    for (int i = 0; i < KR; ++i) { wr0[i] = w0[KS + i]; wr1[i] = w1[KS + i]; }
    for (int i = j; i < 2 * NW * H; i += 2 * H) hbuf[i] = 0.f;
    const int b0 = blockIdx.x * NW;
    float c[NW], pre0[NW], pre1[NW];
    unsigned int mk[NW];                        // raw keep-mask byte of the NEXT step (converted only when consumed)
    bool valid[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        c[w] = 0.f;
        valid[w] = (b0 + w) < B;
        pre0[w] = valid[w] ? gx[(size_t)(b0 + w) * G4 + r0] : 0.f;
        pre1[w] = valid[w] ? gx[(size_t)(b0 + w) * G4 + r1] : 0.f;
        mk[w] = (mask && valid[w]) ? mask[((size_t)(b0 + w) * T) * H + u] : 1u;
        if (valid[w] && p == 0) { hs[(size_t)(b0 + w) * H + u] = 0.f; cs[(size_t)(b0 + w) * H + u] = 0.f; }
    }
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        float a0[NW], a1[NW];
        unsigned int mcur[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) { a0[w] = pre0[w]; a1[w] = pre1[w]; mcur[w] = mk[w]; }
        if (t + 1 < T) {                        // next step's inputs: issued a full step before they are consumed
#pragma unroll
            for (int w = 0; w < NW; ++w)
                if (valid[w]) {
                    const size_t o = (size_t)(t + 1) * gx_tstride + (size_t)(b0 + w) * G4;
                    pre0[w] = gx[o + r0];
                    pre1[w] = gx[o + r1];
                    if (mask) mk[w] = mask[((size_t)(b0 + w) * T + t + 1) * H + u];
                }
        }
        const float* hb = hbuf + (t & 1) * NW * H;
#pragma unroll
        for (int k4 = 0; k4 < KS / 4; ++k4) {
            const float4 x0 = Wt[k4 * G4 + j];
            const float4 x1 = Wt[k4 * G4 + 2 * H + j];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb + w * H + 4 * k4);
                a0[w] = fmaf(x0.x, h4.x, a0[w]); a0[w] = fmaf(x0.y, h4.y, a0[w]);
                a0[w] = fmaf(x0.z, h4.z, a0[w]); a0[w] = fmaf(x0.w, h4.w, a0[w]);
                a1[w] = fmaf(x1.x, h4.x, a1[w]); a1[w] = fmaf(x1.y, h4.y, a1[w]);
                a1[w] = fmaf(x1.z, h4.z, a1[w]); a1[w] = fmaf(x1.w, h4.w, a1[w]);
            }
        }
#pragma unroll
        for (int i = 0; i < KR; i += 4) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb + w * H + KS + i);
                a0[w] = fmaf(wr0[i], h4.x, a0[w]); a0[w] = fmaf(wr0[i + 1], h4.y, a0[w]);
                a0[w] = fmaf(wr0[i + 2], h4.z, a0[w]); a0[w] = fmaf(wr0[i + 3], h4.w, a0[w]);
                a1[w] = fmaf(wr1[i], h4.x, a1[w]); a1[w] = fmaf(wr1[i + 1], h4.y, a1[w]);
                a1[w] = fmaf(wr1[i + 2], h4.z, a1[w]); a1[w] = fmaf(wr1[i + 3], h4.w, a1[w]);
            }
        }
        float* hn = hbuf + ((t + 1) & 1) * NW * H;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float v0 = (p == 1) ? tanhf_acc(a0[w]) : sigmoidf_acc(a0[w]);   // p=0: i, p=1: g
            const float v1 = sigmoidf_acc(a1[w]);                              // p=0: f, p=1: o
            const float o0 = __shfl_xor_sync(0xffffffffu, v0, 1);
            const float o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
            const float gi = p ? o0 : v0, gf = p ? o1 : v1, gg = p ? v0 : o0, go = p ? v1 : o1;
            c[w] = fmaf(gf, c[w], gi * gg);
            const float h = go * tanhf_acc(c[w]);
            if (p == 0) hn[w * H + u] = h;
            if (valid[w]) {
                const size_t tb = (size_t)t * B + (b0 + w);
                act[tb * G4 + r0] = v0;
                act[tb * G4 + r1] = v1;
                const size_t o = ((size_t)(t + 1) * B + (b0 + w)) * H + u;
                if (p == 0) { hs[o] = h; cs[o] = c[w]; }
                else if (hd) hd[tb * H + u] = mask ? (mcur[w] ? h * scale : 0.f) : h;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------
// Recurrence, backward (BPTT).  Phase A (thread <-> (unit, pair) as in the forward kernel): pre-activation gradients of
// the thread's two gate rows at step t; phase B (thread <-> (k, half q) with q = tid / H): the partial
// dh_{t-1}[k] = sum over the 2H gate rows of half q of dG[n] * W_hh[n][k], W_hh^T resident (shared memory + registers).
// act_dg holds the saved gate activations on entry and the pre-activation gradients dG on exit (in place).
// ------------------------------------------------------------------------------------------------------------
template <int H, int NW>
__global__ void __launch_bounds__(2 * H)
lstm_rec_bwd_kernel(const float* __restrict__ w_hh, float* act_dg, const float* __restrict__ cs,
                    const float* __restrict__ dH, const uint8_t* __restrict__ mask, float scale,
                    const float* __restrict__ dh_last, float* __restrict__ dgsum, int T, int B) {
    constexpr int G4 = 4 * H, H2 = 2 * H, MS = RecCfg<H>::MS, MR = RecCfg<H>::MR;
    extern __shared__ float4 smem4[];
    float4* Wb = smem4;                                           // [MS/4][2H]
    float* dgs = reinterpret_cast<float*>(Wb + (MS / 4) * H2);    // [NW][4H]   (reference row order n = gate*H + unit)
    float* red = dgs + NW * G4;                                   // [2][NW][H]
    const int tid = threadIdx.x;
    const int p = tid & 1, u = tid >> 1;                          // phase A role
    const int r0 = (2 * p) * H + u, r1 = (2 * p + 1) * H + u;
    const int q = tid / H, k = tid % H;                           // phase B role: rows [q*2H, (q+1)*2H), column k
    for (int m4 = 0; m4 < MS / 4; ++m4) {
        const float* src = w_hh + (size_t)(q * H2 + 4 * m4) * H + k;
        Wb[m4 * H2 + tid] = make_float4(src[0], src[H], src[2 * H], src[3 * H]);
    }
    float wreg[MR];
#pragma unroll
    for (int i = 0; i < MR; ++i) wreg[i] = w_hh[(size_t)(q * H2 + MS + i) * H + k];
    for (int i = tid; i < 2 * NW * H; i += H2) red[i] = 0.f;
    const int b0 = blockIdx.x * NW;
    float dc[NW], gs0[NW], gs1[NW], a0_n[NW], a1_n[NW], ct_n[NW], cp_n[NW], dh_n[NW], dl_n[NW];
    unsigned int mk_n[NW];
    bool valid[NW];
    auto fetch = [&](int t) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (!valid[w]) { a0_n[w] = 0.f; a1_n[w] = 0.f; ct_n[w] = 0.f; cp_n[w] = 0.f; dh_n[w] = 0.f; dl_n[w] = 0.f; mk_n[w] = 1u; continue; }
            const int b = b0 + w;
            const size_t tb = (size_t)t * B + b;
            a0_n[w] = act_dg[tb * G4 + r0];
            a1_n[w] = act_dg[tb * G4 + r1];
            ct_n[w] = cs[((size_t)(t + 1) * B + b) * H + u];
            cp_n[w] = cs[tb * H + u];
            // raw loads only: nothing here may consume a loaded value, or the prefetch turns into a stall
            dh_n[w] = dH ? dH[tb * H + u] : 0.f;
            mk_n[w] = mask ? mask[((size_t)b * T + t) * H + u] : 1u;
            dl_n[w] = (dh_last && t == T - 1) ? dh_last[(size_t)b * H + u] : 0.f;
        }
    };
#pragma unroll
    for (int w = 0; w < NW; ++w) { dc[w] = 0.f; gs0[w] = 0.f; gs1[w] = 0.f; valid[w] = (b0 + w) < B; }
    fetch(T - 1);
    __syncthreads();
    for (int t = T - 1; t >= 0; --t) {
        float v0c[NW], v1c[NW], ct[NW], cp[NW], dho[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            v0c[w] = a0_n[w]; v1c[w] = a1_n[w]; ct[w] = ct_n[w]; cp[w] = cp_n[w];
            dho[w] = (mask ? (mk_n[w] ? dh_n[w] * scale : 0.f) : dh_n[w]) + dl_n[w];
        }
        if (t > 0) fetch(t - 1);
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float dh = dho[w] + (red[w * H + u] + red[(NW + w) * H + u]);
            const float o0 = __shfl_xor_sync(0xffffffffu, v0c[w], 1);
            const float o1 = __shfl_xor_sync(0xffffffffu, v1c[w], 1);
            const float gi = p ? o0 : v0c[w], gf = p ? o1 : v1c[w], gg = p ? v0c[w] : o0, go = p ? v1c[w] : o1;
            const float tc = tanhf_acc(ct[w]);
            const float dct = fmaf(dh * go, 1.f - tc * tc, dc[w]);
            dc[w] = dct * gf;
            float g0, g1;
            if (p == 0) { g0 = dct * gg * gi * (1.f - gi); g1 = dct * cp[w] * gf * (1.f - gf); }      // d pre-act of i, f
            else { g0 = dct * gi * (1.f - gg * gg); g1 = dh * tc * go * (1.f - go); }                // d pre-act of g, o
            if (!valid[w]) { g0 = 0.f; g1 = 0.f; }
            dgs[w * G4 + r0] = g0;
            dgs[w * G4 + r1] = g1;
            gs0[w] += g0; gs1[w] += g1;
            if (valid[w]) {
                const size_t tb = (size_t)t * B + b0 + w;
                act_dg[tb * G4 + r0] = g0;
                act_dg[tb * G4 + r1] = g1;
            }
        }
        __syncthreads();
        float acc[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) acc[w] = 0.f;
#pragma unroll
        for (int m4 = 0; m4 < MS / 4; ++m4) {
            const float4 w4 = Wb[m4 * H2 + tid];
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float4 d4 = *reinterpret_cast<const float4*>(dgs + w * G4 + q * H2 + 4 * m4);
                acc[w] = fmaf(w4.x, d4.x, acc[w]); acc[w] = fmaf(w4.y, d4.y, acc[w]);
                acc[w] = fmaf(w4.z, d4.z, acc[w]); acc[w] = fmaf(w4.w, d4.w, acc[w]);
            }
        }
#pragma unroll
        for (int i = 0; i < MR; i += 4) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const float4 d4 = *reinterpret_cast<const float4*>(dgs + w * G4 + q * H2 + MS + i);
                acc[w] = fmaf(wreg[i], d4.x, acc[w]); acc[w] = fmaf(wreg[i + 1], d4.y, acc[w]);
                acc[w] = fmaf(wreg[i + 2], d4.z, acc[w]); acc[w] = fmaf(wreg[i + 3], d4.w, acc[w]);
            }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) red[(q * NW + w) * H + k] = acc[w];
        __syncthreads();
    }
    if (dgsum) {
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (valid[w]) { dgsum[(size_t)(b0 + w) * G4 + r0] = gs0[w]; dgsum[(size_t)(b0 + w) * G4 + r1] = gs1[w]; }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Recurrence for H = 128 on a CLUSTER OF TWO CTAs (r02).  profiles/r01_train_rec_v3_raw.csv: the single-CTA kernels above are
// bound by shared-memory wavefronts -- half of W_hh (128 KB) is re-read from shared memory every step next to the h_{t-1}
// broadcasts (2.4 k wavefronts per step, LSU data pipe 53 % busy, short-scoreboard stalls dominate, 5.2 k clk per step).  Two SMs
// together hold all of W_hh in REGISTERS: CTA c owns the four gate rows of units [64c, 64c+64) (256 rows x 128 k = 128 registers per
// thread), so the only shared-memory traffic left is the h_{t-1} broadcast; each CTA writes its half of h_t into both CTAs'
// buffers (distributed shared memory) and one cluster barrier per step replaces the block barrier.  Thread j <-> (unit 64c + (j>>2),
// gate j&3): the 4 gates of a unit sit in 4 adjacent lanes (warp-shuffle cell update).
// ------------------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

// distributed-shared-memory store that signals the RECEIVER's mbarrier (complete_tx): data and "it has landed" travel together,
// so a step needs no cluster-wide barrier -- cluster.sync() compiles to MEMBAR.ALL.GPU + CCTL.IVALL, which waits for every
// outstanding global store of the step (measured: 3.5 us per step with it)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t peer_addr, float v, uint32_t peer_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(peer_addr), "r"(__float_as_uint(v)), "r"(peer_mbar) : "memory");
}
// fast gate activations: ex2.approx + rcp.approx (2 ulp each), far inside the 1e-4 gradient tolerance
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return fmaf(2.f, sigmoid_fast(2.f * x), -1.f); }

// sigma(x) for the i, f, o gates and tanh(x) = 2 sigma(2x) - 1 for the g gate through one branch-free formula
__device__ __forceinline__ float gate_act(float x, int g) {
    const float s = g == 2 ? 2.f : 1.f;
    const float sg = 1.f / (1.f + expf(-s * x));
    return g == 2 ? fmaf(2.f, sg, -1.f) : sg;
}

template <int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
lstm_rec_fwd_c2_kernel(const float* __restrict__ w_hh, const float* gx, long long gx_tstride, float* act,
                       float* __restrict__ hs, float* __restrict__ cs, float* __restrict__ hd, const uint8_t* __restrict__ mask,
                       float scale, int T, int B) {
    constexpr int H = 128, G4 = 512;
    __shared__ __align__(16) float hbuf[2][NW][H];
    __shared__ __align__(8) uint64_t full[2];                   // "the peer's half of h in buffer b has landed"
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = cluster.block_rank();
    const uint32_t hbuf_peer = mapa_u32(tc::smem_u32(&hbuf[0][0][0]), crank ^ 1u);
    const uint32_t full_peer = mapa_u32(tc::smem_u32(&full[0]), crank ^ 1u);
    const int j = threadIdx.x, g = j & 3, u = (int)crank * 64 + (j >> 2);
    if (j == 0) { tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::fence_mbar_init(); }
    const int r = g * H + u;                                    // the row whose activation this thread ends up with (gate g = k-slice id)
    // A broadcast LDS.128 still costs 4 wavefronts (one per quarter warp), so h_{t-1} is NOT broadcast to one-row threads: the 4
    // lanes of a unit split K -- lane g holds columns {16m + 4g .. 16m + 4g + 3, m < 8} of all FOUR gate rows of its unit (128
    // registers; interleaved so the 4 lanes read 64 contiguous bytes: no bank conflict), reads only
    // its 32 values of h (8 LDS.128 per window instead of 32, 4 distinct addresses per quarter warp) and a 3-shuffle
    // reduce-scatter leaves the full sum of gate g in lane g (measured before: 0.7 us per window-step, LSU bound).
    float wr[4][32];
#pragma unroll
    for (int gg = 0; gg < 4; ++gg)
#pragma unroll
        for (int i = 0; i < 32; ++i) wr[gg][i] = w_hh[(size_t)(gg * H + u) * H + 16 * (i >> 2) + 4 * g + (i & 3)];
    for (int i = j; i < 2 * NW * H; i += 256) (&hbuf[0][0][0])[i] = 0.f;
    const int b0 = (blockIdx.x >> 1) * NW;
    const int lane_base = (j & 31) & ~3;
    float c[NW], pre[NW];
    unsigned int mk[NW];
    bool valid[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        c[w] = 0.f;
        valid[w] = (b0 + w) < B;
        pre[w] = valid[w] ? gx[(size_t)(b0 + w) * G4 + r] : 0.f;
        mk[w] = (mask && valid[w] && g == 1) ? mask[((size_t)(b0 + w) * T) * H + u] : 1u;
        if (valid[w] && g == 0) { hs[(size_t)(b0 + w) * H + u] = 0.f; cs[(size_t)(b0 + w) * H + u] = 0.f; }
    }
    cluster.sync();                              // buffers zeroed and barriers initialised before a peer writes into them
    for (int t = 0; t < T; ++t) {
        float a[NW];
        unsigned int mcur[NW];
        const int nbuf = (t + 1) & 1;
        if (j == 0) tc::mbar_arrive_expect_tx(&full[nbuf], NW * 64 * 4);      // the peer's 64 units x NW windows of h_{t+1}
#pragma unroll
        for (int w = 0; w < NW; ++w) { a[w] = pre[w]; mcur[w] = mk[w]; }
        if (t + 1 < T) {                        // next step's inputs: issued a full step before they are consumed
#pragma unroll
            for (int w = 0; w < NW; ++w)
                if (valid[w]) {
                    pre[w] = gx[(size_t)(t + 1) * gx_tstride + (size_t)(b0 + w) * G4 + r];
                    if (mask && g == 1) mk[w] = mask[((size_t)(b0 + w) * T + t + 1) * H + u];
                }
        }
        const float* hb = &hbuf[t & 1][0][0] + 4 * g;
        const int nb = nbuf * NW * H;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;         // this lane's K-slice of the four gate rows
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 h4 = *reinterpret_cast<const float4*>(hb + w * H + 4 * i);
                p0 = fmaf(wr[0][i], h4.x, p0); p1 = fmaf(wr[1][i], h4.x, p1); p2 = fmaf(wr[2][i], h4.x, p2); p3 = fmaf(wr[3][i], h4.x, p3);
                p0 = fmaf(wr[0][i + 1], h4.y, p0); p1 = fmaf(wr[1][i + 1], h4.y, p1); p2 = fmaf(wr[2][i + 1], h4.y, p2); p3 = fmaf(wr[3][i + 1], h4.y, p3);
                p0 = fmaf(wr[0][i + 2], h4.z, p0); p1 = fmaf(wr[1][i + 2], h4.z, p1); p2 = fmaf(wr[2][i + 2], h4.z, p2); p3 = fmaf(wr[3][i + 2], h4.z, p3);
                p0 = fmaf(wr[0][i + 3], h4.w, p0); p1 = fmaf(wr[1][i + 3], h4.w, p1); p2 = fmaf(wr[2][i + 3], h4.w, p2); p3 = fmaf(wr[3][i + 3], h4.w, p3);
            }
            // reduce-scatter over the 4 lanes of the unit: lane g ends with the full sum of gate g
            const bool hi2 = (g & 2) != 0, hi1 = (g & 1) != 0;
            const float r0 = __shfl_xor_sync(0xffffffffu, hi2 ? p0 : p2, 2);
            const float r1 = __shfl_xor_sync(0xffffffffu, hi2 ? p1 : p3, 2);
            const float k0 = (hi2 ? p2 : p0) + r0, k1 = (hi2 ? p3 : p1) + r1;
            const float r2 = __shfl_xor_sync(0xffffffffu, hi1 ? k0 : k1, 1);
            const float x = a[w] + ((hi1 ? k1 : k0) + r2);
            const float sg = sigmoid_fast(g == 2 ? 2.f * x : x);
            const float v = g == 2 ? fmaf(2.f, sg, -1.f) : sg;
            const float gi = __shfl_sync(0xffffffffu, v, lane_base);
            const float gf = __shfl_sync(0xffffffffu, v, lane_base + 1);
            const float gg = __shfl_sync(0xffffffffu, v, lane_base + 2);
            const float go = __shfl_sync(0xffffffffu, v, lane_base + 3);
            c[w] = fmaf(gf, c[w], gi * gg);
            const float h = go * tanh_fast(c[w]);
            if (g == 0) (&hbuf[0][0][0])[nb + w * H + u] = h;
            else if (g == 3) st_async_f32(hbuf_peer + (uint32_t)(nb + w * H + u) * 4u, h, full_peer + (uint32_t)nbuf * 8u);
            if (valid[w]) {
                const size_t tb = (size_t)t * B + (b0 + w);
                act[tb * G4 + r] = v;
                const size_t o = ((size_t)(t + 1) * B + (b0 + w)) * H + u;
                if (g == 0) hs[o] = h;
                else if (g == 2) cs[o] = c[w];
                else if (g == 1 && hd) hd[tb * H + u] = mask ? (mcur[w] ? h * scale : 0.f) : h;
            }
        }
        __syncthreads();                         // this CTA's half of h_{t+1} is in place, everyone is done reading h_t
        // the peer's half: buffer 1 is used at t+1 = 1, 3, 5 ..., buffer 0 at t+1 = 2, 4, ...
        tc::mbar_wait(&full[nbuf], (uint32_t)((((t + 1) >> 1) - (nbuf ? 0 : 1)) & 1));
    }
    cluster.sync();                              // no CTA leaves while its peer may still write into its shared memory
}

// Backward on the same cluster.  Phase A (thread <-> (unit 64c + (tid>>2), gate tid&3)): the pre-activation gradient of the thread's
// gate row, written into BOTH CTAs' dG buffers; cluster barrier; phase B (thread <-> (k = 64c + (tid&63), quarter q = tid>>6)): the
// partial dh_{t-1}[k] over the 128 rows of gate q with W_hh[.,k] in registers; block barrier; the next phase A sums the 4 partials.
template <int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
lstm_rec_bwd_c2_kernel(const float* __restrict__ w_hh, float* act_dg, const float* __restrict__ cs,
                       const float* __restrict__ dH, const uint8_t* __restrict__ mask, float scale,
                       const float* __restrict__ dh_last, float* __restrict__ dgsum, int T, int B) {
    constexpr int H = 128, G4 = 512, HL = 64;
    __shared__ __align__(16) float dgs[2][NW][G4];                // double buffered: a fast CTA may already write step t-1
    __shared__ float dhs[NW][HL];                                 // dh_{t-1}[k] of this CTA's 64 units (phase B -> next phase A)
    __shared__ __align__(8) uint64_t full[2];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int crank = cluster.block_rank();
    const uint32_t dgs_peer = mapa_u32(tc::smem_u32(&dgs[0][0][0]), crank ^ 1u);
    const uint32_t full_peer = mapa_u32(tc::smem_u32(&full[0]), crank ^ 1u);
    const int tid = threadIdx.x;
    if (tid == 0) { tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::fence_mbar_init(); }
    const int g = tid & 3, ul = tid >> 2, u = (int)crank * HL + ul;      // phase A role
    const int r = g * H + u;
    // phase B role: thread <-> (4 columns k0..k0+3, rows n = 64m + 4*ns + j, m < 8, j < 4): 128 registers of W_hh^T; the 16 lanes
    // of a column group read 256 contiguous bytes of dG (8 LDS.128 per window instead of 32 broadcast ones) and a 5-shuffle
    // reduction leaves dh[k0 + (ns >> 2)] in every lane of the group
    const int kg = tid >> 4, ns = tid & 15, k0 = (int)crank * HL + kg * 4;
    float wreg[32][4];
#pragma unroll
    for (int i = 0; i < 32; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) wreg[i][jj] = w_hh[(size_t)(64 * (i >> 2) + 4 * ns + (i & 3)) * H + k0 + jj];
    for (int i = tid; i < NW * HL; i += 256) (&dhs[0][0])[i] = 0.f;
    const int b0 = (blockIdx.x >> 1) * NW;
    const int lane_base = (tid & 31) & ~3;
    float dc[NW], gs[NW], a_n[NW], ct_n[NW], cp_n[NW], dh_n[NW], dl_n[NW];
    unsigned int mk_n[NW];
    bool valid[NW];
    auto fetch = [&](int t) {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            if (!valid[w]) { a_n[w] = 0.f; ct_n[w] = 0.f; cp_n[w] = 0.f; dh_n[w] = 0.f; dl_n[w] = 0.f; mk_n[w] = 1u; continue; }
            const int b = b0 + w;
            const size_t tb = (size_t)t * B + b;
            a_n[w] = act_dg[tb * G4 + r];
            ct_n[w] = cs[((size_t)(t + 1) * B + b) * H + u];
            cp_n[w] = cs[tb * H + u];
            dh_n[w] = dH ? dH[tb * H + u] : 0.f;
            mk_n[w] = mask ? mask[((size_t)b * T + t) * H + u] : 1u;
            dl_n[w] = (dh_last && t == T - 1) ? dh_last[(size_t)b * H + u] : 0.f;
        }
    };
#pragma unroll
    for (int w = 0; w < NW; ++w) { dc[w] = 0.f; gs[w] = 0.f; valid[w] = (b0 + w) < B; }
    fetch(T - 1);
    cluster.sync();
    for (int t = T - 1; t >= 0; --t) {
        float vc[NW], ct[NW], cp[NW], dho[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            vc[w] = a_n[w]; ct[w] = ct_n[w]; cp[w] = cp_n[w];
            dho[w] = (mask ? (mk_n[w] ? dh_n[w] * scale : 0.f) : dh_n[w]) + dl_n[w];
        }
        if (t > 0) fetch(t - 1);
        const int dbuf = t & 1;
        const int db = dbuf * NW * G4;
        if (tid == 0) tc::mbar_arrive_expect_tx(&full[dbuf], NW * 256 * 4);      // the peer's 256 rows x NW windows of dG_t
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float dh = dho[w] + dhs[w][ul];
            const float gi = __shfl_sync(0xffffffffu, vc[w], lane_base);
            const float gf = __shfl_sync(0xffffffffu, vc[w], lane_base + 1);
            const float gg = __shfl_sync(0xffffffffu, vc[w], lane_base + 2);
            const float go = __shfl_sync(0xffffffffu, vc[w], lane_base + 3);
            const float tch = tanh_fast(ct[w]);
            const float dct = fmaf(dh * go, 1.f - tch * tch, dc[w]);
            dc[w] = dct * gf;
            float gv;
            if (g == 0) gv = dct * gg * gi * (1.f - gi);            // d pre-act of i
            else if (g == 1) gv = dct * cp[w] * gf * (1.f - gf);    // f
            else if (g == 2) gv = dct * gi * (1.f - gg * gg);       // g
            else gv = dh * tch * go * (1.f - go);                   // o
            if (!valid[w]) gv = 0.f;
            (&dgs[0][0][0])[db + w * G4 + r] = gv;
            st_async_f32(dgs_peer + (uint32_t)(db + w * G4 + r) * 4u, gv, full_peer + (uint32_t)dbuf * 8u);
            gs[w] += gv;
            if (valid[w]) act_dg[((size_t)t * B + b0 + w) * G4 + r] = gv;
        }
        __syncthreads();                         // this CTA's rows of dG_t are in place
        // the peer's rows: steps t = T-1, T-2, ... alternate the buffers; use n of buffer b is (T-1-t) >> 1
        tc::mbar_wait(&full[dbuf], (uint32_t)(((T - 1 - t) >> 1) & 1));
        const float* dg = &dgs[0][0][0] + db + 4 * ns;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;         // partial dh of the 4 columns over this lane's 32 rows
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float4 d4 = *reinterpret_cast<const float4*>(dg + w * G4 + 64 * m);
                p0 = fmaf(wreg[4 * m][0], d4.x, p0); p1 = fmaf(wreg[4 * m][1], d4.x, p1); p2 = fmaf(wreg[4 * m][2], d4.x, p2); p3 = fmaf(wreg[4 * m][3], d4.x, p3);
                p0 = fmaf(wreg[4 * m + 1][0], d4.y, p0); p1 = fmaf(wreg[4 * m + 1][1], d4.y, p1); p2 = fmaf(wreg[4 * m + 1][2], d4.y, p2); p3 = fmaf(wreg[4 * m + 1][3], d4.y, p3);
                p0 = fmaf(wreg[4 * m + 2][0], d4.z, p0); p1 = fmaf(wreg[4 * m + 2][1], d4.z, p1); p2 = fmaf(wreg[4 * m + 2][2], d4.z, p2); p3 = fmaf(wreg[4 * m + 2][3], d4.z, p3);
                p0 = fmaf(wreg[4 * m + 3][0], d4.w, p0); p1 = fmaf(wreg[4 * m + 3][1], d4.w, p1); p2 = fmaf(wreg[4 * m + 3][2], d4.w, p2); p3 = fmaf(wreg[4 * m + 3][3], d4.w, p3);
            }
            // reduce-scatter over lane bits 3, 2 (column (ns >> 2) stays), then all-reduce over bits 1, 0
            const bool hi8 = (ns & 8) != 0, hi4 = (ns & 4) != 0;
            const float r0 = __shfl_xor_sync(0xffffffffu, hi8 ? p0 : p2, 8);
            const float r1 = __shfl_xor_sync(0xffffffffu, hi8 ? p1 : p3, 8);
            const float k0s = (hi8 ? p2 : p0) + r0, k1s = (hi8 ? p3 : p1) + r1;
            const float r2 = __shfl_xor_sync(0xffffffffu, hi4 ? k0s : k1s, 4);
            float v = (hi4 ? k1s : k0s) + r2;
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if ((ns & 3) == 0) dhs[w][kg * 4 + (ns >> 2)] = v;
        }
        __syncthreads();
    }
    if (dgsum) {
#pragma unroll
        for (int w = 0; w < NW; ++w)
            if (valid[w]) dgsum[(size_t)(b0 + w) * G4 + r] = gs[w];
    }
    cluster.sync();                              // no CTA leaves while its peer may still write into its shared memory
}

// ------------------------------------------------------------------------------------------------------------
// Heads.  grid = B, block = H threads (thread <-> hidden unit).
// ------------------------------------------------------------------------------------------------------------
template <int H>
__device__ __forceinline__ float block_sum(float v, float* red) {
    constexpr int NWARP = H / 32;
    v = warp_sum(v);
    if constexpr (NWARP == 1) return v;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NWARP; ++i) s += red[i];
    return s;
}

struct HeadW {
    const float *ln_w, *ln_b, *mu_w, *mu_b, *lv_w, *lv_b, *l2h_w, *l2h_b;
    int has_ln, Z;
    float ln_eps;
};

// hn [B][H] -> LayerNorm -> mu, logvar -> z = mu + eps*exp(0.5 logvar) -> h0 = tanh(W z + b)
// (temporal_vae.py:55-58,60-63,66).  Saves xn, rstd, y, mu, logvar, z, h0 for the backward pass.
template <int H>
__global__ void __launch_bounds__(H) head_fwd_kernel(const float* __restrict__ hn, const float* __restrict__ eps, HeadW w,
                                                     float* __restrict__ xn, float* __restrict__ rstd, float* __restrict__ y,
                                                     float* __restrict__ mu, float* __restrict__ lv, float* __restrict__ z,
                                                     float* __restrict__ h0, float* __restrict__ mu_out, float* __restrict__ lv_out) {
    __shared__ float red[4], ys[H], zs[16], ms[16], ls[16];
    const int b = blockIdx.x, u = threadIdx.x, Z = w.Z;
    const float v = hn[(size_t)b * H + u];
    float yv = v;
    if (w.has_ln) {
        const float mean = block_sum<H>(v, red) * (1.f / H);
        const float d = v - mean;
        const float var = block_sum<H>(d * d, red) * (1.f / H);
        const float rs = 1.f / sqrtf(var + w.ln_eps);
        const float x = d * rs;
        xn[(size_t)b * H + u] = x;
        if (u == 0) rstd[b] = rs;
        yv = fmaf(x, w.ln_w[u], w.ln_b[u]);
    }
    y[(size_t)b * H + u] = yv;
    ys[u] = yv;
    __syncthreads();
    const int warp = u >> 5, lane = u & 31;
    for (int o = warp; o < 2 * Z; o += H / 32) {
        const float* wr = (o < Z) ? (w.mu_w + (size_t)o * H) : (w.lv_w + (size_t)(o - Z) * H);
        float s = 0.f;
        for (int kk = lane; kk < H; kk += 32) s = fmaf(wr[kk], ys[kk], s);
        s = warp_sum(s);
        if (lane == 0) {
            if (o < Z) ms[o] = s + w.mu_b[o]; else ls[o - Z] = s + w.lv_b[o - Z];
        }
    }
    __syncthreads();
    if (u < Z) {
        const float m = ms[u], l = ls[u];
        const float zz = fmaf(eps[(size_t)b * Z + u], expf(0.5f * l), m);
        zs[u] = zz;
        mu[(size_t)b * Z + u] = m; lv[(size_t)b * Z + u] = l; z[(size_t)b * Z + u] = zz;
        if (mu_out) mu_out[(size_t)b * Z + u] = m;
        if (lv_out) lv_out[(size_t)b * Z + u] = l;
    }
    __syncthreads();
    float s = w.l2h_b[u];
    for (int i = 0; i < Z; ++i) s = fmaf(w.l2h_w[(size_t)u * Z + i], zs[i], s);
    h0[(size_t)b * H + u] = tanhf(s);
}

// dh0 [B][H], upstream d_mu/d_lv [B][Z] (nullable) -> dpre, dmu_t, dlv_t, dy, dhn.
template <int H>
__global__ void __launch_bounds__(H) head_bwd_kernel(const float* __restrict__ dh0, const float* __restrict__ d_mu,
                                                     const float* __restrict__ d_lv, const float* __restrict__ eps, HeadW w,
                                                     const float* __restrict__ xn, const float* __restrict__ rstd,
                                                     const float* __restrict__ lv, const float* __restrict__ h0,
                                                     float* __restrict__ dpre, float* __restrict__ dmu_t, float* __restrict__ dlv_t,
                                                     float* __restrict__ dy, float* __restrict__ dhn) {
    __shared__ float red[4], ps[H], dms[16], dls[16];
    const int b = blockIdx.x, u = threadIdx.x, Z = w.Z;
    const float hv = h0[(size_t)b * H + u];
    const float dp = dh0[(size_t)b * H + u] * (1.f - hv * hv);
    dpre[(size_t)b * H + u] = dp;
    ps[u] = dp;
    __syncthreads();
    const int warp = u >> 5, lane = u & 31;
    for (int o = warp; o < Z; o += H / 32) {
        float s = 0.f;
        for (int kk = lane; kk < H; kk += 32) s = fmaf(ps[kk], w.l2h_w[(size_t)kk * Z + o], s);
        s = warp_sum(s);                                    // dz[o]
        if (lane == 0) {
            const float l = lv[(size_t)b * Z + o];
            const float dm = s + (d_mu ? d_mu[(size_t)b * Z + o] : 0.f);
            const float dl = fmaf(s * eps[(size_t)b * Z + o], 0.5f * expf(0.5f * l), d_lv ? d_lv[(size_t)b * Z + o] : 0.f);
            dms[o] = dm; dls[o] = dl;
            dmu_t[(size_t)b * Z + o] = dm; dlv_t[(size_t)b * Z + o] = dl;
        }
    }
    __syncthreads();
    float d = 0.f;
    for (int i = 0; i < Z; ++i) d = fmaf(dms[i], w.mu_w[(size_t)i * H + u], fmaf(dls[i], w.lv_w[(size_t)i * H + u], d));
    dy[(size_t)b * H + u] = d;
    if (w.has_ln) {
        const float x = xn[(size_t)b * H + u];
        const float dxn = d * w.ln_w[u];
        const float m1 = block_sum<H>(dxn, red) * (1.f / H);
        const float m2 = block_sum<H>(dxn * x, red) * (1.f / H);
        dhn[(size_t)b * H + u] = rstd[b] * (dxn - m1 - x * m2);
    } else {
        dhn[(size_t)b * H + u] = d;
    }
}

// ------------------------------------------------------------------------------------------------------------
// ELBO (03_train_vae.py:264-266): recon = mean (xhat-x)^2, kl = -0.5 mean(1 + lv - mu^2 - e^lv),
// loss = recon + kl_w*kl; upstream gradients for the backward pass.  acc = loss3 (device float[3], zeroed first).
// ------------------------------------------------------------------------------------------------------------
__global__ void elbo_kernel(const float* __restrict__ x, const float* __restrict__ xhat, size_t n_x, const float* __restrict__ mu,
                            const float* __restrict__ lv, size_t n_z, float kl_w, float* __restrict__ d_xhat,
                            float* __restrict__ d_mu, float* __restrict__ d_lv, float* __restrict__ loss3) {
    __shared__ double rs[8], ks[8];
    double r = 0.0, kk = 0.0;
    const float gx = 2.f / (float)n_x, gz = kl_w / (float)n_z;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_x; i += (size_t)gridDim.x * blockDim.x) {
        const float d = xhat[i] - x[i];
        r += (double)d * d;
        if (d_xhat) d_xhat[i] = gx * d;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_z; i += (size_t)gridDim.x * blockDim.x) {
        const float m = mu[i], l = lv[i], e = expf(l);
        kk += (double)(1.f + l - m * m - e);
        if (d_mu) d_mu[i] = gz * m;
        if (d_lv) d_lv[i] = -0.5f * gz * (1.f - e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { r += __shfl_xor_sync(0xffffffffu, r, o); kk += __shfl_xor_sync(0xffffffffu, kk, o); }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = r; ks[threadIdx.x >> 5] = kk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, c = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { a += rs[i]; c += ks[i]; }
        const float recon = (float)(a / (double)n_x), kl = (float)(-0.5 * c / (double)n_z);
        atomicAdd(loss3 + 1, recon);
        atomicAdd(loss3 + 2, kl);
        atomicAdd(loss3 + 0, recon + kl_w * kl);
    }
}

// ------------------------------------------------------------------------------------------------------------
// clip_grad_norm_(max_norm) + Adam(weight_decay as L2-in-grad) (03_train_vae.py:222,269-270).
// ------------------------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ acc) {
    __shared__ double s[8];
    double v = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) v += (double)g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) a += s[i];
        atomicAdd(acc, (float)a);
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr, float b1, float b2, float omb1, float omb2, float eps, float wd, int decoupled,
                            float max_norm, float grad_scale, float bc1, float bc2_sqrt, float* __restrict__ norm2) {
    const float total_norm = sqrtf(norm2[0]) * grad_scale;
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(1.f, max_norm / (total_norm + 1e-6f));
    const float gs = grad_scale * coef;
    const float step = lr / bc1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float pv = p[i];
        float gr = g[i] * gs;
        if (decoupled) pv *= 1.f - lr * wd;                          // AdamW: param.mul_(1 - lr * weight_decay)
        else gr = fmaf(wd, pv, gr);                                  // Adam: weight decay added to the gradient
        const float mv = fmaf(b1, m[i], omb1 * gr);                 // exp_avg.lerp_(grad, 1-beta1); 1-beta in double like torch
        const float vv = fmaf(b2, v[i], omb2 * gr * gr);
        m[i] = mv; v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[i] = pv - step * (mv / denom);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) norm2[1] = total_norm;
}

}  // namespace shm

using namespace shm;

// ------------------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------------------
struct shm_vae_trainer {
    shm_vae_cfg cfg;
    int T, Bmax, device, nsm;
    ParamLayout pl;
    float* ws;
    uint8_t* masks;                 // [2][L-1][Bmax][T][H] copies of the caller's keep masks
    size_t ws_floats;
    // workspace offsets (floats)
    size_t x_tm, xhat_tm, G[2][SHM_MAX_L], hs[2][SHM_MAX_L], cs[2][SHM_MAX_L], hd[2][SHM_MAX_L], dHa, dHb;
    size_t xn, rstd, y, mu, lv, z, h0, eps, gconst, dgsum, dh0, dpre, dmu_t, dlv_t, dy, dhn;
    // state of the last forward
    int B, have_fwd, use_mask;
    float scale;
};

template <int H>
static int launch_rec_fwd(cudaStream_t st, int nsm, const float* w_hh, const float* gx, long long tstride, float* act, float* hs,
                          float* cs, float* hd, const uint8_t* mask, float scale, int T, int B) {
    constexpr int KS = RecCfg<H>::KS;
    if constexpr (H == 128) {
        if (g_rec_cluster) {                       // two CTAs per cluster, W_hh in registers (see lstm_rec_fwd_c2_kernel)
            const int pairs = nsm / 2;
            const int nwc = B <= pairs ? 1 : (B <= 2 * pairs ? 2 : 4);
            const int gridc = 2 * ((B + nwc - 1) / nwc);
            if (nwc == 1) lstm_rec_fwd_c2_kernel<1><<<gridc, 256, 0, st>>>(w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
            else if (nwc == 2) lstm_rec_fwd_c2_kernel<2><<<gridc, 256, 0, st>>>(w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
            else lstm_rec_fwd_c2_kernel<4><<<gridc, 256, 0, st>>>(w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
            SHM_LAUNCH_CHECK();
            return SHM_OK;
        }
    }
    const int nw = B <= nsm ? 1 : (B <= 2 * nsm ? 2 : 4);
    const size_t smem = (size_t)(KS / 4) * 4 * H * sizeof(float4) + (size_t)2 * nw * H * sizeof(float);
    const int grid = (B + nw - 1) / nw;
    constexpr int kThreads = 2 * H;
#define SHM_REC_FWD(NW)                                                                                                         \
    do {                                                                                                                        \
        SHM_CUDA(cudaFuncSetAttribute(lstm_rec_fwd_kernel<H, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        lstm_rec_fwd_kernel<H, NW><<<grid, kThreads, smem, st>>>(w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);           \
    } while (0)
    if (nw == 1) SHM_REC_FWD(1); else if (nw == 2) SHM_REC_FWD(2); else SHM_REC_FWD(4);
#undef SHM_REC_FWD
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

template <int H>
static int launch_rec_bwd(cudaStream_t st, int nsm, const float* w_hh, float* act_dg, const float* cs, const float* dH,
                          const uint8_t* mask, float scale, const float* dh_last, float* dgsum, int T, int B) {
    constexpr int MS = RecCfg<H>::MS;
    if constexpr (H == 128) {
        if (g_rec_cluster) {
            const int pairs = nsm / 2;
            const int nwc = B <= pairs ? 1 : (B <= 2 * pairs ? 2 : 4);
            const int gridc = 2 * ((B + nwc - 1) / nwc);
            if (nwc == 1) lstm_rec_bwd_c2_kernel<1><<<gridc, 256, 0, st>>>(w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
            else if (nwc == 2) lstm_rec_bwd_c2_kernel<2><<<gridc, 256, 0, st>>>(w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
            else lstm_rec_bwd_c2_kernel<4><<<gridc, 256, 0, st>>>(w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
            SHM_LAUNCH_CHECK();
            return SHM_OK;
        }
    }
    const int nw = B <= nsm ? 1 : (B <= 2 * nsm ? 2 : 4);
    const size_t smem = (size_t)(MS / 4) * 2 * H * sizeof(float4) + (size_t)(nw * 4 * H + 2 * nw * H) * sizeof(float);
    const int grid = (B + nw - 1) / nw;
    constexpr int kThreads = 2 * H;
#define SHM_REC_BWD(NW)                                                                                                         \
    do {                                                                                                                        \
        SHM_CUDA(cudaFuncSetAttribute(lstm_rec_bwd_kernel<H, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
        lstm_rec_bwd_kernel<H, NW><<<grid, kThreads, smem, st>>>(w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);         \
    } while (0)
    if (nw == 1) SHM_REC_BWD(1); else if (nw == 2) SHM_REC_BWD(2); else SHM_REC_BWD(4);
#undef SHM_REC_BWD
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

static int rec_fwd(int H, cudaStream_t st, int nsm, const float* w_hh, const float* gx, long long tstride, float* act, float* hs,
                   float* cs, float* hd, const uint8_t* mask, float scale, int T, int B) {
    switch (H) {
        case 32: return launch_rec_fwd<32>(st, nsm, w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
        case 64: return launch_rec_fwd<64>(st, nsm, w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
        case 128: return launch_rec_fwd<128>(st, nsm, w_hh, gx, tstride, act, hs, cs, hd, mask, scale, T, B);
    }
    return SHM_ERR_UNSUPPORTED;
}

static int rec_bwd(int H, cudaStream_t st, int nsm, const float* w_hh, float* act_dg, const float* cs, const float* dH,
                   const uint8_t* mask, float scale, const float* dh_last, float* dgsum, int T, int B) {
    switch (H) {
        case 32: return launch_rec_bwd<32>(st, nsm, w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
        case 64: return launch_rec_bwd<64>(st, nsm, w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
        case 128: return launch_rec_bwd<128>(st, nsm, w_hh, act_dg, cs, dH, mask, scale, dh_last, dgsum, T, B);
    }
    return SHM_ERR_UNSUPPORTED;
}

static HeadW head_weights(const shm_vae_trainer* h, const float* p) {
    HeadW w;
    w.ln_w = h->cfg.has_ln ? p + h->pl.ln_w : nullptr;
    w.ln_b = h->cfg.has_ln ? p + h->pl.ln_b : nullptr;
    w.mu_w = p + h->pl.mu_w; w.mu_b = p + h->pl.mu_b;
    w.lv_w = p + h->pl.lv_w; w.lv_b = p + h->pl.lv_b;
    w.l2h_w = p + h->pl.l2h_w; w.l2h_b = p + h->pl.l2h_b;
    w.has_ln = h->cfg.has_ln; w.Z = h->cfg.Z; w.ln_eps = h->cfg.ln_eps;
    return w;
}

extern "C" int64_t shm_vae_param_count(const shm_vae_cfg* cfg) {
    if (!cfg || cfg->L < 1 || cfg->L > SHM_MAX_L) return SHM_ERR_ARG;
    return (int64_t)param_layout(*cfg).total;
}

extern "C" int shm_vae_trainer_create(shm_vae_trainer** out, const shm_vae_cfg* cfg, int32_t T, int32_t max_batch, int device) {
    if (!out || !cfg) return SHM_ERR_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != SHM_OK) return rc;
    const int D = cfg->D, H = cfg->H, Z = cfg->Z, L = cfg->L;
    if (D < 1 || Z < 1 || L < 1 || T < 1 || max_batch < 1) return SHM_ERR_ARG;
    if (!(H == 32 || H == 64 || H == 128) || L > SHM_MAX_L || D > SHM_MAX_D || Z > 16) return SHM_ERR_UNSUPPORTED;
    int prev = 0;
    SHM_CUDA(cudaGetDevice(&prev));
    SHM_CUDA(cudaSetDevice(device));
    shm_vae_trainer* h = new (std::nothrow) shm_vae_trainer();
    if (!h) { cudaSetDevice(prev); return SHM_ERR_NOMEM; }
    h->cfg = *cfg; h->T = T; h->Bmax = max_batch; h->device = device; h->nsm = device_sm_count(device);
    h->pl = param_layout(*cfg);
    h->have_fwd = 0; h->B = 0; h->use_mask = 0; h->scale = 1.f;
    const size_t B = max_batch, TB = (size_t)T * B;
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 3) / 4 * 4; return r; };     // keep 16-byte alignment
    h->x_tm = take(TB * D); h->xhat_tm = take(TB * D);
    for (int s = 0; s < 2; ++s)
        for (int l = 0; l < L; ++l) {
            h->G[s][l] = take(TB * 4 * H);
            h->hs[s][l] = take((TB + B) * H);
            h->cs[s][l] = take((TB + B) * H);
            h->hd[s][l] = (l < L - 1) ? take(TB * H) : 0;
        }
    h->dHa = take(TB * H); h->dHb = take(TB * H);
    h->xn = take(B * H); h->rstd = take(B); h->y = take(B * H); h->mu = take(B * Z); h->lv = take(B * Z); h->z = take(B * Z);
    h->h0 = take(B * H); h->eps = take(B * Z); h->gconst = take(B * 4 * H); h->dgsum = take(B * 4 * H); h->dh0 = take(B * H);
    h->dpre = take(B * H); h->dmu_t = take(B * Z); h->dlv_t = take(B * Z); h->dy = take(B * H); h->dhn = take(B * H);
    h->ws_floats = o;
    h->ws = nullptr; h->masks = nullptr;
    if (cudaMalloc(&h->ws, o * sizeof(float)) != cudaSuccess) { cudaGetLastError(); delete h; cudaSetDevice(prev); return SHM_ERR_NOMEM; }
    if (L > 1 && cudaMalloc(&h->masks, (size_t)2 * (L - 1) * TB * H) != cudaSuccess) {
        cudaGetLastError(); cudaFree(h->ws); delete h; cudaSetDevice(prev); return SHM_ERR_NOMEM;
    }
    cudaSetDevice(prev);
    *out = h;
    return SHM_OK;
}

extern "C" int shm_vae_trainer_destroy(shm_vae_trainer* h) {
    if (!h) return SHM_OK;
    int prev = 0;
    const bool have_prev = cudaGetDevice(&prev) == cudaSuccess;
    cudaSetDevice(h->device);
    cudaFree(h->ws);
    if (h->masks) cudaFree(h->masks);
    delete h;
    if (have_prev) cudaSetDevice(prev);
    return SHM_OK;
}

extern "C" int shm_vae_train_forward(shm_vae_trainer* h, const float* params, const float* x, int32_t B, const float* eps,
                                     const uint8_t* drop_enc, const uint8_t* drop_dec, float drop_p, float* xhat, float* mu,
                                     float* logvar, void* stream) {
    if (!h || !params || !x || !eps || B < 1) return SHM_ERR_ARG;
    if (B > h->Bmax) return SHM_ERR_ARG;
    if ((drop_enc == nullptr) != (drop_dec == nullptr)) return SHM_ERR_ARG;
    if (drop_enc && !(drop_p >= 0.f && drop_p < 1.f)) return SHM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->cfg.D, H = h->cfg.H, Z = h->cfg.Z, L = h->cfg.L, T = h->T;
    const size_t TB = (size_t)T * B;
    float* ws = h->ws;
    const ParamLayout& pl = h->pl;
    h->B = B; h->have_fwd = 0;
    h->use_mask = (drop_enc && L > 1 && drop_p > 0.f) ? 1 : 0;
    h->scale = h->use_mask ? 1.f / (1.f - drop_p) : 1.f;
    const size_t mask_layer = TB * H;
    if (h->use_mask) {
        SHM_CUDA(cudaMemcpyAsync(h->masks, drop_enc, (size_t)(L - 1) * mask_layer, cudaMemcpyDeviceToDevice, st));
        SHM_CUDA(cudaMemcpyAsync(h->masks + (size_t)(L - 1) * mask_layer, drop_dec, (size_t)(L - 1) * mask_layer,
                                 cudaMemcpyDeviceToDevice, st));
    }
    SHM_CUDA(cudaMemcpyAsync(ws + h->eps, eps, (size_t)B * Z * sizeof(float), cudaMemcpyDeviceToDevice, st));
    swap01_kernel<<<296, 256, 0, st>>>(x, ws + h->x_tm, B, T, D);     // [B][T][D] -> [T][B][D]
    SHM_LAUNCH_CHECK();
    int rc;
    // encoder stack (temporal_vae.py:53-54)
    for (int l = 0; l < L; ++l) {
        const float* in = l == 0 ? ws + h->x_tm : (ws + h->hd[0][l - 1]);
        const int K = l == 0 ? D : H;
        if ((rc = gemm(st, in, K, 1, params + pl.enc_wih[l], 1, K, ws + h->G[0][l], 4 * H, (int)TB, 4 * H, K,
                       params + pl.enc_bih[l], params + pl.enc_bhh[l], false, SHM_GEMM_TC_F16X3))) return rc;
        const bool last = l == L - 1;
        const uint8_t* mk = (!last && h->use_mask) ? h->masks + (size_t)l * mask_layer : nullptr;
        if ((rc = rec_fwd(H, st, h->nsm, params + pl.enc_whh[l], ws + h->G[0][l], (long long)B * 4 * H, ws + h->G[0][l],
                          ws + h->hs[0][l], ws + h->cs[0][l], last ? nullptr : ws + h->hd[0][l], mk, h->scale, T, B))) return rc;
    }
    // heads (temporal_vae.py:55-63,66)
    const float* hn = ws + h->hs[0][L - 1] + TB * H;
    const HeadW hw = head_weights(h, params);
#define SHM_HEAD_FWD(HH)                                                                                                       \
    head_fwd_kernel<HH><<<B, HH, 0, st>>>(hn, ws + h->eps, hw, ws + h->xn, ws + h->rstd, ws + h->y, ws + h->mu, ws + h->lv,    \
                                          ws + h->z, ws + h->h0, mu, logvar)
    if (H == 32) SHM_HEAD_FWD(32); else if (H == 64) SHM_HEAD_FWD(64); else SHM_HEAD_FWD(128);
#undef SHM_HEAD_FWD
    SHM_LAUNCH_CHECK();
    // decoder stack: layer 0 sees the same input h0 at every step (temporal_vae.py:67-69), so its input projection
    // is computed once per window
    for (int l = 0; l < L; ++l) {
        const bool last = l == L - 1;
        const uint8_t* mk = (!last && h->use_mask) ? h->masks + (size_t)(L - 1 + l) * mask_layer : nullptr;
        if (l == 0) {
            if ((rc = gemm(st, ws + h->h0, H, 1, params + pl.dec_wih[0], 1, H, ws + h->gconst, 4 * H, B, 4 * H, H,
                           params + pl.dec_bih[0], params + pl.dec_bhh[0], false))) return rc;
            if ((rc = rec_fwd(H, st, h->nsm, params + pl.dec_whh[0], ws + h->gconst, 0, ws + h->G[1][0], ws + h->hs[1][0],
                              ws + h->cs[1][0], last ? nullptr : ws + h->hd[1][0], mk, h->scale, T, B))) return rc;
        } else {
            if ((rc = gemm(st, ws + h->hd[1][l - 1], H, 1, params + pl.dec_wih[l], 1, H, ws + h->G[1][l], 4 * H, (int)TB, 4 * H, H,
                           params + pl.dec_bih[l], params + pl.dec_bhh[l], false, SHM_GEMM_TC_F16X3))) return rc;
            if ((rc = rec_fwd(H, st, h->nsm, params + pl.dec_whh[l], ws + h->G[1][l], (long long)B * 4 * H, ws + h->G[1][l],
                              ws + h->hs[1][l], ws + h->cs[1][l], last ? nullptr : ws + h->hd[1][l], mk, h->scale, T, B))) return rc;
        }
    }
    // output layer (temporal_vae.py:70)
    const float* htop = ws + h->hs[1][L - 1] + (size_t)B * H;
    if ((rc = gemm(st, htop, H, 1, params + pl.out_w, 1, H, ws + h->xhat_tm, D, (int)TB, D, H, params + pl.out_b, nullptr, false)))
        return rc;
    if (xhat) {
        swap01_kernel<<<296, 256, 0, st>>>(ws + h->xhat_tm, xhat, T, B, D);   // [T][B][D] -> [B][T][D]
        SHM_LAUNCH_CHECK();
    }
    h->have_fwd = 1;
    return SHM_OK;
}

extern "C" int shm_vae_train_backward(shm_vae_trainer* h, const float* params, const float* d_xhat, const float* d_mu,
                                      const float* d_logvar, float* grads, void* stream) {
    if (!h || !params || !d_xhat || !grads) return SHM_ERR_ARG;
    if (!h->have_fwd) return SHM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int D = h->cfg.D, H = h->cfg.H, Z = h->cfg.Z, L = h->cfg.L, T = h->T, B = h->B;
    const size_t TB = (size_t)T * B;
    float* ws = h->ws;
    const ParamLayout& pl = h->pl;
    const size_t mask_layer = TB * H;
    int rc;
    SHM_CUDA(cudaMemsetAsync(grads, 0, pl.total * sizeof(float), st));
    // output layer
    float* dx_tm = ws + h->x_tm + 0;      // x_tm is still needed (encoder dW_ih): use xhat_tm's slot for the upstream gradient
    dx_tm = ws + h->xhat_tm;
    swap01_kernel<<<296, 256, 0, st>>>(d_xhat, dx_tm, B, T, D);
    SHM_LAUNCH_CHECK();
    const float* htop = ws + h->hs[1][L - 1] + (size_t)B * H;
    if ((rc = gemm(st, dx_tm, 1, D, htop, H, 1, grads + pl.out_w, H, D, H, (int)TB, nullptr, nullptr, true, SHM_GEMM_TC_BF16X3))) return rc;
    if ((rc = colsum(st, dx_tm, nullptr, (int)TB, D, grads + pl.out_b, nullptr))) return rc;
    float* dH = ws + h->dHa;
    float* dH2 = ws + h->dHb;
    if ((rc = gemm(st, dx_tm, D, 1, params + pl.out_w, H, 1, dH, H, (int)TB, H, D, nullptr, nullptr, false, SHM_GEMM_TC_BF16X3))) return rc;
    // decoder stack, top down
    for (int l = L - 1; l >= 0; --l) {
        // dH is the gradient w.r.t. this layer's (dropped, if not the top) output sequence
        const uint8_t* mk = (l < L - 1 && h->use_mask) ? h->masks + (size_t)(L - 1 + l) * mask_layer : nullptr;
        float* dG = ws + h->G[1][l];
        if ((rc = rec_bwd(H, st, h->nsm, params + pl.dec_whh[l], dG, ws + h->cs[1][l], dH, mk, h->scale, nullptr,
                          l == 0 ? ws + h->dgsum : nullptr, T, B))) return rc;
        if ((rc = gemm(st, dG, 1, 4 * H, ws + h->hs[1][l], H, 1, grads + pl.dec_whh[l], H, 4 * H, H, (int)TB, nullptr, nullptr, true, SHM_GEMM_TC_BF16X3)))
            return rc;
        if ((rc = colsum(st, dG, nullptr, (int)TB, 4 * H, grads + pl.dec_bih[l], grads + pl.dec_bhh[l]))) return rc;
        if (l > 0) {
            if ((rc = gemm(st, dG, 1, 4 * H, ws + h->hd[1][l - 1], H, 1, grads + pl.dec_wih[l], H, 4 * H, H, (int)TB, nullptr, nullptr, true, SHM_GEMM_TC_BF16X3)))
                return rc;
            if ((rc = gemm(st, dG, 4 * H, 1, params + pl.dec_wih[l], H, 1, dH2, H, (int)TB, H, 4 * H, nullptr, nullptr, false, SHM_GEMM_TC_BF16X3))) return rc;
            float* t = dH; dH = dH2; dH2 = t;
        } else {
            if ((rc = gemm(st, ws + h->dgsum, 1, 4 * H, ws + h->h0, H, 1, grads + pl.dec_wih[0], H, 4 * H, H, B, nullptr, nullptr, true)))
                return rc;
            if ((rc = gemm(st, ws + h->dgsum, 4 * H, 1, params + pl.dec_wih[0], H, 1, ws + h->dh0, H, B, H, 4 * H, nullptr, nullptr, false)))
                return rc;
        }
    }
    // heads
    const HeadW hw = head_weights(h, params);
#define SHM_HEAD_BWD(HH)                                                                                                       \
    head_bwd_kernel<HH><<<B, HH, 0, st>>>(ws + h->dh0, d_mu, d_logvar, ws + h->eps, hw, ws + h->xn, ws + h->rstd, ws + h->lv,  \
                                          ws + h->h0, ws + h->dpre, ws + h->dmu_t, ws + h->dlv_t, ws + h->dy, ws + h->dhn)
    if (H == 32) SHM_HEAD_BWD(32); else if (H == 64) SHM_HEAD_BWD(64); else SHM_HEAD_BWD(128);
#undef SHM_HEAD_BWD
    SHM_LAUNCH_CHECK();
    if ((rc = gemm(st, ws + h->dpre, 1, H, ws + h->z, Z, 1, grads + pl.l2h_w, Z, H, Z, B, nullptr, nullptr, true))) return rc;
    if ((rc = colsum(st, ws + h->dpre, nullptr, B, H, grads + pl.l2h_b, nullptr))) return rc;
    if ((rc = gemm(st, ws + h->dmu_t, 1, Z, ws + h->y, H, 1, grads + pl.mu_w, H, Z, H, B, nullptr, nullptr, true))) return rc;
    if ((rc = colsum(st, ws + h->dmu_t, nullptr, B, Z, grads + pl.mu_b, nullptr))) return rc;
    if ((rc = gemm(st, ws + h->dlv_t, 1, Z, ws + h->y, H, 1, grads + pl.lv_w, H, Z, H, B, nullptr, nullptr, true))) return rc;
    if ((rc = colsum(st, ws + h->dlv_t, nullptr, B, Z, grads + pl.lv_b, nullptr))) return rc;
    if (h->cfg.has_ln) {
        if ((rc = colsum(st, ws + h->dy, ws + h->xn, B, H, grads + pl.ln_w, nullptr))) return rc;
        if ((rc = colsum(st, ws + h->dy, nullptr, B, H, grads + pl.ln_b, nullptr))) return rc;
    }
    // encoder stack, top down: only the top layer's final hidden state feeds the heads (temporal_vae.py:54)
    const float* dHin = nullptr;
    for (int l = L - 1; l >= 0; --l) {
        const uint8_t* mk = (l < L - 1 && h->use_mask) ? h->masks + (size_t)l * mask_layer : nullptr;
        float* dG = ws + h->G[0][l];
        if ((rc = rec_bwd(H, st, h->nsm, params + pl.enc_whh[l], dG, ws + h->cs[0][l], dHin, mk, h->scale,
                          l == L - 1 ? ws + h->dhn : nullptr, nullptr, T, B))) return rc;
        if ((rc = gemm(st, dG, 1, 4 * H, ws + h->hs[0][l], H, 1, grads + pl.enc_whh[l], H, 4 * H, H, (int)TB, nullptr, nullptr, true, SHM_GEMM_TC_BF16X3)))
            return rc;
        if ((rc = colsum(st, dG, nullptr, (int)TB, 4 * H, grads + pl.enc_bih[l], grads + pl.enc_bhh[l]))) return rc;
        if (l > 0) {
            if ((rc = gemm(st, dG, 1, 4 * H, ws + h->hd[0][l - 1], H, 1, grads + pl.enc_wih[l], H, 4 * H, H, (int)TB, nullptr, nullptr, true, SHM_GEMM_TC_BF16X3)))
                return rc;
            if ((rc = gemm(st, dG, 4 * H, 1, params + pl.enc_wih[l], H, 1, ws + h->dHa, H, (int)TB, H, 4 * H, nullptr, nullptr, false, SHM_GEMM_TC_BF16X3)))
                return rc;
            dHin = ws + h->dHa;
        } else {
            if ((rc = gemm(st, dG, 1, 4 * H, ws + h->x_tm, D, 1, grads + pl.enc_wih[0], D, 4 * H, D, (int)TB, nullptr, nullptr, true)))
                return rc;
        }
    }
    h->have_fwd = 0;      // the saved gate activations were overwritten by their gradients
    return SHM_OK;
}

extern "C" int shm_vae_elbo_grad(const float* x, const float* xhat, const float* mu, const float* logvar, int64_t n_x, int64_t n_z,
                                 float kl_w, float* d_xhat, float* d_mu, float* d_logvar, float* loss3, void* stream) {
    if (!x || !xhat || !mu || !logvar || !loss3 || n_x < 1 || n_z < 1) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SHM_CUDA(cudaMemsetAsync(loss3, 0, 3 * sizeof(float), st));
    const int grid = (int)((n_x + 255) / 256 < 296 ? (n_x + 255) / 256 : 296);
    elbo_kernel<<<grid, 256, 0, st>>>(x, xhat, (size_t)n_x, mu, logvar, (size_t)n_z, kl_w, d_xhat, d_mu, d_logvar, loss3);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

namespace shm {
int adam_step(cudaStream_t st, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, int step, float lr,
              float beta1, float beta2, float eps, float weight_decay, int decoupled, float max_norm, float grad_scale, float* norm2) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !norm2 || n < 1 || step < 1) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    SHM_CUDA(cudaMemsetAsync(norm2, 0, 2 * sizeof(float), st));
    const int grid = (int)((n + 255) / 256 < 296 ? (n + 255) / 256 : 296);
    sumsq_kernel<<<grid, 256, 0, st>>>(grads, (size_t)n, norm2);
    SHM_LAUNCH_CHECK();
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<grid, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (size_t)n, lr, beta1, beta2, (float)(1.0 - (double)beta1),
                                      (float)(1.0 - (double)beta2), eps, weight_decay, decoupled, max_norm, grad_scale, (float)bc1,
                                      (float)sqrt(bc2), norm2);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
}  // namespace shm

extern "C" int shm_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                  float grad_scale, float* norm2, void* stream) {
    return shm::adam_step((cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay, 0, max_norm,
                          grad_scale, norm2);
}

extern "C" int shm_adamw_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                                   float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm,
                                   float grad_scale, float* norm2, void* stream) {
    return shm::adam_step((cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, weight_decay, 1, max_norm,
                          grad_scale, norm2);
}

extern "C" int shm_train_set_tensor_cores(int enable) {
    // 0: fp32 FMA contractions + single-CTA recurrence (the r01 engines); 1: both r02 engines (default);
    // 2: cluster recurrence only; 3: same as 1; 5: tensor-core contractions only
    shm::g_train_tc = (enable == 1 || enable == 3 || enable == 5) ? 1 : 0;
    shm::g_rec_cluster = (enable == 1 || enable == 2 || enable == 3) ? 1 : 0;
    return SHM_OK;
}
