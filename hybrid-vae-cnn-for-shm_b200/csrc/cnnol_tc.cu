// Tensor-core engine of the openLAB attribution CNN (20250506_openLAB_tests/Codes/Models/cnn_model.py:16-43,54-57,
// eval mode), used by shm_cnnol_forward for the flagged windows of 10_test_hybrid_pipeline.py:265-302.
//
// Blocks 2..4 hold 99 % of the 133.9 MFLOP per window; each is an implicit GEMM on the 5th-gen tensor cores
// (tcgen05.mma kind::f16, fp32 accumulators in TMEM) with the same 3-pass fp16 hi/lo split as the LSTM scorer, so the
// logits keep fp32-grade accuracy (1e-4 tolerance):
//
//   rows  m = (h_local, window-in-group, w): 128 rows = HT time steps x WPT windows x 4 sensor columns,
//   cols  n = output channel (64 / 128 / 256),    K = (ci chunk of 16) x (dt, df) taps.
//
// The A operand is never im2col-expanded: the activated + pooled input of a tile PAIR (2*HT + KT-1 time steps) is stored
// ONCE per ci-chunk as UMMA K-major core-matrix images -- three copies pre-shifted by df = -1/0/+1 in the sensor dimension
// with zero fill, hi and lo halves -- and every (dt, df) tap is the SAME image addressed dt*RPH rows further down
// (RPH = rows per time step is a multiple of 8, so a time shift is a whole number of 8-row core-matrix groups).
// Both tiles of a pair consume each 16 KB weight stage (6 MMAs), which halves the L2->SM weight stream.
//
// Per block:  act_stage (GroupNorm affine + SiLU + MaxPool(2,1) + fp16 split of the previous block's raw output, written
// straight into the image layout)  ->  conv_gemm (persistent, warp-specialised: 2 bulk-copy producers, 1 MMA issuer,
// 4 epilogue warps: +bias, channels-last store, GroupNorm sum/sumsq in fp64 atomics).  Block 1 (1->32 channels, 0.8 % of
// the FLOPs) and the GAP + classifier tail run on the CUDA cores.
#include "cnnol_tc.cuh"
#include "tcgen05.cuh"

namespace shm {
using namespace tc;

namespace {

constexpr int OLT_T = 200, OLT_F = 4;

// geometry of conv blocks 2..4
template <int L> struct OlGeo;
template <> struct OlGeo<1> { static constexpr int CIN = 32, NOUT = 64, KT = 5, HIN = 100, RPH = 16; };
template <> struct OlGeo<2> { static constexpr int CIN = 64, NOUT = 128, KT = 5, HIN = 50, RPH = 32; };
template <> struct OlGeo<3> { static constexpr int CIN = 128, NOUT = 256, KT = 3, HIN = 25, RPH = 64; };

template <int L> struct OlDer {
    using G = OlGeo<L>;
    static constexpr int WPT = G::RPH / 4;                 // windows per tile
    static constexpr int HT = 128 / G::RPH;                // time steps per 128-row tile
    static constexpr int PH = 2 * HT;                      // time steps per tile pair
    static constexpr int HH = PH + G::KT - 1;              // staged time steps (with halo)
    static constexpr int R = HH * G::RPH;                  // staged rows
    static constexpr int NPAIR = (G::HIN + PH - 1) / PH;
    static constexpr int NCC = G::CIN / 16;
    static constexpr int A_IMG = R * 32;                   // bytes: R rows x 16 fp16 (2 k-core blocks of R x 16 B: row r at r*16)
    static constexpr int A_STAGE = 6 * A_IMG;              // shared memory: 3 df shifts x {hi, lo}
    static constexpr int A_GSTAGE = 2 * A_IMG;             // global memory: the unshifted image only, {hi, lo}
    static constexpr int TG = (G::NOUT <= 64) ? 3 : 1;     // taps per weight stage (small-N MMAs are short: fewer barrier round trips)
    static constexpr int B_TAP = G::NOUT * 64;             // {hi, lo} x [NOUT x 16] fp16 of one tap
    static constexpr int B_STAGE = TG * B_TAP;
    static constexpr int NB = 4;                           // weight ring stages (8 measured slower: the shared memory it takes is L1 the fused producer's loads need)
    static constexpr int NACC = (4 * G::NOUT <= 512) ? 2 : 1;   // accumulator sets (pair = 2*NOUT columns): double buffered when TMEM allows
    static constexpr int OFF_BAR = 2 * A_STAGE + NB * B_STAGE;
    static constexpr int OFF_BIAS = OFF_BAR + 256;         // fp32 bias[NOUT]
    static constexpr int OFF_RED = OFF_BIAS + G::NOUT * 4; // fp64 [4 warps][8 windows][16] GroupNorm partials
    static constexpr int SMEM = OFF_RED + 4 * 8 * 16 * 8;
    static constexpr int TCOLS = (2 * NACC * G::NOUT <= 128) ? 128 : ((2 * NACC * G::NOUT <= 256) ? 256 : 512);
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert((G::KT * 3) % TG == 0, "tap groups");
};

__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }
// ex2.approx (2 ulp) + rcp.approx (1 ulp): ~3e-7 relative, 4x fewer instructions; used where SiLU is the whole kernel
__device__ __forceinline__ float silu_fast(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(-1.4426950408889634f * x, 80.f)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

// packed fp32 pairs (sm_100 fma.rn.f32x2 -> FFMA2: two lanes per issue slot, same flops per clock; scripts/ffma2_probe.cu)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ long long eff_windows(long long n_total, const int* n_dev, long long base, long long n_chunk) {
    long long e = n_total;
    if (n_dev) e = min(e, (long long)__ldg(n_dev));
    e -= base;
    return e < 0 ? 0 : (e > n_chunk ? n_chunk : e);
}

// ------------------------------------------------------------------------------------------------------------
// block 1: gather + standardise (10_test_hybrid_pipeline.py:272-278) + Conv(7x3, 1->32) on the CUDA cores.
// out: channels-last [w][200][4][32] fp32 (+bias); stats[w][8][2] = sum, sum of squares per GroupNorm group (fp64).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ol_conv1_kernel(const float* __restrict__ w1, const float* __restrict__ b1, WinSrc src,
                                                       const int* __restrict__ idx, const int* __restrict__ n_dev, long long n_total,
                                                       long long base, long long n_chunk, float* __restrict__ out,
                                                       double* __restrict__ stats) {
    // thread item = (time step h, block of 4 output channels = one GroupNorm group) for all 4 sensor columns: 16 accumulators,
    // per dt one padded input row (6 floats) and 3 float4 weight vectors feed 48 FMAs
    __shared__ __align__(16) float xin[(OLT_T + 6) * 8];    // zero-padded plane: rows -3..202, cols -1..4 (+2 pad)
    __shared__ __align__(16) float wk[21 * 32];             // [tap][co]
    __shared__ float bs[32];
    __shared__ float red[8][16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cb = tid & 7;                                 // channel block / GroupNorm group
    const long long n_eff = eff_windows(n_total, n_dev, base, n_chunk);
    for (int i = tid; i < 21 * 32; i += 256) { const int co = i & 31, tap = i >> 5; wk[i] = __ldg(w1 + co * 21 + tap); }
    if (tid < 32) bs[tid] = __ldg(b1 + tid);
    for (long long w = blockIdx.x; w < n_eff; w += gridDim.x) {
        const long long win = idx ? (long long)idx[base + w] : (base + w);
        __syncthreads();
        for (int i = tid; i < (OLT_T + 6) * 8; i += 256) {
            const int r = (i >> 3) - 3, c = (i & 7) - 1;
            xin[i] = (r >= 0 && r < OLT_T && c >= 0 && c < OLT_F) ? win_fetch(src, win, r, c) : 0.f;
        }
        __syncthreads();
        float gs = 0.f, gq = 0.f;
        const float4 bias4 = *reinterpret_cast<const float4*>(bs + cb * 4);
        for (int h = tid >> 3; h < OLT_T; h += 32) {
            // two output channels per packed FMA (fma.rn.f32x2 -> FFMA2): 24 issue slots per filter row instead of 48
            f32x2 acc2[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) { acc2[q][0] = pk2(bias4.x, bias4.y); acc2[q][1] = pk2(bias4.z, bias4.w); }
#pragma unroll
            for (int dt = 0; dt < 7; ++dt) {
                const float4 xa = *reinterpret_cast<const float4*>(xin + (h + dt) * 8);
                const float2 xb = *reinterpret_cast<const float2*>(xin + (h + dt) * 8 + 4);
                const f32x2 xr[6] = {pk2(xa.x, xa.x), pk2(xa.y, xa.y), pk2(xa.z, xa.z), pk2(xa.w, xa.w), pk2(xb.x, xb.x), pk2(xb.y, xb.y)};
#pragma unroll
                for (int df = 0; df < 3; ++df) {
                    const ulonglong2 w4 = *reinterpret_cast<const ulonglong2*>(wk + (dt * 3 + df) * 32 + cb * 4);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        acc2[q][0] = fma2(xr[q + df], w4.x, acc2[q][0]);
                        acc2[q][1] = fma2(xr[q + df], w4.y, acc2[q][1]);
                    }
                }
            }
            float acc[4][4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { unpk2(acc2[q][0], acc[q][0], acc[q][1]); unpk2(acc2[q][1], acc[q][2], acc[q][3]); }
            float* o = out + ((size_t)w * OLT_T * OLT_F + (size_t)h * OLT_F) * 32 + cb * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                *reinterpret_cast<float4*>(o + q * 32) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
                gs += (acc[q][0] + acc[q][1]) + (acc[q][2] + acc[q][3]);
                gq += fmaf(acc[q][0], acc[q][0], fmaf(acc[q][1], acc[q][1], fmaf(acc[q][2], acc[q][2], acc[q][3] * acc[q][3])));
            }
        }
        // lanes with the same channel block (lane & 7) reduce, then the 8 warps through shared memory (fixed order)
        gs += __shfl_xor_sync(0xffffffffu, gs, 8); gq += __shfl_xor_sync(0xffffffffu, gq, 8);
        gs += __shfl_xor_sync(0xffffffffu, gs, 16); gq += __shfl_xor_sync(0xffffffffu, gq, 16);
        if (lane < 8) { red[warp][2 * lane] = gs; red[warp][2 * lane + 1] = gq; }
        __syncthreads();
        if (tid < 16) {
            double t = 0.0;
            for (int i = 0; i < 8; ++i) t += (double)red[i][tid];
            stats[(size_t)w * 16 + tid] = t;
        }
    }
}

// stats (sum, sumsq over 3200 elements per group) + affine -> per (window, channel) scale / shift:
// GroupNorm(x)[c] = x*sc + sh with sc = gamma[c]*rstd, sh = beta[c] - mean*sc (biased variance, eps inside the sqrt)
__global__ void ol_gn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      int C, float eps, const int* __restrict__ n_dev, long long n_total, long long base,
                                      long long n_chunk, float2* __restrict__ scsh) {
    const long long n_eff = eff_windows(n_total, n_dev, base, n_chunk);
    const int cpg = C / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_eff * C; i += (long long)gridDim.x * blockDim.x) {
        const long long w = i / C;
        const int c = (int)(i - w * C), g = c / cpg;
        const double s = stats[w * 16 + 2 * g], q = stats[w * 16 + 2 * g + 1];
        const double mean = s / 3200.0;
        double var = q / 3200.0 - mean * mean;
        if (var < 0.0) var = 0.0;
        const float rstd = 1.0f / sqrtf((float)var + eps);
        const float sc = __ldg(gamma + c) * rstd;
        scsh[w * 256 + c] = make_float2(sc, fmaf(-(float)mean, sc, __ldg(beta + c)));
    }
}

// ------------------------------------------------------------------------------------------------------------
// A-operand staging of block L: SiLU(GroupNorm(prev)) -> MaxPool(2,1) -> fp16 hi/lo -> K-major core-matrix images.
// One thread item = 8 channels of one (pair-local time step, window, source column).
// ------------------------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(256) ol_act_stage_kernel(const float* __restrict__ prev, const float2* __restrict__ scsh,
                                                           const int* __restrict__ n_dev, long long n_total, long long base,
                                                           long long n_chunk, unsigned char* __restrict__ staged) {
    using G = OlGeo<L>;
    using D = OlDer<L>;
    constexpr int PT = G::KT / 2;
    const long long n_eff = eff_windows(n_total, n_dev, base, n_chunk);
    const long long groups = (n_eff + D::WPT - 1) / D::WPT;
    const long long per_group = (long long)D::NPAIR * D::NCC * D::HH * D::WPT * 4 * 2;
    const long long total = groups * per_group;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i;
        const int c8 = (int)(r & 1); r >>= 1;
        const int ws = (int)(r & 3); r >>= 2;
        const int wi = (int)(r % D::WPT); r /= D::WPT;
        const int hh = (int)(r % D::HH); r /= D::HH;
        const int cc = (int)(r % D::NCC); r /= D::NCC;
        const int p = (int)(r % D::NPAIR);
        const long long g = r / D::NPAIR;
        const int h = p * D::PH + hh - PT;
        const long long win = g * D::WPT + wi;
        uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
        if (h >= 0 && h < G::HIN && win < n_eff) {
            const int ch = cc * 16 + c8 * 8;
            const float* pa = prev + (((size_t)win * (2 * G::HIN) + 2 * h) * 4 + ws) * G::CIN + ch;
            const float* pb = pa + (size_t)4 * G::CIN;
            const float2* ss = scsh + win * 256 + ch;
            float v[8];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 a = *reinterpret_cast<const float4*>(pa + 4 * q);
                const float4 b = *reinterpret_cast<const float4*>(pb + 4 * q);
                const float2 s0 = ss[4 * q], s1 = ss[4 * q + 1], s2 = ss[4 * q + 2], s3 = ss[4 * q + 3];
                v[4 * q + 0] = fmaxf(silu_acc(fmaf(a.x, s0.x, s0.y)), silu_acc(fmaf(b.x, s0.x, s0.y)));
                v[4 * q + 1] = fmaxf(silu_acc(fmaf(a.y, s1.x, s1.y)), silu_acc(fmaf(b.y, s1.x, s1.y)));
                v[4 * q + 2] = fmaxf(silu_acc(fmaf(a.z, s2.x, s2.y)), silu_acc(fmaf(b.z, s2.x, s2.y)));
                v[4 * q + 3] = fmaxf(silu_acc(fmaf(a.w, s3.x, s3.y)), silu_acc(fmaf(b.w, s3.x, s3.y)));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) split_f16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
        }
        // only the unshifted image goes to global memory: in the K-major core-matrix layout row r of a k-core block sits at
        // r*16 bytes, so the GEMM's producer realises the df = -1/+1 copies as the same bytes landed 16 B later / earlier
        unsigned char* img = staged + (((size_t)g * D::NPAIR + p) * D::NCC + cc) * D::A_GSTAGE + (size_t)c8 * (D::R * 16);
        const int row = hh * G::RPH + wi * 4 + ws;
        *reinterpret_cast<uint4*>(img + (size_t)row * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(img + D::A_IMG + (size_t)row * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// implicit-GEMM convolution of block L on the tensor cores
// ------------------------------------------------------------------------------------------------------------
struct OlGemmArgs {
    const unsigned char* staged;   // pre-staged operand images (separate staging kernel) ...
    const float* prev;             // ... or the previous block's raw output + its GroupNorm scale / shift (fused staging)
    const float2* scsh;
    const unsigned char* wimg;
    const float* bias;
    const float* wsc;          // [2]: weight scale, 1/scale
    float* out;                // channels-last [w][HIN][4][NOUT]
    double* stats;             // [w][8][2]
    const int* n_dev;
    long long n_total, base, n_chunk;
};

template <int L> struct OlBars {
    uint64_t a_landed[2], a_full[2], a_empty[2], b_full[OlDer<L>::NB], b_empty[OlDer<L>::NB], acc_full[2], acc_empty[2];
};
static_assert(sizeof(OlBars<1>) <= 240 && sizeof(OlBars<3>) <= 240, "barriers, then the TMEM base word at +240");

// NPW = 0: the A stages are bulk copies of images written by ol_act_stage_kernel (warps 4 + 7).
// NPW > 0: NPW extra warps (8 ..) build the stages themselves from the previous block's raw fp32 output -- GroupNorm affine,
// SiLU, MaxPool(2,1), fp16 hi/lo split, the three sensor-shifted copies -- so the activations make ONE trip through HBM per block
// (raw write by the epilogue, raw read here) instead of three (raw write, raw read + image write, image read).
template <int L, int NPW>
__global__ void __launch_bounds__(256 + 32 * NPW, 1) ol_conv_gemm_kernel(const OlGemmArgs a) {
    using G = OlGeo<L>;
    using D = OlDer<L>;
    constexpr int NOUT = G::NOUT, KT = G::KT, RPH = G::RPH, HIN = G::HIN;
    constexpr bool FUSED = NPW > 0;
    constexpr int NTAP = KT * 3;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* As = smem;
    unsigned char* Bs = smem + 2 * D::A_STAGE;
    OlBars<L>* bars = reinterpret_cast<OlBars<L>*>(smem + D::OFF_BAR);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + D::OFF_BAR + 240);
    float* bias_s = reinterpret_cast<float*>(smem + D::OFF_BIAS);
    double* red_s = reinterpret_cast<double*>(smem + D::OFF_RED);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_eff = eff_windows(a.n_total, a.n_dev, a.base, a.n_chunk);
    const long long groups = (n_eff + D::WPT - 1) / D::WPT;
    const long long items = groups * D::NPAIR;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->a_landed[i], 1); mbar_init(&bars->a_full[i], FUSED ? NPW : 1); mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < D::NB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], 4); }
        fence_mbar_init();
    }
    for (int i = tid; i < NOUT; i += 256 + 32 * NPW) bias_s[i] = __ldg(a.bias + i);
    if (warp == 6) tmem_alloc(tmem_holder, D::TCOLS);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_holder;

    if (FUSED && warp >= 8) {                          // ---- fused producer: A stages built from the raw activations
        // Thread = one (k-core block of 8 channels, window, sensor column) for every NPT/2-th staged row: its NIT items differ only
        // in the time step, so the 8 (scale, shift) pairs are loaded once per stage.  A warp's 32 lanes are 32 consecutive rows.
        constexpr int NPT = (NPW > 0 ? NPW : 4) * 32, RS = NPT / 2, NIT = (D::R + RS - 1) / RS, PT = KT / 2;
        constexpr uint32_t KB = D::R * 16;
        static_assert(RS % RPH == 0 && D::R % 32 == 0, "rows of one thread share (window, column); whole warps per k-core block");
        const int pt = tid - 256;
        const int kc = pt / RS, rsub = pt - kc * RS;
        const int wi = (rsub % RPH) >> 2, ws = rsub & 3;
        const uint32_t so0 = (uint32_t)kc * KB + (uint32_t)rsub * 16u;              // byte offset of (kc, row rsub) inside an image
        uint32_t it = 0;
        for (long long item = blockIdx.x; item < items; item += gridDim.x) {
            const long long g = item / D::NPAIR;
            const int p = (int)(item - g * D::NPAIR);
            const long long win = g * D::WPT + wi;
            const bool wok = win < n_eff;
            const int h0 = p * D::PH + rsub / RPH - PT;                            // time step of item j: h0 + j * (RS / RPH)
            const float* pa = a.prev + (((size_t)(wok ? win : 0) * (2 * HIN)) * 4 + ws) * G::CIN + kc * 8;
            const float4* sp0 = reinterpret_cast<const float4*>(a.scsh + (wok ? win : 0) * 256 + kc * 8);
            for (int cc = 0; cc < D::NCC; ++cc, ++it) {
                const uint32_t s = it & 1;
                float4 ra[NIT][4];                                 // two pooled time steps x 8 channels
                bool ok[NIT];
#pragma unroll
                for (int j = 0; j < NIT; ++j) {
                    const int h = h0 + j * (RS / RPH);
                    ok[j] = wok && (rsub + j * RS < D::R) && h >= 0 && h < HIN;
                    if (ok[j]) {
                        const float4* q = reinterpret_cast<const float4*>(pa + (size_t)(2 * h) * 4 * G::CIN + cc * 16);
                        ra[j][0] = __ldg(q); ra[j][1] = __ldg(q + 1);
                        ra[j][2] = __ldg(q + G::CIN); ra[j][3] = __ldg(q + G::CIN + 1);       // + 4 sensor columns = next time step
                    }
                }
                const float4 s01 = __ldg(sp0 + cc * 8), s23 = __ldg(sp0 + cc * 8 + 1), s45 = __ldg(sp0 + cc * 8 + 2), s67 = __ldg(sp0 + cc * 8 + 3);
                // the NEXT stage's activations are pulled from DRAM into L2 while this one is computed (no registers held)
                {
                    const bool same = cc + 1 < D::NCC;
                    const long long item2 = item + gridDim.x;
                    if (same || item2 < items) {
                        const long long g2 = item2 / D::NPAIR;
                        const int p2 = (int)(item2 - g2 * D::NPAIR);
                        const long long win2 = same ? win : g2 * D::WPT + wi;
                        const int h02 = same ? h0 : p2 * D::PH + rsub / RPH - PT;
                        const float* pn = a.prev + (((size_t)win2 * (2 * HIN)) * 4 + ws) * G::CIN + kc * 8 + (same ? (cc + 1) * 16 : 0);
#pragma unroll
                        for (int j = 0; j < NIT; ++j) {
                            const int h = h02 + j * (RS / RPH);
                            if (win2 < n_eff && (rsub + j * RS < D::R) && h >= 0 && h < HIN) {
                                prefetch_l2(pn + (size_t)(2 * h) * 4 * G::CIN); prefetch_l2(pn + (size_t)(2 * h + 1) * 4 * G::CIN);
                            }
                        }
                    }
                }
                // values first, slot second: the stage is computed into registers while the MMAs still read the slot it will go to
                uint4 vh[NIT], vl[NIT];
#pragma unroll
                for (int j = 0; j < NIT; ++j) {
                    uint32_t hi[4] = {0u, 0u, 0u, 0u}, lo[4] = {0u, 0u, 0u, 0u};
                    if (ok[j]) {
                        float v[8];
                        const float4 x0 = ra[j][0], x1 = ra[j][2], y0 = ra[j][1], y1 = ra[j][3];
                        v[0] = fmaxf(silu_fast(fmaf(x0.x, s01.x, s01.y)), silu_fast(fmaf(x1.x, s01.x, s01.y)));
                        v[1] = fmaxf(silu_fast(fmaf(x0.y, s01.z, s01.w)), silu_fast(fmaf(x1.y, s01.z, s01.w)));
                        v[2] = fmaxf(silu_fast(fmaf(x0.z, s23.x, s23.y)), silu_fast(fmaf(x1.z, s23.x, s23.y)));
                        v[3] = fmaxf(silu_fast(fmaf(x0.w, s23.z, s23.w)), silu_fast(fmaf(x1.w, s23.z, s23.w)));
                        v[4] = fmaxf(silu_fast(fmaf(y0.x, s45.x, s45.y)), silu_fast(fmaf(y1.x, s45.x, s45.y)));
                        v[5] = fmaxf(silu_fast(fmaf(y0.y, s45.z, s45.w)), silu_fast(fmaf(y1.y, s45.z, s45.w)));
                        v[6] = fmaxf(silu_fast(fmaf(y0.z, s67.x, s67.y)), silu_fast(fmaf(y1.z, s67.x, s67.y)));
                        v[7] = fmaxf(silu_fast(fmaf(y0.w, s67.z, s67.w)), silu_fast(fmaf(y1.w, s67.z, s67.w)));
#pragma unroll
                        for (int q = 0; q < 4; ++q) split_f16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
                    }
                    vh[j] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    vl[j] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                mbar_wait(&bars->a_empty[s], ((it >> 1) & 1) ^ 1);
                unsigned char* sbase = As + s * D::A_STAGE + so0;
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int j = 0; j < NIT; ++j) {
                    if (rsub + j * RS >= D::R) continue;
                    unsigned char* q1 = sbase + j * (RS * 16);
                    // df = 1: in[w] at its own row.  df = 0 holds in[w-1]: this value lands one row down, and the quad's last lane
                    // writes the zero of the w = 0 row instead; df = 2 (in[w+1]) mirrors it -- one store per image and thread.
                    *reinterpret_cast<uint4*>(q1 + 2 * D::A_IMG) = vh[j];
                    *reinterpret_cast<uint4*>(q1 + 3 * D::A_IMG) = vl[j];
                    unsigned char* q0 = q1 + (ws != 3 ? 16 : -48);
                    *reinterpret_cast<uint4*>(q0) = ws != 3 ? vh[j] : z;
                    *reinterpret_cast<uint4*>(q0 + D::A_IMG) = ws != 3 ? vl[j] : z;
                    unsigned char* q2 = q1 + 4 * D::A_IMG + (ws != 0 ? -16 : 48);
                    *reinterpret_cast<uint4*>(q2) = ws != 0 ? vh[j] : z;
                    *reinterpret_cast<uint4*>(q2 + D::A_IMG) = ws != 0 ? vl[j] : z;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full[s]);
            }
        }
    } else if (!FUSED && warp == 4) {                  // ---- producer: A stages (one per item x ci-chunk)
        if (lane == 0) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x) {
                const unsigned char* src = a.staged + (size_t)item * D::NCC * D::A_GSTAGE;
                constexpr uint32_t KB = D::R * 16;                         // one k-core block: R rows x 16 B
                for (int cc = 0; cc < D::NCC; ++cc, ++it) {
                    const uint32_t s = it & 1;
                    mbar_wait(&bars->a_empty[s], ((it >> 1) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->a_landed[s], 4u * (KB + 2u * (KB - 16u)));
                    const unsigned char* gimg = src + (size_t)cc * D::A_GSTAGE;
                    unsigned char* sbase = As + s * D::A_STAGE;
#pragma unroll
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int kc = 0; kc < 2; ++kc) {
                            const unsigned char* g0 = gimg + half * D::A_IMG + kc * KB;
                            // df = 0 holds in[w-1]: row r <- row r-1 (row 0 / every w = 0 row is zeroed by the fix-up warp)
                            bulk_g2s(sbase + (0 * 2 + half) * D::A_IMG + kc * KB + 16, g0, KB - 16, &bars->a_landed[s]);
                            bulk_g2s(sbase + (1 * 2 + half) * D::A_IMG + kc * KB, g0, KB, &bars->a_landed[s]);
                            // df = 2 holds in[w+1]: row r <- row r+1 (every w = 3 row zeroed by the fix-up warp)
                            bulk_g2s(sbase + (2 * 2 + half) * D::A_IMG + kc * KB, g0 + 16, KB - 16, &bars->a_landed[s]);
                        }
                }
            }
        }
    } else if (!FUSED && warp == 7) {                  // ---- fix-up: zero fill of the shifted copies' out-of-range column
        uint32_t it = 0;
        constexpr uint32_t KB = D::R * 16;
        for (long long item = blockIdx.x; item < items; item += gridDim.x) {
            for (int cc = 0; cc < D::NCC; ++cc, ++it) {
                const uint32_t s = it & 1;
                mbar_wait(&bars->a_landed[s], (it >> 1) & 1);
                unsigned char* sbase = As + s * D::A_STAGE;
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                for (int i = lane; i < (D::R / 4) * 4; i += 32) {            // (row group of 4) x (half, k-core)
                    const int r4 = i >> 2, hk = i & 3;
                    const uint32_t off = (uint32_t)(hk >> 1) * D::A_IMG + (uint32_t)(hk & 1) * KB;
                    *reinterpret_cast<uint4*>(sbase + 0 * D::A_IMG + off + (uint32_t)(r4 * 4) * 16) = z;                    // df = 0, w = 0
                    *reinterpret_cast<uint4*>(sbase + 4 * D::A_IMG + off + (uint32_t)(r4 * 4 + 3) * 16) = z;                // df = 2, w = 3
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->a_full[s]);
            }
        }
    } else if (warp == 5) {                            // ---- producer: weight stages (one per ci-chunk x tap)
        if (lane == 0) {
            uint32_t it = 0;
            for (long long item = blockIdx.x; item < items; item += gridDim.x) {
                for (int st = 0; st < D::NCC * NTAP / D::TG; ++st, ++it) {
                    const uint32_t s = it % D::NB;
                    mbar_wait(&bars->b_empty[s], ((it / D::NB) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bars->b_full[s], (uint32_t)D::B_STAGE);
                    bulk_g2s(Bs + s * D::B_STAGE, a.wimg + (size_t)st * D::B_STAGE, (uint32_t)D::B_STAGE, &bars->b_full[s]);
                }
            }
        }
    } else if (warp == 6) {                            // ---- MMA issuer (warp-uniform; one elected lane issues)
        constexpr uint32_t IDESC = make_idesc_f16(128, NOUT);
        const uint32_t as_a = smem_u32(As), bs_a = smem_u32(Bs);
        uint32_t a_it = 0, b_it = 0, n_item = 0;
        for (long long item = blockIdx.x; item < items; item += gridDim.x, ++n_item) {
            const uint32_t ab_set = n_item % D::NACC;                      // accumulator set of this pair
            mbar_wait(&bars->acc_empty[ab_set], ((n_item / D::NACC) & 1) ^ 1);   // epilogue drained the pair that used it last
            tc_fence_after_sync();
            for (int cc = 0; cc < D::NCC; ++cc, ++a_it) {
                const uint32_t sa = a_it & 1;
                mbar_wait(&bars->a_full[sa], (a_it >> 1) & 1);
                tc_fence_after_sync();
                for (int tg = 0; tg < NTAP / D::TG; ++tg, ++b_it) {
                    const uint32_t sb = b_it % D::NB;
                    mbar_wait(&bars->b_full[sb], (b_it / D::NB) & 1);
                    tc_fence_after_sync();
                    if (elect_one()) {
#pragma unroll
                        for (int tq = 0; tq < D::TG; ++tq) {
                            const int tap = tg * D::TG + tq;
                            const int dt = tap / 3, df = tap - dt * 3;
                            const uint32_t bb = bs_a + sb * D::B_STAGE + tq * D::B_TAP;
                            const uint64_t b_hi = make_smem_desc(bb, NOUT * 16, 128);
                            const uint64_t b_lo = make_smem_desc(bb + NOUT * 32, NOUT * 16, 128);
                            const uint32_t ab = as_a + sa * D::A_STAGE + (uint32_t)(df * 2) * D::A_IMG;
#pragma unroll
                            for (int tile = 0; tile < 2; ++tile) {
                                const uint32_t roff = (uint32_t)((dt + tile * D::HT) * RPH / 8) * 128u;
                                const uint64_t a_hi = make_smem_desc(ab + roff, D::R * 16, 128);
                                const uint64_t a_lo = make_smem_desc(ab + D::A_IMG + roff, D::R * 16, 128);
                                const uint32_t acc = tbase + (uint32_t)((ab_set * 2 + tile) * NOUT);
                                mma_ss(acc, a_hi, b_hi, IDESC, (cc | tap) ? 1u : 0u);
                                mma_ss(acc, a_lo, b_hi, IDESC, 1u);
                                mma_ss(acc, a_hi, b_lo, IDESC, 1u);
                            }
                        }
                        mma_commit(&bars->b_empty[sb]);
                    }
                    __syncwarp();
                }
                if (elect_one()) mma_commit(&bars->a_empty[sa]);
                __syncwarp();
            }
            if (elect_one()) mma_commit(&bars->acc_full[ab_set]);
            __syncwarp();
        }
    } else if (warp < 4) {                             // ---- epilogue: thread = accumulator row
        constexpr int GS = NOUT / 8;                   // channels per GroupNorm group (8, 16, 32)
        constexpr int GPC = 32 / GS;                   // groups per 32-column chunk (4, 2, 1)
        constexpr int WW = D::WPT < 8 ? D::WPT : 8;    // windows covered by one warp
        const int row = warp * 32 + lane;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        const float inv = __ldg(a.wsc + 1);
        const int wi = (row % RPH) >> 2, w = row & 3;
        uint32_t n_item = 0;
        for (long long item = blockIdx.x; item < items; item += gridDim.x, ++n_item) {
            const long long g = item / D::NPAIR;
            const int p = (int)(item - g * D::NPAIR);
            const uint32_t ab_set = n_item % D::NACC;
            const long long win = g * D::WPT + wi;
            float gs[8], gq[8];                        // this row's GroupNorm partials over both tiles (same window)
#pragma unroll
            for (int i = 0; i < 8; ++i) { gs[i] = 0.f; gq[i] = 0.f; }
            mbar_wait(&bars->acc_full[ab_set], (n_item / D::NACC) & 1);
            tc_fence_after_sync();
            // the accumulator is drained in 32-column chunks with the NEXT chunk's TMEM load in flight under the current chunk's
            // stores and statistics (two register buffers): with one accumulator set (block 4) the MMA issuer waits for this drain
            constexpr int NCK = NOUT / 32;
            static_assert(NCK % 2 == 0, "chunk 0 of the second tile lands in buffer A again");
            const uint32_t acc0 = tbase + lane_base + (uint32_t)(ab_set * 2 * NOUT);
            uint32_t va[32], vb[32];
            tmem_ld32(acc0, va);
#pragma unroll 1
            for (int tile = 0; tile < 2; ++tile) {
                const int h = p * D::PH + tile * D::HT + row / RPH;
                const bool valid = h < HIN && win < n_eff;
                float* orow = a.out + (((size_t)(valid ? win : 0) * HIN + (valid ? h : 0)) * 4 + w) * NOUT;
#pragma unroll
                for (int ci = 0; ci < NCK; ++ci) {
                    const int c0 = ci * 32;
                    tmem_ld_wait();
                    uint32_t (&v)[32] = (ci & 1) ? vb : va;
                    uint32_t (&vn)[32] = (ci & 1) ? va : vb;
                    if (ci + 1 < NCK) tmem_ld32(acc0 + (uint32_t)(tile * NOUT + c0 + 32), vn);
                    else if (tile == 0) tmem_ld32(acc0 + (uint32_t)NOUT, vn);
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = fmaf(__uint_as_float(v[j]), inv, bias_s[c0 + j]);
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(orow + c0 + 4 * q) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
#pragma unroll
                        for (int gg = 0; gg < GPC; ++gg) {
                            float s1 = 0.f, q2 = 0.f;
#pragma unroll
                            for (int j = 0; j < GS && j < 32; ++j) { const float x = f[gg * GS + j]; s1 += x; q2 = fmaf(x, x, q2); }
                            gs[c0 / GS + gg] += s1;
                            gq[c0 / GS + gg] += q2;
                        }
                    }
                }
            }
            // accumulators are in registers: hand the TMEM set back before the reduction
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->acc_empty[ab_set]);
            // GroupNorm statistics: reduce over the 4 sensor columns (and the 2 time steps a warp holds when RPH = 16),
            // then across the warps through shared memory in a fixed order; one fp64 atomic per (window, group, stat)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                gs[i] += __shfl_xor_sync(0xffffffffu, gs[i], 1); gq[i] += __shfl_xor_sync(0xffffffffu, gq[i], 1);
                gs[i] += __shfl_xor_sync(0xffffffffu, gs[i], 2); gq[i] += __shfl_xor_sync(0xffffffffu, gq[i], 2);
                if (RPH == 16) { gs[i] += __shfl_xor_sync(0xffffffffu, gs[i], 16); gq[i] += __shfl_xor_sync(0xffffffffu, gq[i], 16); }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");            // previous item's partials have been consumed
            if (w == 0 && (RPH != 16 || lane < 16)) {
                double* dst = red_s + ((size_t)warp * 8 + (wi % WW)) * 16;
#pragma unroll
                for (int i = 0; i < 8; ++i) { dst[2 * i] = (double)gs[i]; dst[2 * i + 1] = (double)gq[i]; }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = row; i < D::WPT * 16; i += 128) {
                const int wq = i >> 4, st = i & 15;                    // window of the group, (group, stat) slot
                const long long wn = g * D::WPT + wq;
                if (wn < n_eff) {
                    double tot;
                    if (D::WPT <= 8) tot = (red_s[(0 * 8 + wq) * 16 + st] + red_s[(1 * 8 + wq) * 16 + st]) +
                                           (red_s[(2 * 8 + wq) * 16 + st] + red_s[(3 * 8 + wq) * 16 + st]);
                    else tot = red_s[(((wq >> 3) + 0) * 8 + (wq & 7)) * 16 + st] + red_s[(((wq >> 3) + 2) * 8 + (wq & 7)) * 16 + st];
                    atomicAdd(a.stats + (size_t)wn * 16 + st, tot);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 6) tmem_dealloc(tbase, D::TCOLS);
}

// ------------------------------------------------------------------------------------------------------------
// tail: GroupNorm + SiLU + global average pool (25x4) -> Linear 256->128 + SiLU -> Linear 128->2 -> softmax[:,1] (fp64 out)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ol_head_kernel(const float* __restrict__ raw4, const float2* __restrict__ scsh,
                                                      const float* __restrict__ fc1t, const float* __restrict__ fc1b,
                                                      const float* __restrict__ fc2w, const float* __restrict__ fc2b,
                                                      const int* __restrict__ n_dev, long long n_total, long long base,
                                                      long long n_chunk, float* __restrict__ logits, double* __restrict__ prob) {
    __shared__ float s_gap[256], s_h[128];
    const int tid = threadIdx.x;
    const long long n_eff = eff_windows(n_total, n_dev, base, n_chunk);
    for (long long w = blockIdx.x; w < n_eff; w += gridDim.x) {
        __syncthreads();
        {
            const float2 ss = scsh[w * 256 + tid];
            const float* p = raw4 + (size_t)w * 25600 + tid;
            float s = 0.f;
#pragma unroll 10
            for (int i = 0; i < 100; ++i) s += silu_fast(fmaf(p[(size_t)i * 256], ss.x, ss.y));
            s_gap[tid] = s / 100.f;
        }
        __syncthreads();
        if (tid < 128) {
            float y = __ldg(fc1b + tid);
            for (int k = 0; k < 256; ++k) y = fmaf(s_gap[k], __ldg(fc1t + k * 128 + tid), y);
            s_h[tid] = silu_acc(y);
        }
        __syncthreads();
        if (tid < 32) {
            float l0 = 0.f, l1 = 0.f;
            for (int k = tid; k < 128; k += 32) {
                l0 = fmaf(s_h[k], __ldg(fc2w + k), l0);
                l1 = fmaf(s_h[k], __ldg(fc2w + 128 + k), l1);
            }
            l0 = warp_sum(l0) + __ldg(fc2b);
            l1 = warp_sum(l1) + __ldg(fc2b + 1);
            if (tid == 0) {
                logits[(base + w) * 2] = l0;
                logits[(base + w) * 2 + 1] = l1;
                if (prob) {
                    const float m = fmaxf(l0, l1);
                    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
                    prob[base + w] = (double)(e1 / (e0 + e1));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// weight images: conv.weight [NOUT][CIN][KT][3] fp32 -> per (ci-chunk, dt, df): {hi, lo} x [NOUT x 16] fp16, K-major
// core matrices, pre-multiplied by a power-of-two scale that lifts the lo halves out of the fp16 subnormal range
// ------------------------------------------------------------------------------------------------------------
__global__ void ol_tc_scale_kernel(const float* __restrict__ w, int n, float* __restrict__ wsc) {
    __shared__ float red[32];
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, red[i]);
        int e = 0;
        if (m > 0.f && isfinite(m)) e = (int)floorf(log2f(8192.f / m));          // |w * 2^e| <= 8192 < fp16 max
        e = e < 0 ? 0 : (e > 12 ? 12 : e);
        wsc[0] = exp2f((float)e);
        wsc[1] = exp2f(-(float)e);
    }
}

__global__ void ol_tc_pack_kernel(const float* __restrict__ w, int NOUT, int CIN, int KT, const float* __restrict__ wsc,
                                  unsigned short* __restrict__ out) {
    const float scale = wsc[0];
    const int total = NOUT * CIN * KT * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int df = r % 3; r /= 3;
        const int dt = r % KT; r /= KT;
        const int ci = r % CIN;
        const int n = r / CIN;
        const int cc = ci >> 4, c = ci & 15;
        const float v = w[i] * scale;
        const __half hv = __float2half_rn(v);
        const __half lv = __float2half_rn(v - __half2float(hv));
        const size_t stage = ((size_t)(cc * KT + dt) * 3 + df) * (size_t)(2 * NOUT * 16);
        const size_t off = (size_t)(c >> 3) * (NOUT * 8) + (size_t)(n >> 3) * 64 + (n & 7) * 8 + (c & 7);
        out[stage + off] = *reinterpret_cast<const unsigned short*>(&hv);
        out[stage + (size_t)NOUT * 16 + off] = *reinterpret_cast<const unsigned short*>(&lv);
    }
}

constexpr int kOlCin[4] = {1, 32, 64, 128};
constexpr int kOlCout[4] = {32, 64, 128, 256};
constexpr int kOlKt[4] = {7, 5, 5, 3};
constexpr size_t kStagedPerWindow = 86016;       // max over blocks of (NPAIR*NCC*A_GSTAGE / WPT)
static_assert((size_t)OlDer<1>::NPAIR * OlDer<1>::NCC * OlDer<1>::A_GSTAGE / OlDer<1>::WPT <= kStagedPerWindow, "staged size");
static_assert((size_t)OlDer<2>::NPAIR * OlDer<2>::NCC * OlDer<2>::A_GSTAGE / OlDer<2>::WPT <= kStagedPerWindow, "staged size");
static_assert((size_t)OlDer<3>::NPAIR * OlDer<3>::NCC * OlDer<3>::A_GSTAGE / OlDer<3>::WPT <= kStagedPerWindow, "staged size");

template <int L, int NPW>
int launch_gemm(CnnOlTc* t, const OlGemmArgs& a, long long n_chunk, cudaStream_t st) {
    using D = OlDer<L>;
    SHM_CUDA(cudaFuncSetAttribute(ol_conv_gemm_kernel<L, NPW>, cudaFuncAttributeMaxDynamicSharedMemorySize, D::SMEM));
    const long long max_items = (n_chunk + D::WPT - 1) / D::WPT * D::NPAIR;
    const int grid = (int)(max_items < t->nsm ? max_items : t->nsm);
    ol_conv_gemm_kernel<L, NPW><<<grid, 256 + 32 * NPW, D::SMEM, st>>>(a);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

constexpr int kOlProducerWarps = 8;

// SHM_OL_SEPARATE_STAGING=1 selects the round-1 structure (ol_act_stage_kernel writes the operand images to HBM, the GEMM's producer
// bulk-copies them): kept as the A/B reference of the fused producer (4.58 vs 3.90 ms per 8192 windows), not a fallback
bool ol_separate_staging() {
    static const bool separate = [] { const char* e = getenv("SHM_OL_SEPARATE_STAGING"); return e && e[0] == '1'; }();
    return separate;
}

template <int L>
int run_block(CnnOlTc* t, const float* prev, float* out, const float* bias, const int* n_dev, long long n_total, long long base,
              long long n_chunk, cudaStream_t st) {
    const bool separate = ol_separate_staging();
    OlGemmArgs a{t->staged, prev, t->scsh, t->wimg[L - 1], bias, t->wscale + 2 * (L - 1), out, t->stats + (size_t)L * t->chunk * 16,
                 n_dev, n_total, base, n_chunk};
    if (separate) {
        ol_act_stage_kernel<L><<<t->nsm * 8, 256, 0, st>>>(prev, t->scsh, n_dev, n_total, base, n_chunk, t->staged);
        SHM_LAUNCH_CHECK();
        return launch_gemm<L, 0>(t, a, n_chunk, st);
    }
    return launch_gemm<L, kOlProducerWarps>(t, a, n_chunk, st);
}

}  // namespace

int cnnol_tc_init(CnnOlTc* t, int device) {
    memset(t, 0, sizeof(*t));
    t->nsm = device_sm_count(device);
    for (int b = 1; b < 4; ++b) {
        const size_t bytes = (size_t)kOlCout[b] * kOlCin[b] * kOlKt[b] * 3 * 2 * sizeof(unsigned short);
        if (cudaMalloc(&t->wimg[b - 1], bytes) != cudaSuccess) { cudaGetLastError(); return SHM_ERR_NOMEM; }
    }
    if (cudaMalloc(&t->wscale, 6 * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return SHM_ERR_NOMEM; }
    return SHM_OK;
}

void cnnol_tc_free(CnnOlTc* t) {
    for (int b = 0; b < 3; ++b) if (t->wimg[b]) cudaFree(t->wimg[b]);
    if (t->wscale) cudaFree(t->wscale);
    if (t->raw[0]) cudaFree(t->raw[0]);
    if (t->raw[1]) cudaFree(t->raw[1]);
    if (t->staged) cudaFree(t->staged);
    if (t->stats) cudaFree(t->stats);
    if (t->scsh) cudaFree(t->scsh);
    memset(t, 0, sizeof(*t));
}

int cnnol_tc_pack(CnnOlTc* t, const float* const raw_w[4], cudaStream_t st) {
    for (int b = 1; b < 4; ++b) {
        const int total = kOlCout[b] * kOlCin[b] * kOlKt[b] * 3;
        ol_tc_scale_kernel<<<1, 1024, 0, st>>>(raw_w[b], total, t->wscale + 2 * (b - 1));
        SHM_LAUNCH_CHECK();
        ol_tc_pack_kernel<<<148, 256, 0, st>>>(raw_w[b], kOlCout[b], kOlCin[b], kOlKt[b], t->wscale + 2 * (b - 1),
                                               reinterpret_cast<unsigned short*>(t->wimg[b - 1]));
        SHM_LAUNCH_CHECK();
    }
    return SHM_OK;
}

static int ensure_workspace(CnnOlTc* t, long long n) {
    // 9472 = 148 * 64: with 7 tile pairs per group of 4 / 8 / 16 windows the three GEMMs get 112 / 56 / 28 whole waves of items on
    // 148 SMs (8192 left block 4 a 25th, 20 %-filled wave)
    constexpr long long kMaxChunk = 9472;
    long long want = n < kMaxChunk ? n : kMaxChunk;
    want = (want + 255) / 256 * 256;
    if (t->chunk >= want) return SHM_OK;
    if (t->raw[0]) { cudaFree(t->raw[0]); cudaFree(t->raw[1]); if (t->staged) cudaFree(t->staged); cudaFree(t->stats); cudaFree(t->scsh); }
    t->raw[0] = t->raw[1] = nullptr; t->staged = nullptr; t->stats = nullptr; t->scsh = nullptr; t->chunk = 0;
    const size_t raw_bytes = (size_t)want * 25600 * sizeof(float);
    if (cudaMalloc(&t->raw[0], raw_bytes) != cudaSuccess || cudaMalloc(&t->raw[1], raw_bytes) != cudaSuccess ||
        (ol_separate_staging() && cudaMalloc(&t->staged, (size_t)want * kStagedPerWindow) != cudaSuccess) ||
        cudaMalloc(&t->stats, (size_t)4 * want * 16 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&t->scsh, (size_t)want * 256 * sizeof(float2)) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaMalloc(cnnol tensor-core workspace)");
        return SHM_ERR_NOMEM;
    }
    t->chunk = (int)want;
    return SHM_OK;
}

int cnnol_tc_forward(CnnOlTc* t, const float* const conv1_w, const float* const bias[4], const float* const gn_w[4],
                     const float* const gn_b[4], const float* fc1t, const float* fc1b, const float* fc2w, const float* fc2b,
                     float gn_eps, const WinSrc& src, const int* idx, const int* n_dev, long long n, float* logits, double* prob,
                     cudaStream_t st) {
    int rc = ensure_workspace(t, n);
    if (rc != SHM_OK) return rc;
    const long long CH = t->chunk;
    for (long long base = 0; base < n; base += CH) {
        const long long nc = (n - base) < CH ? (n - base) : CH;
        SHM_CUDA(cudaMemsetAsync(t->stats, 0, (size_t)4 * CH * 16 * sizeof(double), st));
        const int grid_w = (int)(nc < (long long)t->nsm * 8 ? nc : (long long)t->nsm * 8);
        ol_conv1_kernel<<<grid_w, 256, 0, st>>>(conv1_w, bias[0], src, idx, n_dev, n, base, nc, t->raw[0], t->stats);
        SHM_LAUNCH_CHECK();
        const float* prev = t->raw[0];
        float* cur = t->raw[1];
        for (int b = 1; b < 4; ++b) {
            // GroupNorm affine of block b-1's output, folded into block b's operand staging
            ol_gn_finalize_kernel<<<t->nsm * 4, 256, 0, st>>>(t->stats + (size_t)(b - 1) * CH * 16, gn_w[b - 1], gn_b[b - 1], kOlCout[b - 1],
                                                              gn_eps, n_dev, n, base, nc, t->scsh);
            SHM_LAUNCH_CHECK();
            if (b == 1) rc = run_block<1>(t, prev, cur, bias[1], n_dev, n, base, nc, st);
            else if (b == 2) rc = run_block<2>(t, prev, cur, bias[2], n_dev, n, base, nc, st);
            else rc = run_block<3>(t, prev, cur, bias[3], n_dev, n, base, nc, st);
            if (rc != SHM_OK) return rc;
            const float* tmp = prev; prev = cur; cur = const_cast<float*>(tmp);
        }
        ol_gn_finalize_kernel<<<t->nsm * 4, 256, 0, st>>>(t->stats + (size_t)3 * CH * 16, gn_w[3], gn_b[3], 256, gn_eps, n_dev, n, base,
                                                          nc, t->scsh);
        SHM_LAUNCH_CHECK();
        ol_head_kernel<<<grid_w, 256, 0, st>>>(prev, t->scsh, fc1t, fc1b, fc2w, fc2b, n_dev, n, base, nc, logits, prob);
        SHM_LAUNCH_CHECK();
    }
    return SHM_OK;
}

}  // namespace shm
