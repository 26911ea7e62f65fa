// Tensor-core engine (SHM_ENGINE_TC_BF16X3) of the fused LSTM-VAE scorer.
//
// Same math and same fusion as vae_fp32.cuh (TemporalVAE.forward + per-window MSE,
// 4DOF/Scripts/Models/temporal_vae.py:51-77, 04_vae_thresholding.py:113-124), but every gate
// pre-activation GEMM runs on the 5th-gen tensor cores:
//
//   * tcgen05.mma.kind::f16, M = 128 windows x N = 128 (32 hidden units x 4 gates) x K = 16, fp32
//     accumulators in TMEM.  fp32-grade accuracy (the 1e-4 score tolerance rules out plain bf16/tf32,
//     SURVEY.md section 7) comes from a 3-pass split: x = hi + lo, D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo,
//     with fp16 halves (2^-22 operand error) for the bounded operands -- hidden states, decoder input,
//     weights -- and bf16 halves (2^-17) only for the raw window values, whose range is unbounded.
//   * the recurrent operand h_{t-1} never leaves the SM: the epilogue writes h_t (hi|lo) straight back
//     into TMEM with tcgen05.st and the next step's MMAs read A from TMEM (the ".ts" form).
//   * the layer input (x_t, the lower layer's h_t stream, or the constant decoder input u) is an
//     A operand in shared memory (K-major, no-swizzle core-matrix image).
//   * weights (pre-scaled by -log2(e) / -2log2(e) so the accumulator is directly the ex2 argument)
//     stream L2 -> smem through a 5-stage ring of 1-D bulk async copies (TMA engine) gated by mbarriers;
//     the two CTAs of a cluster walk the same stream in lock step, each loads half of every stage and
//     multicasts it into both rings (TC_CLUSTER).
//   * warp-specialised persistent CTA (one per SM): 16 epilogue warps (LSTM cell in registers,
//     ex2/rcp with merged divisions: 7 MUFU per cell), 4 window-staging / output warps, 1 copy-producer
//     warp, 1 MMA-issuer warp.  Two 128-column accumulator buffers let chunk c+1's MMAs overlap chunk c's
//     cell update; the input-projection MMAs of step t+1 overlap the tail of step t.
//   * layers of a stack run one after the other over all T steps; the lower layer's h_t stream goes
//     through a per-CTA global scratch (written once, read once by bulk copies).
//   * the decoder's first layer sees the same input at every step: its input projection is computed once
//     per tile and added by the cell update instead of the bias (IN_HOIST); a re-score call (io.mu_in)
//     starts at the heads stage with the encoder outputs of an earlier call.
//   * the single-layer H = 64 model runs two tiles per CTA with resident weights: vae_tc_dual.cuh.
#include "tcgen05.cuh"
#include "vae_tc.cuh"

namespace shm {
using namespace tc;

constexpr int TCM = 128;                       // windows per tile (UMMA M)
constexpr int TC_NST = 5;                      // weight ring stages
constexpr int TC_STAGE = 128 * 64 * 2;         // bytes per stage: [128 N-rows x 64 K] bf16
constexpr int TC_XSTAGE = 128 * 16 * 2;        // [128 x 16] bf16 (window-input tiles)
constexpr int TC_EPI_WARPS = 16;              // 4 warps per SM sub-partition: the cell update hides MUFU / TMEM latency across warps
constexpr int TC_UPT = 32 / (TC_EPI_WARPS / 4);   // hidden units per epilogue thread per 32-unit chunk (8)
constexpr int TC_WARP_AUX0 = TC_EPI_WARPS;    // 4 warps: window staging (encoder input) / output Linear + error (decoder)
constexpr int TC_WARP_PROD = TC_EPI_WARPS + 4, TC_WARP_MMA = TC_EPI_WARPS + 5;
constexpr int TC_THREADS = (TC_EPI_WARPS + 8) * 32;   // epilogue | staging/output group | copy producer | MMA issuer | 2 idle
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_CLUSTER = 2;                 // CTAs per cluster: each loads half of every weight stage and multicasts it to both
constexpr int TC_HEADS_SCRATCH = 3 * 16 * 128 * 4;   // per-CTA global scratch of the heads stage (mu | logvar | z), bytes
// setmaxnreg budgets: 16*32*88 + 8*32*64 = 61440 of 65536 registers.  Do not raise the epilogue's share: 96 / 56 (63,488 in total, one
// spill pair less in the epilogue loops) DEAD-LOCKED the kernel on the B200 -- like the budgets that added up to exactly 65,536 before it.
constexpr int TC_REG_EPI = 88, TC_REG_AUX = 64;
constexpr float NLOG2E = -1.4426950408889634f;

enum { IN_X = 0, IN_STREAM = 1, IN_CONST = 2, IN_HOIST = 3 };
// IN_HOIST: the layer input is constant over time (decoder layer 0: u = tanh(W z + b) repeated T times, temporal_vae.py:65-70), so its
// projection G = u W_ih^T + b is computed ONCE per tile (one MMA sweep, drained to a per-CTA global image) and each step's cell update
// adds its chunk of G, streamed back through the two idle input buffers, instead of the bias: the pass issues half the MMAs.
enum { SINK_STREAM = 0, SINK_LAST_ENC = 1, SINK_LAST_DEC = 2 };

struct TcPassDev {
    const unsigned char* w;     // per chunk: [in part][hh part]; part = [kt][hi|lo] stage images
    const float* bias;          // [H][4] pre-scaled b_ih + b_hh
    int in_kind, sink;
};
struct TcDev {
    TcPassDev pass[2 * SHM_MAX_L];
    int n_pass, L;
    unsigned char* scratch;     // per-CTA h_t stream: [grid][T][hi|lo image]
    unsigned long long scratch_stride;
    unsigned long long g_bytes; // per-CTA image of the hoisted input projection (IN_HOIST), 0 if unused
    long long* dbg;             // optional [grid][8] profiling counters
};

template <int H>
struct TcSmem {
    static constexpr int NCH = H / 32;
    static constexpr int IMGH = TCM * H * 2;                // one bf16 image [128 x H]
    static constexpr int IMG = 2 * IMGH;                    // hi + lo
    static constexpr int off_ring = 0;
    static constexpr int off_in = off_ring + TC_NST * TC_STAGE;
    static constexpr int WOIMG = 16 * H * 2;                // W_o as one fp16 K-major image [16 x H]
    static constexpr int off_wo = off_in + 2 * IMG;         // hi image, lo image, then b_o[16] fp32
    static constexpr int off_bias = off_wo + 2 * WOIMG + 64;
    static constexpr int off_bar = off_bias + H * 4 * 4;
    static constexpr int total = off_bar + 32 * 8 + 16;
    static_assert(total <= 232448, "shared memory budget");
};

struct TcBars {
    uint64_t w_full[TC_NST], w_empty[TC_NST], in_full[2], in_empty[2], acc_full[2], acc_empty[2], h_full[2], xhat_full, xhat_empty;
    uint64_t g_full[2], g_empty[2];     // IN_HOIST: chunk images of G in the two input buffers
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void aux_bar_sync() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// LSTM cell from pre-scaled gate arguments (ai = -log2e*a_i, af, ao likewise, ag = -2log2e*a_g):
// sigma(a) = 1/(1+2^ai), tanh(a) = (1-2^ag)/(1+2^ag); the three divisions of c' = f*c + i*g share one
// reciprocal, the two of h = o*tanh(c') another: 5 ex2 + 2 rcp per cell.
//   e_x = 2^min(a_x, clamp);  a = 1+e_i, b = 1+e_f, d = 1+e_g, n = 2-d;  P = a*d;  c' = (c*P + n*b) / (P*b);
//   q = 1 + 2^min(2*NLOG2E*c', 30);  h = (2-q) / ((1+e_o)*q)
// Where a pass adds the same bias for every window (all but the hoisted decoder pass) the pre-scaled bias is folded into a per-gate
// MULTIPLIER B = 2^bias: 2^(a + bias) = 2^a * B, so `1 + e` becomes one FMA and the four bias additions disappear.  Arguments are
// then clamped at 22 and multipliers at 2^20 (by the caller): 1 + e*B <= 2^42, so P*b <= 2^126 stays finite when three gates
// saturate at once; the saturation floor 2^-22 is below what the 1e-4 score tolerance can see.  The two-tile H = 64 scorer is bound
// by issue slots (74 % busy next to XU 70 %); moving a reciprocal onto the FMA pipe instead made it slower (76.8 vs 74.4 ms per 2^20
// windows), the multiplier gained 3 % and the packed pair update below another 3 %.
//
// Two cells at once on the packed fp32 pipe (sm_100 fma/mul/add.f32x2 -> FFMA2: two lanes per issue slot, same flops per clock as
// two FFMAs -- scripts/ffma2_probe.cu): the cell update's 16 FMA-pipe instructions per cell become 15 per PAIR, which frees ~8 of
// ~35 issue slots per cell next to the 7 MUFU operations that bound it.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 ex2_2(float x0, float x1, float clampv) {
    return pk2(ex2_approx(fminf(x0, clampv)), ex2_approx(fminf(x1, clampv)));
}
__device__ __forceinline__ f32x2 rcp_2(f32x2 v) { float a, b; unpk2(v, a, b); return pk2(rcp_approx(a), rcp_approx(b)); }
// 1/x0, 1/x1 from ONE reciprocal: r = rcp(x0*x1), 1/x0 = r*x1, 1/x1 = r*x0 (1 MUFU + 3 FMUL instead of 2 MUFU).  Only for the
// h = (2-q) / ((1+e_o) q) denominators: they are >= 1 (no underflow) and <= 2^73, so the product overflows (r = 0, both h = 0) only
// when each exceeds 2^55, i.e. 1+e_o >= 2^25 and the true |h| < 2^-25 for both cells.  Not for c' = num / (P b): there a saturated
// input + cell gate with an OPEN forget gate makes P b huge while c' = c is not small.
#ifndef SHM_PAIR_RCP_H
#define SHM_PAIR_RCP_H 0
#endif
__device__ __forceinline__ f32x2 rcp_pair(f32x2 v) {
#if SHM_PAIR_RCP_H
    float a, b;
    unpk2(v, a, b);
    const float r = rcp_approx(a * b);
    return pk2(r * b, r * a);
#else
    return rcp_2(v);
#endif
}
// gate arguments of units (u, u+1); B4 = multipliers {Bi(u),Bi(u+1)}, {Bf..}, {Bg..}, {Bo..}; c, h in/out for both units
__device__ __forceinline__ void lstm_cell_bmul2(float ai0, float ai1, float af0, float af1, float ag0, float ag1, float ao0, float ao1,
                                                const f32x2 (&B4)[4], float& c0, float& c1, float& h0, float& h1) {
    const f32x2 one = pk2(1.f, 1.f), two = pk2(2.f, 2.f), m1 = pk2(-1.f, -1.f), k2 = pk2(2.f * NLOG2E, 2.f * NLOG2E);
    const f32x2 ei = ex2_2(ai0, ai1, 22.f), ef = ex2_2(af0, af1, 22.f), eg = ex2_2(ag0, ag1, 22.f), eo = ex2_2(ao0, ao1, 22.f);
    const f32x2 a = fma2(ei, B4[0], one), b = fma2(ef, B4[1], one), d = fma2(eg, B4[2], one);
    const f32x2 n = fma2(d, m1, two);
    const f32x2 P = mul2(a, d);
    const f32x2 num = fma2(pk2(c0, c1), P, mul2(n, b));
    const f32x2 c = mul2(num, rcp_2(mul2(P, b)));
    unpk2(c, c0, c1);
    float k0, k1;
    unpk2(mul2(c, k2), k0, k1);
    const f32x2 q = add2(ex2_2(k0, k1, 30.f), one);
    const f32x2 h = mul2(fma2(q, m1, two), rcp_pair(mul2(fma2(eo, B4[3], one), q)));
    unpk2(h, h0, h1);
}

// The same pair update from complete gate arguments (bias already inside, as in the hoisted decoder pass): lstm_cell's formulas.
__device__ __forceinline__ void lstm_cell2(f32x2 ai, f32x2 af, f32x2 ag, f32x2 ao, float& c0, float& c1, float& h0, float& h1) {
    const f32x2 one = pk2(1.f, 1.f), two = pk2(2.f, 2.f), m1 = pk2(-1.f, -1.f), k2 = pk2(2.f * NLOG2E, 2.f * NLOG2E);
    float x0, x1;
    unpk2(ai, x0, x1); const f32x2 a = add2(ex2_2(x0, x1, 30.f), one);
    unpk2(af, x0, x1); const f32x2 b = add2(ex2_2(x0, x1, 30.f), one);
    unpk2(ag, x0, x1); const f32x2 d = add2(ex2_2(x0, x1, 30.f), one);
    unpk2(ao, x0, x1); const f32x2 o = add2(ex2_2(x0, x1, 30.f), one);
    const f32x2 n = fma2(d, m1, two);
    const f32x2 P = mul2(a, d);
    const f32x2 num = fma2(pk2(c0, c1), P, mul2(n, b));
    const f32x2 c = mul2(num, rcp_2(mul2(P, b)));
    unpk2(c, c0, c1);
    unpk2(mul2(c, k2), x0, x1);
    const f32x2 q = add2(ex2_2(x0, x1, 30.f), one);
    const f32x2 h = mul2(fma2(q, m1, two), rcp_pair(mul2(o, q)));
    unpk2(h, h0, h1);
}

// chunk loop of the epilogue passes: 1 = rolled (the cell state rotates through the register arrays: ~45 moves per 290-instruction
// chunk), 2 = two chunks per iteration (accumulator buffer c & 1 static, half the moves), 4 = fully unrolled.  Same-box A/B at the
// power cap, 2^20 windows: 1 -> 4.44 M, 2 -> 4.47 M, 4 -> 4.38 M windows/s (4 x 290 instructions x 4 pass bodies press on the
// 32 KB instruction cache again); scripts/ab_unroll.sh.
#ifndef TC_EPI_UNROLL
#define TC_EPI_UNROLL 2
#endif
constexpr int kTcEpiUnroll = TC_EPI_UNROLL;

// optional role profiling (TcDev.dbg != nullptr): cycles the MMA issuer spends in each kind of wait
#define TC_TWAIT(slot, bar, par)                                   \
    do {                                                           \
        const long long _t0 = TC_CLOCK();                           \
        mbar_wait(bar, par);                                       \
        prof[slot] += TC_CLOCK() - _t0;                             \
    } while (0)

// mbarrier use-counters: two buffers each; (b ? n1 : n0) keeps them in registers
struct Cnt2 {
    uint32_t n0, n1;
    __device__ __forceinline__ uint32_t get(int b) const { return b ? n1 : n0; }
    __device__ __forceinline__ void inc(int b) { if (b) ++n1; else ++n0; }
};

constexpr uint64_t TC_DESC_HI = ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
__device__ __forceinline__ uint64_t kdesc(uint32_t smem_addr) { return TC_DESC_HI | (uint64_t)((smem_addr & 0x3FFFFu) >> 4); }

struct EpiCtx {
    TcBars* bars;
    uint32_t t_acc, hbuf, lane_base;
    const float* bias_s;
    unsigned char* img;          // scratch image of this step (SINK_STREAM)
    float* hT;                   // fp32 h_T [H][128] (SINK_LAST_ENC)
    const unsigned char* gs;     // input buffers carrying G chunk images [2 planes][16 unit pairs][128 rows][4] fp32 (HOIST)
    int wg, row, lane;
    bool last_step, first_step;
    long long* prof;
};

// One chunk (32 hidden units x 4 gates) of one step for this thread's (row, 16-unit) slice, in two batches of
// 8 units to bound register pressure: TMEM -> registers, LSTM cell, h_t -> TMEM (fp16 hi|lo) and the pass's sink.
// `c` is a RUN-TIME chunk index: the chunk loop is a real loop (the cell state rotates through the register arrays),
// so the hot loop of the epilogue is ~11 KB of code instead of ~41 KB -- the fully unrolled form overflowed the 32 KB
// L1.5 instruction cache (ncu: 25 % of the epilogue's issue slots were stall_no_inst).
template <int H, int SINK, bool HOIST>
__device__ __forceinline__ void epi_chunk(const EpiCtx& x, int c, uint32_t acc_parity, uint32_t g_parity, float (&cst)[TC_UPT]) {
    using S = TcSmem<H>;
    static_assert(TC_UPT == 8, "one batch of 8 units per thread per chunk");
    const int b = c & 1;
    {
        const long long tw0 = TC_CLOCK();
        mbar_wait(&x.bars->acc_full[b], acc_parity);
        x.prof[0] += TC_CLOCK() - tw0;
    }
    tc_fence_after_sync();
    const int u0 = c * 32 + x.wg * 8;                                // first hidden unit of this thread's slice
    uint32_t g0[8], g1[8], g2[8], g3[8];
    const uint32_t abase = x.t_acc + x.lane_base + (uint32_t)(b * 128 + x.wg * 8);
    tmem_ld8(abase + 0, g0);
    tmem_ld8(abase + 32, g1);
    tmem_ld8(abase + 64, g2);
    tmem_ld8(abase + 96, g3);
    tmem_ld_wait();
    tc_fence_before_sync();                                          // accumulator slice in registers: release it
    __syncwarp();
    if (x.lane == 0) mbar_arrive(&x.bars->acc_empty[b]);
    float hv[8];
    if (HOIST) {
        if (x.first_step) {                                          // h_{-1} = 0: no MMA was issued, the accumulator is stale
#pragma unroll
            for (int u = 0; u < 8; ++u) { g0[u] = 0u; g1[u] = 0u; g2[u] = 0u; g3[u] = 0u; }
        }
        mbar_wait(&x.bars->g_full[b], g_parity);                    // this chunk's G image has landed in input buffer b
        // G = u W_ih^T + b per unit PAIR and row: plane 0 {i(u),i(u+1),f(u),f(u+1)}, plane 1 {g.., o..}, [plane][pair][row] 16 B each
        const ulonglong2* G = reinterpret_cast<const ulonglong2*>(x.gs + b * S::IMG) + (x.wg * 4) * TCM + x.row;
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            const ulonglong2 gif = G[(u >> 1) * TCM], ggo = G[(16 + (u >> 1)) * TCM];
            lstm_cell2(add2(pk2(__uint_as_float(g0[u]), __uint_as_float(g0[u + 1])), gif.x),
                       add2(pk2(__uint_as_float(g1[u]), __uint_as_float(g1[u + 1])), gif.y),
                       add2(pk2(__uint_as_float(g2[u]), __uint_as_float(g2[u + 1])), ggo.x),
                       add2(pk2(__uint_as_float(g3[u]), __uint_as_float(g3[u + 1])), ggo.y), cst[u], cst[u + 1], hv[u], hv[u + 1]);
        }
        __syncwarp();
        if (x.lane == 0) mbar_arrive(&x.bars->g_empty[b]);
    } else {
#pragma unroll
        for (int u = 0; u < 8; u += 2) {                                 // gate multipliers 2^bias, stored per unit pair
            const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(x.bias_s + (u0 + u) * 4);
            const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(x.bias_s + (u0 + u) * 4 + 4);
            const f32x2 B4[4] = {b01.x, b01.y, b23.x, b23.y};
            lstm_cell_bmul2(__uint_as_float(g0[u]), __uint_as_float(g0[u + 1]), __uint_as_float(g1[u]), __uint_as_float(g1[u + 1]),
                            __uint_as_float(g2[u]), __uint_as_float(g2[u + 1]), __uint_as_float(g3[u]), __uint_as_float(g3[u + 1]), B4,
                            cst[u], cst[u + 1], hv[u], hv[u + 1]);
        }
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_f16x2(hv[2 * j], hv[2 * j + 1], hi[j], lo[j]);
    tmem_st4(x.hbuf + x.lane_base + (uint32_t)(u0 >> 1), hi);
    tmem_st4(x.hbuf + x.lane_base + (uint32_t)(H / 2 + (u0 >> 1)), lo);
    if (SINK == SINK_STREAM) {
        const int off = ((u0 >> 3) * 16 + (x.row >> 3)) * 128 + (x.row & 7) * 16;
        *reinterpret_cast<uint4*>(x.img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(x.img + S::IMGH + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    } else if (SINK == SINK_LAST_ENC) {
        if (x.last_step) {
#pragma unroll
            for (int u = 0; u < 8; ++u) x.hT[(u0 + u) * TCM + x.row] = hv[u];
        }
    }
}

struct PassCtx {
    TcBars* bars;
    unsigned char* ring;
    unsigned char* inbuf;
    float* bias_s;
    unsigned char* wo_img;       // W_o fp16 hi|lo K-major images (B operand of the output Linear)
    float* bo_s;
    unsigned char* scratch;
    float* heads_scratch;
    unsigned char* gbuf;         // global image of G, [NCH][2 planes][16 unit pairs][128 rows][4] fp32 (IN_HOIST)
    uint32_t t_acc, t_h;
    int T, nvalid, hT_buf, u_buf;
    long long n0;
};

// ------------------------------------------------------------------------------------------------ epilogue warps
template <int H, int SINK, bool HOIST = false>
__device__ __forceinline__ void epi_pass(const PassCtx& pc, const VaeDev& P, const WinSrc& src, const VaeIO& io, const float* bias_g,
                                      Cnt2 acc_cnt, long long (&prof)[8], Cnt2 g_cnt = Cnt2{0, 0}) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = warp >> 2;                                  // units [16*wg, 16*wg+16) of each chunk
    const int row = (warp & 3) * 32 + lane;                    // TMEM lane == window row
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int T = pc.T;
    // non-hoisted passes fold the bias into the cell update as a multiplier 2^bias (lstm_cell_bmul); the hoisted pass reads b from G
    if (!HOIST)
        for (int i = tid; i < H * 4; i += TC_EPI_THREADS) {               // {Bi(u),Bi(u+1)} {Bf..} {Bg..} {Bo..} per unit pair (lstm_cell_bmul2)
            const int unit = i >> 2, gate = i & 3;
            pc.bias_s[(unit >> 1) * 8 + gate * 2 + (unit & 1)] = exp2f(fminf(__ldg(bias_g + i), 20.f));
        }
    epi_bar_sync();
    float cst[NCH][TC_UPT];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int u = 0; u < TC_UPT; ++u) cst[c][u] = 0.f;
    uint32_t nacc0 = acc_cnt.n0, nacc1 = acc_cnt.n1, ng0 = g_cnt.n0, ng1 = g_cnt.n1;
    for (int t = 0; t < T; ++t) {
        const uint32_t hbuf = pc.t_h + (uint32_t)((t & 1) * H);
        EpiCtx ctx{pc.bars, pc.t_acc, hbuf, lane_base, pc.bias_s, pc.scratch + (size_t)t * S::IMG,
                   reinterpret_cast<float*>(pc.inbuf + pc.hT_buf * S::IMG), pc.inbuf, wg, row, lane, t == T - 1, t == 0, prof};
#pragma unroll kTcEpiUnroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t par = ((c & 1) ? nacc1 : nacc0) & 1;
            if (c & 1) ++nacc1; else ++nacc0;
            const uint32_t gpar = ((c & 1) ? ng1 : ng0) & 1;
            if (HOIST) { if (c & 1) ++ng1; else ++ng0; }
            epi_chunk<H, SINK, HOIST>(ctx, c, par, gpar, cst[0]);
            if constexpr (NCH > 1) {                           // rotate the cell state: chunk c+1's state moves to cst[0]
#pragma unroll
                for (int u = 0; u < TC_UPT; ++u) {
                    const float t0 = cst[0][u];
#pragma unroll
                    for (int k = 0; k + 1 < NCH; ++k) cst[k][u] = cst[k + 1][u];
                    cst[NCH - 1][u] = t0;
                }
            }
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pc.bars->h_full[t & 1]);
    }
    if (SINK == SINK_STREAM) fence_proxy_async_all();          // scratch writes -> visible to the bulk-copy engine
}

// IN_HOIST, once per tile: drain G = u W_ih^T (+ b) chunk by chunk from the accumulators into the per-CTA global image
// [chunk][2 planes][16 unit pairs][128 rows][4] fp32 -- the layout the packed cell update reads back from shared memory (two 16-byte
// loads per row and unit pair: {i, i', f, f'} and {g, g', o, o'}).
template <int H>
__device__ __forceinline__ void epi_hoist_pre(const PassCtx& pc, const float* bias_g, Cnt2 acc_cnt) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = warp >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    for (int i = tid; i < H * 4; i += TC_EPI_THREADS) pc.bias_s[i] = __ldg(bias_g + i);
    epi_bar_sync();
    uint32_t nacc0 = acc_cnt.n0, nacc1 = acc_cnt.n1;
#pragma unroll 1
    for (int c = 0; c < NCH; ++c) {
        const int b = c & 1;
        mbar_wait(&pc.bars->acc_full[b], ((c & 1) ? nacc1 : nacc0) & 1);
        if (c & 1) ++nacc1; else ++nacc0;
        tc_fence_after_sync();
        uint32_t g0[8], g1[8], g2[8], g3[8];
        const uint32_t abase = pc.t_acc + lane_base + (uint32_t)(b * 128 + wg * 8);
        tmem_ld8(abase + 0, g0);
        tmem_ld8(abase + 32, g1);
        tmem_ld8(abase + 64, g2);
        tmem_ld8(abase + 96, g3);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pc.bars->acc_empty[b]);
        float4* G = reinterpret_cast<float4*>(pc.gbuf + (size_t)c * S::IMG) + (wg * 4) * TCM + row;
#pragma unroll
        for (int u = 0; u < 8; u += 2) {                          // unit pairs: plane 0 = {i, i', f, f'}, plane 1 = {g, g', o, o'}
            const float4 b0 = *reinterpret_cast<const float4*>(pc.bias_s + (c * 32 + wg * 8 + u) * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(pc.bias_s + (c * 32 + wg * 8 + u + 1) * 4);
            G[(u >> 1) * TCM] = make_float4(__uint_as_float(g0[u]) + b0.x, __uint_as_float(g0[u + 1]) + b1.x,
                                            __uint_as_float(g1[u]) + b0.y, __uint_as_float(g1[u + 1]) + b1.y);
            G[(16 + (u >> 1)) * TCM] = make_float4(__uint_as_float(g2[u]) + b0.z, __uint_as_float(g2[u + 1]) + b1.z,
                                                   __uint_as_float(g3[u]) + b0.w, __uint_as_float(g3[u + 1]) + b1.w);
        }
    }
    fence_proxy_async_all();                                   // the image is read back by bulk copies
}

// ------------------------------------------------------------------------------------------------ MMA issuer warp
// One part = accumulate A[128 x K] * W_part^T into a 128-column accumulator.  KIND 0: A = layer input image in
// shared memory (K = H), 1: A = window tile in shared memory (K = 16, bf16), 2: A = h_{t-1} in TMEM (K = H).
// Warp-uniform control flow (descriptors live in uniform registers); one elected lane issues.
template <int H, int KIND>
__device__ __forceinline__ void mma_part(TcBars* bars, uint32_t ring_a, uint32_t& ring_it, uint32_t acc, uint32_t a_hi,
                                         uint32_t a_lo, uint32_t first_acc, long long (&prof)[8], bool leader) {
    constexpr int NKT = (KIND == 1) ? 1 : H / 64;
    constexpr int KPER = (KIND == 1) ? 1 : 4;
    constexpr uint32_t IDESC = (KIND == 1) ? make_idesc_bf16(128, 128) : make_idesc_f16(128, 128);
    uint32_t accf = first_acc;
#pragma unroll
    for (int kt = 0; kt < NKT; ++kt) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {           // 0: B_hi stage (A_hi and A_lo), 1: B_lo stage (A_hi)
            const uint32_t slot = ring_it % TC_NST;
            TC_TWAIT(0, &bars->w_full[slot], (ring_it / TC_NST) & 1);
            tc_fence_after_sync();
            if (elect_one()) {
                const uint32_t bbase = ring_a + slot * TC_STAGE;
#pragma unroll
                for (int j = 0; j < KPER; ++j) {
                    const int k = kt * KPER + j;           // k-step (16 elements)
                    const uint64_t bdesc = kdesc(bbase + j * 4096);
                    if (KIND == 2) {
                        mma_ts(acc, a_hi + k * 8, bdesc, IDESC, (j == 0) ? accf : 1u);
                        if (half == 0) mma_ts(acc, a_lo + k * 8, bdesc, IDESC, 1u);
                    } else {
                        mma_ss(acc, kdesc(a_hi + k * 4096), bdesc, IDESC, (j == 0) ? accf : 1u);
                        if (half == 0) mma_ss(acc, kdesc(a_lo + k * 4096), bdesc, IDESC, 1u);
                    }
                }
                mma_commit_mc(&bars->w_empty[slot], (uint16_t)((1u << TC_CLUSTER) - 1));   // the slot is refilled by BOTH producers
            }
            __syncwarp();
            accf = 1u;
            ++ring_it;
        }
    }
}

// xhat = h_t W_o^T : 3-pass fp16 split, M=128 x N=16 x K=H, A = h_t (hi|lo) in TMEM, B = W_o images in smem
template <int H>
__device__ __forceinline__ void mma_xhat(TcBars* bars, uint32_t acc, uint32_t h_hi, uint32_t h_lo, uint32_t wo_a, bool leader) {
    using S = TcSmem<H>;
    constexpr uint32_t IDESC16 = make_idesc_f16(128, 16);
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < H / 16; ++k) {
            const uint64_t bhi = make_smem_desc(wo_a + k * 512, 256, 128);
            const uint64_t blo = make_smem_desc(wo_a + S::WOIMG + k * 512, 256, 128);
            mma_ts(acc, h_hi + k * 8, bhi, IDESC16, k > 0 ? 1u : 0u);
            mma_ts(acc, h_lo + k * 8, bhi, IDESC16, 1u);
            mma_ts(acc, h_hi + k * 8, blo, IDESC16, 1u);
        }
        mma_commit(&bars->xhat_full);
    }
    __syncwarp();
}

template <int H, int IN_KIND, bool LASTDEC>
__device__ __forceinline__ void mma_pass(const PassCtx& pc, uint32_t ring_it, Cnt2 n_in, Cnt2 n_acc, Cnt2 n_h, uint32_t n_xe,
                                         long long (&prof)[8]) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    constexpr int NFIRST = NCH < 2 ? NCH : 2;
    constexpr int INK = (IN_KIND == IN_X) ? 1 : 0;
    static_assert(!LASTDEC || NCH >= 2, "the output Linear borrows accumulator buffer 1");
    TcBars* bars = pc.bars;
    const bool leader = (threadIdx.x & 31) == 0;
    // warp-uniform by construction; the shuffles let the compiler keep descriptors in uniform registers
    const uint32_t ring_a = __shfl_sync(0xffffffffu, smem_u32(pc.ring), 0);
    const uint32_t in_a = __shfl_sync(0xffffffffu, smem_u32(pc.inbuf), 0);
    const uint32_t wo_a = __shfl_sync(0xffffffffu, smem_u32(pc.wo_img), 0);
    const uint32_t t_acc = __shfl_sync(0xffffffffu, pc.t_acc, 0);
    const uint32_t t_h = __shfl_sync(0xffffffffu, pc.t_h, 0);
    const int u_buf = __shfl_sync(0xffffffffu, pc.u_buf, 0);
    ring_it = __shfl_sync(0xffffffffu, ring_it, 0);
    const uint32_t acc1 = t_acc + 128;
    const int T = __shfl_sync(0xffffffffu, pc.T, 0);
    for (int t = 0; t < T; ++t) {
        const int tb = t & 1;
        const bool xh_step = LASTDEC && t > 0;                 // xhat_{t-1} is produced at the head of step t
        if (IN_KIND != IN_CONST) {
            TC_TWAIT(1, &bars->in_full[tb], n_in.get(tb) & 1);
            n_in.inc(tb);
            tc_fence_after_sync();
        }
        uint32_t a_hi, a_lo;
        if (IN_KIND == IN_X) { a_hi = in_a + tb * 2 * TC_XSTAGE; a_lo = a_hi + TC_XSTAGE; }
        else { a_hi = in_a + ((IN_KIND == IN_CONST) ? u_buf : tb) * S::IMG; a_lo = a_hi + S::IMGH; }
        const uint32_t h_hi = t_h + (uint32_t)(((t - 1) & 1) * H), h_lo = h_hi + H / 2;
        // input projections that do not depend on h_{t-1}: chunk 0 always, chunk 1 unless its accumulator
        // buffer first has to carry xhat_{t-1}
#pragma unroll
        for (int c = 0; c < NFIRST; ++c) {
            if (c == 1 && xh_step) break;
            TC_TWAIT(2, &bars->acc_empty[c & 1], (n_acc.get(c & 1) & 1) ^ 1);
            n_acc.inc(c & 1);
            tc_fence_after_sync();
            mma_part<H, INK>(bars, ring_a, ring_it, t_acc + (uint32_t)((c & 1) * 128), a_hi, a_lo, 0u, prof, leader);
        }
        if (NCH <= NFIRST && IN_KIND != IN_CONST && !xh_step) { if (elect_one()) mma_commit(&bars->in_empty[tb]); __syncwarp(); }
        if (t > 0) {
            TC_TWAIT(3, &bars->h_full[(t - 1) & 1], n_h.get((t - 1) & 1) & 1);
            n_h.inc((t - 1) & 1);
            tc_fence_after_sync();
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t acc = t_acc + (uint32_t)((c & 1) * 128);
            if (c == 1 && xh_step) {
                mbar_wait(&bars->xhat_empty, n_xe & 1);        // output group has read xhat_{t-1}
                ++n_xe;
                tc_fence_after_sync();
                mma_part<H, INK>(bars, ring_a, ring_it, acc, a_hi, a_lo, 0u, prof, leader);
                if (NCH <= NFIRST && IN_KIND != IN_CONST) { if (elect_one()) mma_commit(&bars->in_empty[tb]); __syncwarp(); }
            }
            if (c >= NFIRST) {
                TC_TWAIT(2, &bars->acc_empty[c & 1], (n_acc.get(c & 1) & 1) ^ 1);
                n_acc.inc(c & 1);
                tc_fence_after_sync();
                mma_part<H, INK>(bars, ring_a, ring_it, acc, a_hi, a_lo, 0u, prof, leader);
                if (c == NCH - 1 && IN_KIND != IN_CONST) { if (elect_one()) mma_commit(&bars->in_empty[tb]); __syncwarp(); }
            }
            if (t > 0) mma_part<H, 2>(bars, ring_a, ring_it, acc, h_hi, h_lo, 1u, prof, leader);
            if (elect_one()) mma_commit(&bars->acc_full[c & 1]);
            __syncwarp();
            if (c == 0 && xh_step) {
                // off the critical path (chunk 0's recurrent part is already queued): xhat_{t-1} into buffer 1
                TC_TWAIT(2, &bars->acc_empty[1], (n_acc.get(1) & 1) ^ 1);     // last chunk of step t-1 drained
                n_acc.inc(1);
                tc_fence_after_sync();
                mma_xhat<H>(bars, acc1, h_hi, h_lo, wo_a, leader);
            }
        }
    }
    // the last step's h_full arrivals are not consumed by a recurrent MMA: consume them here so the phase
    // bookkeeping stays aligned for the next pass (and xhat_{T-1} needs h_{T-1} anyway)
    mbar_wait(&bars->h_full[(T - 1) & 1], n_h.get((T - 1) & 1) & 1);
    if (LASTDEC) {
        tc_fence_after_sync();
        mbar_wait(&bars->acc_empty[1], (n_acc.get(1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t h_hi = t_h + (uint32_t)(((T - 1) & 1) * H);
        mma_xhat<H>(bars, acc1, h_hi, h_hi + H / 2, wo_a, leader);
    }
}

// IN_HOIST: (1) once per tile, the input projection of every chunk into the accumulators (drained by epi_hoist_pre) ...
template <int H>
__device__ __forceinline__ uint32_t mma_hoist_pre(const PassCtx& pc, uint32_t ring_it, Cnt2& n_acc, long long (&prof)[8]) {
    using S = TcSmem<H>;
    TcBars* bars = pc.bars;
    const bool leader = (threadIdx.x & 31) == 0;
    const uint32_t ring_a = __shfl_sync(0xffffffffu, smem_u32(pc.ring), 0);
    const uint32_t in_a = __shfl_sync(0xffffffffu, smem_u32(pc.inbuf), 0);
    const uint32_t t_acc = __shfl_sync(0xffffffffu, pc.t_acc, 0);
    const int u_buf = __shfl_sync(0xffffffffu, pc.u_buf, 0);
    ring_it = __shfl_sync(0xffffffffu, ring_it, 0);
    const uint32_t a_hi = in_a + u_buf * S::IMG, a_lo = a_hi + S::IMGH;
#pragma unroll 1
    for (int c = 0; c < S::NCH; ++c) {
        TC_TWAIT(2, &bars->acc_empty[c & 1], (n_acc.get(c & 1) & 1) ^ 1);
        n_acc.inc(c & 1);
        tc_fence_after_sync();
        mma_part<H, 0>(bars, ring_a, ring_it, t_acc + (uint32_t)((c & 1) * 128), a_hi, a_lo, 0u, prof, leader);
        if (elect_one()) mma_commit(&bars->acc_full[c & 1]);
        __syncwarp();
    }
    return ring_it;
}

// ... (2) per step only the recurrent part, starting the accumulator afresh (first_acc = 0); at t = 0 there is nothing to issue.
template <int H>
__device__ __forceinline__ void mma_pass_hoist(const PassCtx& pc, uint32_t ring_it, Cnt2 n_acc, Cnt2 n_h, long long (&prof)[8]) {
    using S = TcSmem<H>;
    TcBars* bars = pc.bars;
    const bool leader = (threadIdx.x & 31) == 0;
    const uint32_t ring_a = __shfl_sync(0xffffffffu, smem_u32(pc.ring), 0);
    const uint32_t t_acc = __shfl_sync(0xffffffffu, pc.t_acc, 0);
    const uint32_t t_h = __shfl_sync(0xffffffffu, pc.t_h, 0);
    ring_it = __shfl_sync(0xffffffffu, ring_it, 0);
    const int T = __shfl_sync(0xffffffffu, pc.T, 0);
    for (int t = 0; t < T; ++t) {
        const uint32_t h_hi = t_h + (uint32_t)(((t - 1) & 1) * H), h_lo = h_hi + H / 2;
        if (t > 0) {
            TC_TWAIT(3, &bars->h_full[(t - 1) & 1], n_h.get((t - 1) & 1) & 1);
            n_h.inc((t - 1) & 1);
            tc_fence_after_sync();
        }
#pragma unroll
        for (int c = 0; c < S::NCH; ++c) {
            TC_TWAIT(2, &bars->acc_empty[c & 1], (n_acc.get(c & 1) & 1) ^ 1);
            n_acc.inc(c & 1);
            tc_fence_after_sync();
            if (t > 0) mma_part<H, 2>(bars, ring_a, ring_it, t_acc + (uint32_t)((c & 1) * 128), h_hi, h_lo, 0u, prof, leader);
            if (elect_one()) mma_commit(&bars->acc_full[c & 1]);
            __syncwarp();
        }
    }
    mbar_wait(&bars->h_full[(T - 1) & 1], n_h.get((T - 1) & 1) & 1);      // phase bookkeeping, as in mma_pass
}

// ------------------------------------------------------------------------------------------------ copy producer (one lane)
template <int H>
__device__ __forceinline__ void prod_pass(const PassCtx& pc, const unsigned char* w, int in_kind, bool lastdec, uint32_t ring_it,
                                          Cnt2 n_in) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    constexpr int NFIRST = NCH < 2 ? NCH : 2;
    constexpr int KT_PER_PART = H / 64;
    TcBars* bars = pc.bars;
    const int T = pc.T;
    const int part_in_bytes = (in_kind == IN_X) ? 2 * TC_XSTAGE : KT_PER_PART * 2 * TC_STAGE;
    const int part_hh_bytes = KT_PER_PART * 2 * TC_STAGE;
    // Every CTA of the cluster consumes the same weight stream in the same order: each loads 1/TC_CLUSTER of a stage and
    // multicasts it into the same ring slot of all of them (one L2 read per cluster instead of one per CTA).  A slot is
    // free when the MMA issuers of ALL CTAs have committed it (w_empty counts TC_CLUSTER multicast commits).
    const uint32_t crank = cluster_ctarank();
    auto load_part = [&](const unsigned char* g, int nstage, uint32_t bytes) {
        const uint32_t part = bytes / TC_CLUSTER;
        for (int s = 0; s < nstage; ++s) {
            const uint32_t slot = ring_it % TC_NST;
            mbar_wait(&bars->w_empty[slot], ((ring_it / TC_NST) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars->w_full[slot], bytes);
            bulk_g2s_mc(pc.ring + slot * TC_STAGE + crank * part, g + (size_t)s * bytes + crank * part, part, &bars->w_full[slot],
                        (uint16_t)((1u << TC_CLUSTER) - 1));
            ++ring_it;
        }
    };
    auto load_in = [&](int t) {
        const int b = t & 1;
        mbar_wait(&bars->in_empty[b], (n_in.get(b) & 1) ^ 1);
        n_in.inc(b);
        mbar_arrive_expect_tx(&bars->in_full[b], (uint32_t)S::IMG);
        const unsigned char* g = pc.scratch + (size_t)t * S::IMG;
        for (int q = 0; q < S::IMG / 16384; ++q) bulk_g2s(pc.inbuf + b * S::IMG + q * 16384, g + q * 16384, 16384, &bars->in_full[b]);
    };
    const int in_nst = (in_kind == IN_X) ? 2 : KT_PER_PART * 2;
    const uint32_t in_sb = (in_kind == IN_X) ? TC_XSTAGE : TC_STAGE;
    const size_t chunk_bytes = (size_t)part_in_bytes + part_hh_bytes;
    if (in_kind == IN_STREAM) load_in(0);
    for (int t = 0; t < T; ++t) {
        if (in_kind == IN_STREAM && t + 1 < T) load_in(t + 1);
        // same order as the MMA issuer: input parts of chunks 0/1 lead, except in the decoder's last layer where
        // chunk 1's accumulator first carries xhat_{t-1} (its input part then follows chunk 0's recurrent part)
        const int nlead = (lastdec && t > 0) ? 1 : NFIRST;
        for (int c = 0; c < nlead; ++c) load_part(w + c * chunk_bytes, in_nst, in_sb);
        for (int c = 0; c < NCH; ++c) {
            if (c >= nlead) load_part(w + c * chunk_bytes, in_nst, in_sb);
            if (t > 0) load_part(w + c * chunk_bytes + part_in_bytes, KT_PER_PART * 2, TC_STAGE);
        }
    }
}

// IN_HOIST producer: (pre) the input-part weight stages of every chunk, once per tile; (main) per step and chunk the chunk's G image
// into input buffer c & 1 (this CTA's own data: plain bulk copies) and, from the second step on, its recurrent weight stages.
template <int H>
__device__ __forceinline__ uint32_t prod_hoist(const PassCtx& pc, const unsigned char* w, uint32_t ring_it, bool pre, Cnt2 n_g) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    constexpr int KT_PER_PART = H / 64;
    TcBars* bars = pc.bars;
    const int part_bytes = KT_PER_PART * 2 * TC_STAGE;         // input part == recurrent part: [128 x H] hi|lo per chunk
    const size_t chunk_bytes = (size_t)2 * part_bytes;
    const uint32_t crank = cluster_ctarank();
    auto load_part = [&](const unsigned char* g) {
        const uint32_t part = TC_STAGE / TC_CLUSTER;
        for (int s = 0; s < KT_PER_PART * 2; ++s) {
            const uint32_t slot = ring_it % TC_NST;
            mbar_wait(&bars->w_empty[slot], ((ring_it / TC_NST) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars->w_full[slot], TC_STAGE);
            bulk_g2s_mc(pc.ring + slot * TC_STAGE + crank * part, g + (size_t)s * TC_STAGE + crank * part, part, &bars->w_full[slot],
                        (uint16_t)((1u << TC_CLUSTER) - 1));
            ++ring_it;
        }
    };
    if (pre) {
        for (int c = 0; c < NCH; ++c) load_part(w + c * chunk_bytes);
        return ring_it;
    }
    for (int t = 0; t < pc.T; ++t)
        for (int c = 0; c < NCH; ++c) {
            const int b = c & 1;
            mbar_wait(&bars->g_empty[b], (n_g.get(b) & 1) ^ 1);
            n_g.inc(b);
            mbar_arrive_expect_tx(&bars->g_full[b], (uint32_t)S::IMG);
            const unsigned char* g = pc.gbuf + (size_t)c * S::IMG;
            for (int q = 0; q < S::IMG / 16384; ++q) bulk_g2s(pc.inbuf + b * S::IMG + q * 16384, g + q * 16384, 16384, &bars->g_full[b]);
            if (t > 0) load_part(w + c * chunk_bytes + part_bytes);
        }
    return ring_it;
}

// ------------------------------------------------------------------------------------------------ staging / output group (warps 8-11)
// Encoder input: x_t (gather + normalise fused) -> bf16 hi|lo A image [128 x 16]; one window row per thread.
__device__ __forceinline__ void aux_stage_pass(const PassCtx& pc, const VaeDev& P, const WinSrc& src, const VaeIO& io, Cnt2 n_in,
                                               long long (&prof)[8]) {
    TcBars* bars = pc.bars;
    const int row = threadIdx.x - TC_WARP_AUX0 * 32;
    const bool ok = row < pc.nvalid;
    const long long n = pc.n0 + (ok ? row : 0);
    const long long win = (ok && io.idx) ? (long long)io.idx[n] : n;
    const float* wbase = src.base + win * src.win_stride;
    for (int t = 0; t < pc.T; ++t) {
        const int b = t & 1;
        float raw[SHM_MAX_D];
#pragma unroll
        for (int d = 0; d < SHM_MAX_D; ++d) raw[d] = (ok && d < P.D) ? __ldg(wbase + (long long)t * src.row_stride + src.chan[d]) : 0.f;
        float v[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) v[d] = (ok && d < P.D) ? win_transform_fast(src, raw[d], d) : 0.f;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
        const long long tw0 = TC_CLOCK();
        mbar_wait(&bars->in_empty[b], (n_in.get(b) & 1) ^ 1);
        prof[0] += TC_CLOCK() - tw0;
        n_in.inc(b);
        unsigned char* xhi = pc.inbuf + b * 2 * TC_XSTAGE;
        unsigned char* xlo = xhi + TC_XSTAGE;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int off = (q * 16 + (row >> 3)) * 128 + (row & 7) * 16;
            *reinterpret_cast<uint4*>(xhi + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
            *reinterpret_cast<uint4*>(xlo + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
        }
        fence_proxy_async_smem();
        aux_bar_sync();
        if (row == 0) mbar_arrive(&bars->in_full[b]);
    }
}

// Decoder output: the MMA warp computes xhat_t = h_t W_o^T on the tensor core (16 spare accumulator columns);
// this group adds the bias, forms the squared error against the window, optionally stores the
// reconstruction / CNN input, and writes the per-window score.  One window row per thread.
__device__ __forceinline__ float ld_nc_volatile(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int H>
__device__ __forceinline__ void aux_out_pass(const PassCtx& pc, const VaeDev& P, const WinSrc& src, const VaeIO& io, uint32_t n_x,
                                             long long (&prof)[8]) {
    TcBars* bars = pc.bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = threadIdx.x - TC_WARP_AUX0 * 32;
    const uint32_t xh_addr = pc.t_acc + 128 + ((uint32_t)((warp & 3) * 32) << 16);      // accumulator buffer 1, columns 0..15
    const bool ok = row < pc.nvalid;
    const long long n = pc.n0 + (ok ? row : 0);
    const long long win = (ok && io.idx) ? (long long)io.idx[n] : n;
    const float* wbase = src.base + win * src.win_stride;
    const int T = pc.T;
    const bool exact = io.cnn_in != nullptr || io.recon != nullptr;     // windows handed back: bit-exact transform
    float sse = 0.f;
    for (int t = 0; t < T; ++t) {
        float x[SHM_MAX_D];                                    // issued before the wait so the latency is hidden
#pragma unroll
        for (int d = 0; d < SHM_MAX_D; ++d) x[d] = (ok && d < P.D) ? ld_nc_volatile(wbase + (long long)t * src.row_stride + src.chan[d]) : 0.f;
        const long long tw0 = TC_CLOCK();
        mbar_wait(&bars->xhat_full, n_x & 1);
        prof[0] += TC_CLOCK() - tw0;
        ++n_x;
        tc_fence_after_sync();
        uint32_t acc[16];
        tmem_ld16(xh_addr, acc);
        tmem_ld_wait();
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->xhat_empty);
#pragma unroll
        for (int d = 0; d < SHM_MAX_D; ++d) {
            if (d < P.D) {
                const float y = __uint_as_float(acc[d]) + pc.bo_s[d];
                const float xv = !ok ? 0.f : (exact ? win_transform(src, x[d], d) : win_transform_fast(src, x[d], d));
                const float e = xv - y;
                sse = fmaf(e, e, sse);
                if (ok) {
                    if (io.recon) io.recon[(n * T + t) * P.D + d] = y;
                    if (io.cnn_in) {
                        io.cnn_in[((n * 2 + 0) * T + t) * P.D + d] = xv;
                        io.cnn_in[((n * 2 + 1) * T + t) * P.D + d] = e * e;
                    }
                }
            }
        }
    }
    if (ok && io.score) io.score[n] = sse / (float)(T * P.D);
}

// ------------------------------------------------------------------------------------------------ heads (epilogue warps)
// LayerNorm -> mu/logvar -> z = mu + eps*exp(0.5 logvar) -> u = tanh(W z + b) as the decoder's constant A operand
template <int H>
__device__ __forceinline__ void heads_stage(const PassCtx& pc, const VaeDev& P, const VaeIO& io) {
    using S = TcSmem<H>;
    const int tid = threadIdx.x;
    const int nvalid = pc.nvalid;
    const long long n0 = pc.n0;
    float* hT = reinterpret_cast<float*>(pc.inbuf + pc.hT_buf * S::IMG);      // [H][128] fp32
    // [2Z][128] in a per-CTA GLOBAL scratch: the weight ring cannot be borrowed here, the peer CTA's producer may already be
    // multicasting the next pass's first stages into it
    float* muS = pc.heads_scratch;
    float* zS = muS + 2 * VAE_MAX_Z * TCM;                                    // [Z][128]
    if (io.mu_in) {
        // re-score: the encoder ran in an earlier call; fetch its mu / logvar for this tile's windows
        for (int item = tid; item < 2 * P.Z * TCM; item += TC_EPI_THREADS) {
            const int o = item / TCM, w = item - o * TCM;
            const bool is_lv = o >= P.Z;
            const int zi = is_lv ? o - P.Z : o;
            float y = 0.f;
            if (w < nvalid) {
                const long long win = io.idx ? (long long)io.idx[n0 + w] : n0 + w;
                y = __ldg((is_lv ? io.logvar_in : io.mu_in) + win * P.Z + zi);
            }
            muS[o * TCM + w] = y;
        }
    } else {
    if (P.has_ln) {
        if (tid < TCM) {
            float m = 0.f;
            for (int k = 0; k < H; ++k) m += hT[k * TCM + tid];
            m /= (float)H;
            float v = 0.f;
            for (int k = 0; k < H; ++k) { const float d = hT[k * TCM + tid] - m; v = fmaf(d, d, v); }
            v /= (float)H;
            const float rstd = 1.0f / sqrtf(v + P.ln_eps);
            for (int k = 0; k < H; ++k)
                hT[k * TCM + tid] = fmaf((hT[k * TCM + tid] - m) * rstd, __ldg(P.ln_w + k), __ldg(P.ln_b + k));
        }
        epi_bar_sync();
    }
    for (int item = tid; item < 2 * P.Z * TCM; item += TC_EPI_THREADS) {
        const int o = item / TCM, w = item - o * TCM;
        const bool is_lv = o >= P.Z;
        const int zi = is_lv ? o - P.Z : o;
        const float* wr = (is_lv ? P.lv_w : P.mu_w) + zi * H;
        float y = __ldg((is_lv ? P.lv_b : P.mu_b) + zi);
        for (int k = 0; k < H; ++k) y = fmaf(hT[k * TCM + w], __ldg(wr + k), y);
        muS[o * TCM + w] = y;
        if (w < nvalid) {
            float* dst = is_lv ? io.logvar : io.mu;
            if (dst) dst[(n0 + w) * P.Z + zi] = y;
        }
    }
    }
    epi_bar_sync();
    for (int item = tid; item < P.Z * TCM; item += TC_EPI_THREADS) {
        const int zi = item / TCM, w = item - zi * TCM;
        const float m = muS[zi * TCM + w], lv = muS[(P.Z + zi) * TCM + w];
        float z = m;
        if (io.eps && w < nvalid) z = fmaf(__ldg(io.eps + (n0 + w) * P.Z + zi), expf(0.5f * lv), m);
        zS[zi * TCM + w] = z;
    }
    epi_bar_sync();
    unsigned short* uhi = reinterpret_cast<unsigned short*>(pc.inbuf + pc.u_buf * S::IMG);
    unsigned short* ulo = reinterpret_cast<unsigned short*>(pc.inbuf + pc.u_buf * S::IMG + S::IMGH);
    for (int item = tid; item < H * TCM; item += TC_EPI_THREADS) {
        const int k = item / TCM, w = item - k * TCM;
        float y = __ldg(P.l2h_b + k);
        const float* wr = P.l2h_w + k * P.Z;
        for (int zi = 0; zi < P.Z; ++zi) y = fmaf(zS[zi * TCM + w], __ldg(wr + zi), y);
        const float u = tanhf(y);
        const __half bh = __float2half_rn(u);
        const __half bl = __float2half_rn(u - __half2float(bh));
        const int off = ((k >> 3) * 16 + (w >> 3)) * 64 + (w & 7) * 8 + (k & 7);
        uhi[off] = *reinterpret_cast<const unsigned short*>(&bh);
        ulo[off] = *reinterpret_cast<const unsigned short*>(&bl);
    }
    fence_proxy_async_smem();
}

template <int H>
__global__ void __cluster_dims__(TC_CLUSTER, 1, 1) __launch_bounds__(TC_THREADS, 1)
vae_score_tc_kernel(VaeDev P, TcDev TC, WinSrc src, VaeIO io) {
    using S = TcSmem<H>;
    constexpr int NCH = S::NCH;
    constexpr int KT_PER_PART = H / 64;                      // 64-wide K tiles per H-wide part
    extern __shared__ __align__(1024) unsigned char smem[];
    TcBars* bars = reinterpret_cast<TcBars*>(smem + S::off_bar);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + S::off_bar + 32 * 8);
    unsigned short* wo_hi = reinterpret_cast<unsigned short*>(smem + S::off_wo);
    unsigned short* wo_lo = wo_hi + 16 * H;
    float* bo_s = reinterpret_cast<float*>(smem + S::off_wo + 2 * S::WOIMG);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = src.T;
    long long n_eff = io.n;
    if (io.n_dev) n_eff = min(n_eff, (long long)__ldg(io.n_dev));
    const int n_tiles = (int)((n_eff + TCM - 1) / TCM);

    if (tid == 0) {
        for (int i = 0; i < TC_NST; ++i) { mbar_init(&bars->w_full[i], 1); mbar_init(&bars->w_empty[i], TC_CLUSTER); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->in_full[i], 1); mbar_init(&bars->in_empty[i], 1);
            mbar_init(&bars->acc_full[i], 1); mbar_init(&bars->acc_empty[i], TC_EPI_WARPS);
            mbar_init(&bars->g_full[i], 1); mbar_init(&bars->g_empty[i], TC_EPI_WARPS);
            mbar_init(&bars->h_full[i], TC_EPI_WARPS);
        }
        mbar_init(&bars->xhat_full, 1); mbar_init(&bars->xhat_empty, 4);
        fence_mbar_init();
    }
    if (warp == TC_WARP_MMA) tmem_alloc(tmem_holder, 512);
    for (int i = tid; i < 16 * H; i += TC_THREADS) {           // W_o [D,H] -> fp16 hi|lo images, rows d >= D zero
        const int d = i / H, k = i - d * H;
        const float w = d < P.D ? __ldg(P.out_w + d * H + k) : 0.f;
        const __half bh = __float2half_rn(w);
        const __half bl = __float2half_rn(w - __half2float(bh));
        const int off = ((k >> 3) * 2 + (d >> 3)) * 64 + (d & 7) * 8 + (k & 7);
        wo_hi[off] = *reinterpret_cast<const unsigned short*>(&bh);
        wo_lo[off] = *reinterpret_cast<const unsigned short*>(&bl);
    }
    if (tid < 16) bo_s[tid] = tid < P.D ? __ldg(P.out_b + tid) : 0.f;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // every CTA's mbarriers are initialised before a peer multicasts / commits into them
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_holder;
    // the CTAs of a cluster run in lock step over the weight stream: same number of tile iterations, dummy tiles
    // (nvalid = 0) where a CTA has no tile left
    const int n_iter = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

    PassCtx pc;
    pc.bars = bars; pc.ring = smem + S::off_ring; pc.inbuf = smem + S::off_in;
    pc.bias_s = reinterpret_cast<float*>(smem + S::off_bias);
    pc.wo_img = smem + S::off_wo; pc.bo_s = bo_s;
    pc.scratch = TC.scratch + (size_t)blockIdx.x * (TC.scratch_stride + TC_HEADS_SCRATCH + TC.g_bytes);
    pc.heads_scratch = reinterpret_cast<float*>(pc.scratch + TC.scratch_stride);
    pc.gbuf = pc.scratch + TC.scratch_stride + TC_HEADS_SCRATCH;
    pc.t_acc = tbase;                  // 2 x 128 accumulator columns
    pc.t_h = tbase + 256;              // 2 x H columns: h_t as fp16 pairs, [hi H/2 | lo H/2]
    pc.T = T;

    // Use-counters of the mbarriers.  Every thread carries the same pass-entry values ("base") and
    // advances them analytically at the end of each pass, so roles that skip a barrier in one pass
    // still know its phase in the next; inside a pass each role counts its own uses from the base.
    uint32_t ring_base = 0;
    Cnt2 in_base{0, 0}, acc_base{0, 0}, h_base{0, 0}, g_base{0, 0};
    uint32_t xhat_base = 0;
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};

    // per-pass bookkeeping shared by both role groups
    auto pass_setup = [&](int p, int& in_kind, int& sink) {
        in_kind = TC.pass[p].in_kind; sink = TC.pass[p].sink;
        if (sink == SINK_LAST_ENC) pc.hT_buf = (in_kind == IN_STREAM) ? (T & 1) : 1;
        pc.u_buf = pc.hT_buf ^ 1;
    };
    // IN_HOIST runs a one-sweep precompute (input parts only, one accumulator use per chunk) in front of the pass proper
    auto hoist_pre_advance = [&](uint32_t& ring, Cnt2& acc) {
        ring += (uint32_t)NCH * KT_PER_PART * 2;
        acc.n0 += (NCH + 1) / 2;
        acc.n1 += NCH / 2;
    };
    auto pass_advance = [&](int in_kind, int sink) {
        const uint32_t even = (uint32_t)((T + 1) / 2), odd = (uint32_t)(T / 2);
        if (in_kind == IN_HOIST) {
            hoist_pre_advance(ring_base, acc_base);
            ring_base += (uint32_t)(T - 1) * NCH * KT_PER_PART * 2;
            g_base.n0 += (uint32_t)T * ((NCH + 1) / 2);
            g_base.n1 += (uint32_t)T * (NCH / 2);
        } else {
            const uint32_t stages_step0 = (uint32_t)NCH * ((in_kind == IN_X) ? 2 : KT_PER_PART * 2);
            const uint32_t stages_step = stages_step0 + (uint32_t)NCH * KT_PER_PART * 2;
            ring_base += stages_step0 + (uint32_t)(T - 1) * stages_step;
        }
        if (in_kind != IN_CONST && in_kind != IN_HOIST) { in_base.n0 += even; in_base.n1 += odd; }
        acc_base.n0 += (uint32_t)T * ((NCH + 1) / 2);
        acc_base.n1 += (uint32_t)T * (NCH / 2);
        h_base.n0 += even; h_base.n1 += odd;
        if (sink == SINK_LAST_DEC) xhat_base += (uint32_t)T;
    };
    const bool encode_only_call = !io.score && !io.recon && !io.cnn_in;
    const int p_first = io.mu_in ? TC.L : 0;       // re-score: mu / logvar are given, the encoder passes are skipped

    // Role groups at top level so that each group's code is dominated by its setmaxnreg: the two epilogue
    // warpgroups take 208 registers per thread (cell state + accumulator slices + deep ILP), the
    // producer / MMA / staging warpgroup drops to 88.  CTA-wide phases meet at barrier 0.
    if (warp < TC_EPI_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REG_EPI));
        for (int it = 0; it < n_iter; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            pc.n0 = tile < n_tiles ? (long long)tile * TCM : 0;
            pc.nvalid = tile < n_tiles ? (int)min((long long)TCM, n_eff - pc.n0) : 0;
            pc.hT_buf = 1;                         // in-buffer that receives the encoder's fp32 h_T
            for (int p = p_first; p < TC.n_pass; ++p) {
                int in_kind, sink;
                pass_setup(p, in_kind, sink);
                if (p == TC.L) heads_stage<H>(pc, P, io);
                cta_sync();                        // (A) previous pass / heads complete and visible
                if (p == TC.L && encode_only_call) break;
                const long long pass_t0 = TC_CLOCK();
                if (in_kind == IN_HOIST) {
                    epi_hoist_pre<H>(pc, TC.pass[p].bias, acc_base);
                    cta_sync();                    // (A2) G is in global memory and the u image is dead: the input buffers now stage G
                    uint32_t ring_dummy = 0;
                    Cnt2 acc2 = acc_base;
                    hoist_pre_advance(ring_dummy, acc2);
                    epi_pass<H, SINK_STREAM, true>(pc, P, src, io, TC.pass[p].bias, acc2, prof, g_base);
                } else if (sink == SINK_STREAM) epi_pass<H, SINK_STREAM>(pc, P, src, io, TC.pass[p].bias, acc_base, prof);
                else if (sink == SINK_LAST_ENC) epi_pass<H, SINK_LAST_ENC>(pc, P, src, io, TC.pass[p].bias, acc_base, prof);
                else epi_pass<H, SINK_LAST_DEC>(pc, P, src, io, TC.pass[p].bias, acc_base, prof);
                { const long long _d = TC_CLOCK() - pass_t0; const int _q = p & 3;
                  if (_q == 0) prof[4] += _d; else if (_q == 1) prof[5] += _d; else if (_q == 2) prof[6] += _d; else prof[7] += _d; }
                cta_sync();                        // (B) end of pass
                pass_advance(in_kind, sink);
            }
        }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REG_AUX));
        for (int it = 0; it < n_iter; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            pc.n0 = tile < n_tiles ? (long long)tile * TCM : 0;
            pc.nvalid = tile < n_tiles ? (int)min((long long)TCM, n_eff - pc.n0) : 0;
            pc.hT_buf = 1;
            for (int p = p_first; p < TC.n_pass; ++p) {
                int in_kind, sink;
                pass_setup(p, in_kind, sink);
                cta_sync();                        // (A)
                if (p == TC.L && encode_only_call) break;
                if (in_kind == IN_HOIST) {
                    uint32_t ring2 = ring_base;
                    Cnt2 acc2 = acc_base;
                    hoist_pre_advance(ring2, acc2);
                    if (warp == TC_WARP_PROD) { if (lane == 0) prod_hoist<H>(pc, TC.pass[p].w, ring_base, true, g_base); }
                    else if (warp == TC_WARP_MMA) { Cnt2 a = acc_base; mma_hoist_pre<H>(pc, ring_base, a, prof); }
                    __syncwarp();
                    cta_sync();                    // (A2)
                    if (warp == TC_WARP_PROD) { if (lane == 0) prod_hoist<H>(pc, TC.pass[p].w, ring2, false, g_base); }
                    else if (warp == TC_WARP_MMA) {
                        const long long pass_t0 = TC_CLOCK();
                        mma_pass_hoist<H>(pc, ring2, acc2, h_base, prof);
                        { const long long _d = TC_CLOCK() - pass_t0; const int _q = p & 3;
                          if (_q == 0) prof[4] += _d; else if (_q == 1) prof[5] += _d; else if (_q == 2) prof[6] += _d; else prof[7] += _d; }
                    }
                } else if (warp == TC_WARP_PROD) {
                    if (lane == 0) prod_pass<H>(pc, TC.pass[p].w, in_kind, sink == SINK_LAST_DEC, ring_base, in_base);
                } else if (warp == TC_WARP_MMA) {
                    const long long pass_t0 = TC_CLOCK();
                    if (sink == SINK_LAST_DEC) {
                        if (in_kind == IN_STREAM) mma_pass<H, IN_STREAM, true>(pc, ring_base, in_base, acc_base, h_base, xhat_base, prof);
                        else mma_pass<H, IN_CONST, true>(pc, ring_base, in_base, acc_base, h_base, xhat_base, prof);
                    } else if (in_kind == IN_X) mma_pass<H, IN_X, false>(pc, ring_base, in_base, acc_base, h_base, xhat_base, prof);
                    else if (in_kind == IN_STREAM) mma_pass<H, IN_STREAM, false>(pc, ring_base, in_base, acc_base, h_base, xhat_base, prof);
                    else mma_pass<H, IN_CONST, false>(pc, ring_base, in_base, acc_base, h_base, xhat_base, prof);
                    { const long long _d = TC_CLOCK() - pass_t0; const int _q = p & 3;
                  if (_q == 0) prof[4] += _d; else if (_q == 1) prof[5] += _d; else if (_q == 2) prof[6] += _d; else prof[7] += _d; }
                } else if (warp >= TC_WARP_AUX0 && warp < TC_WARP_AUX0 + 4) {
                    if (in_kind == IN_X) aux_stage_pass(pc, P, src, io, in_base, prof);
                    if (sink == SINK_LAST_DEC) aux_out_pass<H>(pc, P, src, io, xhat_base, prof);
                }
                cta_sync();                        // (B)
                pass_advance(in_kind, sink);
            }
        }
    }
    if (TC.dbg && lane == 0 && (warp == TC_WARP_MMA || warp == TC_WARP_AUX0 || warp == 0)) {
        const int role = (warp == TC_WARP_MMA) ? 0 : (warp == TC_WARP_AUX0 ? 1 : 2);
        for (int i = 0; i < 8; ++i) TC.dbg[(blockIdx.x * 3 + role) * 8 + i] = prof[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();                 // no CTA leaves while a peer can still write into its shared memory
    if (warp == TC_WARP_MMA) tmem_dealloc(tbase, 512);
}

}  // namespace shm
#include "vae_tc_dual.cuh"
namespace shm {

// ------------------------------------------------------------------------------------------------
// weight repacking: state_dict fp32 -> pre-scaled bf16 hi|lo stage images in UMMA K-major layout
//   out, per chunk c: [in part: kt x {hi,lo} images of [128 x KT_in]] [hh part: kt x {hi,lo} of [128 x 64]]
//   N-row n of a chunk = gate (n / 32), hidden unit c*32 + n % 32.
// ------------------------------------------------------------------------------------------------
__global__ void tc_pack_pass_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int Kin, int H,
                                    int x_input, unsigned short* __restrict__ out, float* __restrict__ bias) {
    const int NCH = H / 32;
    const int Kin_pad = x_input ? 16 : H;
    const int kt_in = x_input ? 16 : 64;
    const int in_elems = 128 * Kin_pad * 2;              // hi + lo
    const int hh_elems = 128 * H * 2;
    const int chunk_elems = in_elems + hh_elems;
    const int per_chunk_work = 128 * (Kin_pad + H);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < NCH * per_chunk_work; i += gridDim.x * blockDim.x) {
        const int c = i / per_chunk_work;
        int r = i - c * per_chunk_work;
        const int n = r / (Kin_pad + H);
        const int kk = r - n * (Kin_pad + H);
        const int gate = n >> 5;
        const int row = gate * H + c * 32 + (n & 31);
        const float scale = (gate == 2) ? 2.f * NLOG2E : NLOG2E;
        float w;
        int base, k, ktile;
        if (kk < Kin_pad) {
            k = kk; ktile = kt_in;
            w = (k < Kin) ? w_ih[(size_t)row * Kin + k] : 0.f;
            base = c * chunk_elems;
        } else {
            k = kk - Kin_pad; ktile = 64;
            w = w_hh[(size_t)row * H + k];
            base = c * chunk_elems + in_elems;
        }
        w *= scale;
        unsigned short sh, sl;
        if (x_input && kk < Kin_pad) {              // pairs with the bf16 window operand
            const __nv_bfloat16 bh = __float2bfloat16_rn(w);
            const __nv_bfloat16 bl = __float2bfloat16_rn(w - __bfloat162float(bh));
            sh = *reinterpret_cast<const unsigned short*>(&bh); sl = *reinterpret_cast<const unsigned short*>(&bl);
        } else {
            const __half bh = __float2half_rn(w);
            const __half bl = __float2half_rn(w - __half2float(bh));
            sh = *reinterpret_cast<const unsigned short*>(&bh); sl = *reinterpret_cast<const unsigned short*>(&bl);
        }
        const int kt = k / ktile, kl = k - kt * ktile;
        const int stage_elems = 128 * ktile;
        const int off = base + kt * 2 * stage_elems + ((kl >> 3) * 16 + (n >> 3)) * 64 + (n & 7) * 8 + (kl & 7);
        out[off] = sh;
        out[off + stage_elems] = sl;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * H; i += gridDim.x * blockDim.x) {
        const int u = i >> 2, g = i & 3;
        const float scale = (g == 2) ? 2.f * NLOG2E : NLOG2E;
        bias[i] = (b_ih[g * H + u] + b_hh[g * H + u]) * scale;
    }
}

static size_t tc_pass_bytes(int H, bool x_input) { return (size_t)(H / 32) * 128 * ((x_input ? 16 : H) + H) * 2 * 2; }

bool vae_tc_supported(const shm_vae_cfg& cfg) {
    return (cfg.H == 128 || cfg.H == 64) && cfg.L >= 1 && cfg.L <= SHM_MAX_L && cfg.D <= 16 && cfg.Z <= VAE_MAX_Z;
}

int vae_tc_alloc(VaeTc* tc, const shm_vae_cfg& cfg) {
    size_t bytes = 0;
    for (int l = 0; l < cfg.L; ++l) bytes += tc_pass_bytes(cfg.H, l == 0) + tc_pass_bytes(cfg.H, false);
    tc->wpack_bytes = bytes;
    if (cudaMalloc(&tc->wpack, bytes) != cudaSuccess || cudaMalloc(&tc->bias, (size_t)2 * cfg.L * 4 * cfg.H * sizeof(float)) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaMalloc(vae tc weights)");
        return SHM_ERR_NOMEM;
    }
    cudaError_t e = cudaSuccess;
    if (cfg.H == 128) e = cudaFuncSetAttribute(vae_score_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<128>::total);
    else e = cudaFuncSetAttribute(vae_score_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<64>::total);
    if (e == cudaSuccess && cfg.H == 64)
        e = cudaFuncSetAttribute(vae_score_tc_dual_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, D2Smem::total);
    if (e != cudaSuccess) { set_cuda_error(e, "cudaFuncSetAttribute(vae_score_tc)"); return SHM_ERR_CUDA; }
    tc->H = cfg.H; tc->L = cfg.L; tc->D = cfg.D;
    return SHM_OK;
}

int vae_tc_pack(VaeTc* tc, const shm_vae_cfg& cfg, const VaeTcRaw& raw, cudaStream_t st) {
    const int H = cfg.H, L = cfg.L;
    size_t off = 0;
    for (int p = 0; p < 2 * L; ++p) {
        const bool dec = p >= L;
        const int l = dec ? p - L : p;
        const bool xin = (!dec && l == 0);
        const float* wih = dec ? raw.dec_wih[l] : raw.enc_wih[l];
        const float* whh = dec ? raw.dec_whh[l] : raw.enc_whh[l];
        const float* bih = dec ? raw.dec_bih[l] : raw.enc_bih[l];
        const float* bhh = dec ? raw.dec_bhh[l] : raw.enc_bhh[l];
        tc->pass_off[p] = off;
        tc_pack_pass_kernel<<<296, 256, 0, st>>>(wih, whh, bih, bhh, xin ? cfg.D : H, H, xin ? 1 : 0,
                                                 reinterpret_cast<unsigned short*>(static_cast<unsigned char*>(tc->wpack) + off),
                                                 tc->bias + (size_t)p * 4 * H);
        SHM_LAUNCH_CHECK();
        off += tc_pass_bytes(H, xin);
    }
    return SHM_OK;
}

void vae_tc_free(VaeTc* tc) {
    if (!tc) return;
    if (tc->wpack) cudaFree(tc->wpack);
    if (tc->bias) cudaFree(tc->bias);
    if (tc->scratch) cudaFree(tc->scratch);
    if (tc->dbg) cudaFree(tc->dbg);
    tc->dbg = nullptr;
    tc->wpack = nullptr; tc->bias = nullptr; tc->scratch = nullptr;
}

bool vae_tc_can_rescore(const VaeTc* tc) { return tc->H == 128 || tc->L > 1; }

int vae_tc_score(VaeTc* tc, const VaeDev& P, const WinSrc& src, const VaeIO& io, cudaStream_t st) {
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    const int H = tc->H, L = tc->L;
    const long long tiles = (io.n + TCM - 1) / TCM;
    int grid = (int)min((long long)device_sm_count(dev), tiles);
    grid = (grid + TC_CLUSTER - 1) / TC_CLUSTER * TC_CLUSTER;           // whole clusters; surplus CTAs run dummy tiles
    const size_t img = (size_t)TCM * H * 2 * 2;
    const size_t stride = (L > 1) ? img * (size_t)src.T : 0;
    const bool hoist = (H == 128 && L > 1);                   // G chunk image (32 units x 128 rows x 4 gates fp32) must equal an input buffer
    const size_t g_bytes = hoist ? (size_t)4 * H * TCM * sizeof(float) : 0;
    const size_t need = (stride + TC_HEADS_SCRATCH + g_bytes) * (size_t)(device_sm_count(dev) + TC_CLUSTER);
    if (need > tc->scratch_bytes) {
        SHM_CUDA(cudaStreamSynchronize(st));
        if (tc->scratch) cudaFree(tc->scratch);
        tc->scratch = nullptr; tc->scratch_bytes = 0;
        if (cudaMalloc(&tc->scratch, need) != cudaSuccess) { set_cuda_error(cudaGetLastError(), "cudaMalloc(vae tc scratch)"); return SHM_ERR_NOMEM; }
        tc->scratch_bytes = need;
    }
    TcDev T;
    memset(&T, 0, sizeof(T));
    T.n_pass = 2 * L; T.L = L;
    T.scratch = static_cast<unsigned char*>(tc->scratch); T.scratch_stride = stride; T.g_bytes = g_bytes;
    T.dbg = tc->dbg;
    for (int p = 0; p < 2 * L; ++p) {
        const bool dec = p >= L;
        const int l = dec ? p - L : p;
        T.pass[p].w = static_cast<const unsigned char*>(tc->wpack) + tc->pass_off[p];
        T.pass[p].bias = tc->bias + (size_t)p * 4 * H;
        T.pass[p].in_kind = (l > 0) ? IN_STREAM : (dec ? (hoist ? IN_HOIST : IN_CONST) : IN_X);
        T.pass[p].sink = (l < L - 1) ? SINK_STREAM : (dec ? SINK_LAST_DEC : SINK_LAST_ENC);
    }
    if (H == 128) vae_score_tc_kernel<128><<<grid, TC_THREADS, TcSmem<128>::total, st>>>(P, T, src, io);
    else if (L == 1) {
        // single-layer H = 64 (the openLAB model): two independent tiles per CTA, see vae_tc_dual.cuh
        const int grid2 = (int)min((long long)device_sm_count(dev), (tiles + 1) / 2);
        vae_score_tc_dual_kernel<<<grid2, TC_THREADS, D2Smem::total, st>>>(P, T, src, io);
    } else vae_score_tc_kernel<64><<<grid, TC_THREADS, TcSmem<64>::total, st>>>(P, T, src, io);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

}  // namespace shm
