// Two-tile variant of the tensor-core scorer for the single-layer H = 64 model (openLAB VAE, temporal_vae_model.py:4-66).
// Included by vae_tc.cu (uses its helpers); selected by vae_tc_score when H == 64 and L == 1.
//
// With H = 64 a step is only two 32-unit chunks, the cell update (MUFU/issue bound) is longer than the MMAs, and the
// recurrence serialises them: the single-tile kernel leaves the tensor pipe AND the epilogue warps idle in turn.  Here one
// CTA owns TWO independent 128-window tiles.  TMEM holds, per tile, ONE 128-column accumulator and the double-buffered h_t
// (2 x 64 columns): 2 x (128 + 128) = 512 columns.  The MMA issuer walks (chunk, tile) = (c0,A) (c0,B) (c1,A) (c1,B); the
// epilogue warps follow in the same order, so tile B's MMAs run under tile A's cell update and vice versa.  The whole
// weight set of a pass (80 KB encoder / 128 KB decoder) is resident in shared memory -- no ring, no L2 latency on the
// recurrence -- and is loaded once per pass for 256 windows.  The decoder's xhat_{t-1} accumulates in the dead half of the
// h double buffer (d2_xhat), so neither the issuer nor the accumulators ever wait for the output warps.
#pragma once

namespace shm {

constexpr int D2_H = 64, D2_NCH = 2, D2_NT = 2;
constexpr int D2_IMG = 2 * TCM * D2_H * 2;                 // hi|lo fp16 image [128 x 64]: 32 KB
constexpr int D2_XT = 4 * TC_XSTAGE;                       // per tile: double-buffered hi|lo window tiles (16 KB)
constexpr int D2_WOIMG = 16 * D2_H * 2;
constexpr int D2_WENC = D2_NCH * (2 * TC_XSTAGE + 2 * TC_STAGE);      // resident encoder weights: 80 KB
constexpr int D2_WDEC = D2_NCH * 4 * TC_STAGE;                        // resident decoder weights: 128 KB

// Shared memory in 32 KB units.  Units 0-3: the pass's WHOLE weight set, resident (encoder 80 KB, decoder 128 KB), loaded
// once per tile pair and pass.  Units 4-5: encoder = the two tiles' window tiles (unit 4) and tile A's fp32 h_T (unit 5);
// tile B's h_T sits in unit 3 (free while the encoder weights are resident); heads: u_A -> unit 4, u_B -> unit 5 (tile A's
// h_T is dead by then); decoder = u_A | u_B.
struct D2Smem {
    static constexpr int off_w = 0;
    static constexpr int off_in = 4 * D2_IMG;
    static constexpr int off_wo = off_in + D2_NT * D2_IMG;
    static constexpr int off_bias = off_wo + 2 * D2_WOIMG + 64;
    static constexpr int off_bar = off_bias + D2_H * 4 * 4;
    static constexpr int total = off_bar + 48 * 8 + 16;
    static_assert(total <= 232448, "shared memory budget");
};

struct D2Bars {
    uint64_t w_full[D2_NCH];
    uint64_t in_full[D2_NT][2], in_empty[D2_NT][2];
    uint64_t acc_full[D2_NT], acc_empty[D2_NT], h_full[D2_NT], xhat_full[D2_NT], xhat_empty[D2_NT];      // h_full: once per step
};
static_assert(sizeof(D2Bars) <= 48 * 8, "barrier block");

struct D2Ctx {
    D2Bars* bars;
    unsigned char* smem0;          // unit 0 (resident weights; heads scratch)
    unsigned char* inbuf;          // unit 4
    float* bias_s;
    unsigned char* wo_img;
    float* bo_s;
    uint32_t tbase;
    int T;
    long long n0[D2_NT];
    int nvalid[D2_NT];
};

__device__ __forceinline__ uint32_t d2_acc(uint32_t tbase, int tl) { return tbase + (uint32_t)(tl * 128); }
__device__ __forceinline__ uint32_t d2_h(uint32_t tbase, int tl, int buf) { return tbase + 256u + (uint32_t)(tl * 128 + buf * D2_H); }
// xhat_{t-1} = h_{t-1} W_o^T (16 fp32 columns) is accumulated in columns [16,32) of the h buffer that step t will overwrite
// (h_{t-2}, dead): only the chunk-1 cell update of step t, which stores there, has to wait until the output warps have read it.
__device__ __forceinline__ uint32_t d2_xhat(uint32_t tbase, int tl, int tprev) { return d2_h(tbase, tl, (tprev + 1) & 1) + 16u; }
__device__ __forceinline__ int d2_hT_unit(int tl) { return tl == 0 ? 5 : 3; }

// ------------------------------------------------------------------------------------------------ epilogue warps (16)
// All 16 warps walk the virtual chunks v = (chunk, tile) = (c0,A) (c0,B) (c1,A) (c1,B) together: while they update tile A's
// cells the tensor pipe produces tile B's gates and vice versa.  A thread owns one window row and 8 units of each chunk.
// (Measured alternatives, profiles/experiments: one 8-warp group per tile running out of phase, with one or two issuer
// warps and per-K-block h barriers, were 6-10 % slower: the groups fall into phase through the shared tensor pipe.)
// dec == false: encoder (keeps the fp32 h_T of the last step per tile); dec == true: decoder (h_t only feeds TMEM).
// D2_EPI_UNROLL = 1: the (chunk, tile) loop stays rolled and the cell state rotates through the register arrays (44 of the loop's
// 316 instructions are those moves); 4: fully unrolled, the rotation is register renaming, the body is 4x the code.  Measured equal
// (8.20 vs 8.18 ms per 151,552 windows; 10.91 vs 10.89 M windows/s at 2^20): the pass is bound by the XU (56 MUFU per chunk and warp),
// not by issue slots, so the rolled form stays.
#ifndef D2_EPI_UNROLL
#define D2_EPI_UNROLL 1
#endif
constexpr int kD2EpiUnroll = D2_EPI_UNROLL;
template <bool DEC>
__device__ __forceinline__ void d2_epi_pass(const D2Ctx& cx, const float* bias_g, uint32_t (&n_acc)[D2_NT], uint32_t (&n_xe)[D2_NT],
                                            long long (&prof)[8]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wg = warp >> 2;                                  // units [8*wg, 8*wg+8) of each 32-unit chunk
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // gate multipliers 2^bias (lstm_cell_bmul2), stored per unit PAIR as {Bi(u),Bi(u+1)} {Bf..} {Bg..} {Bo..}
    for (int i = tid; i < D2_H * 4; i += TC_EPI_THREADS) {
        const int unit = i >> 2, gate = i & 3;
        cx.bias_s[(unit >> 1) * 8 + gate * 2 + (unit & 1)] = exp2f(fminf(__ldg(bias_g + i), 20.f));
    }
    epi_bar_sync();
    float cst[D2_NCH * D2_NT][TC_UPT];                         // cell state per virtual chunk, rotated so the next is cst[0]
#pragma unroll
    for (int v = 0; v < D2_NCH * D2_NT; ++v)
#pragma unroll
        for (int u = 0; u < TC_UPT; ++u) cst[v][u] = 0.f;
    const int T = cx.T;
    const uint32_t pa0 = n_acc[0], pa1 = n_acc[1], px0 = n_xe[0], px1 = n_xe[1];     // barrier phases at the start of the pass
    for (int t = 0; t < T; ++t) {
#pragma unroll kD2EpiUnroll
        for (int v = 0; v < D2_NCH * D2_NT; ++v) {
            const int c = v >> 1, tl = v & 1;
            const long long q0 = TC_CLOCK();
            mbar_wait(&cx.bars->acc_full[tl], ((tl ? pa1 : pa0) + (uint32_t)c) & 1);     // use 2t + c of the pass: no counter in local memory
            const long long q1 = TC_CLOCK();
            tc_fence_after_sync();
            const int u0 = c * 32 + wg * 8;
            uint32_t g0[8], g1[8], g2[8], g3[8];
            const uint32_t abase = d2_acc(cx.tbase, tl) + lane_base + (uint32_t)(wg * 8);
            tmem_ld8(abase + 0, g0);
            tmem_ld8(abase + 32, g1);
            tmem_ld8(abase + 64, g2);
            tmem_ld8(abase + 96, g3);
            tmem_ld_wait();
            const long long q2 = TC_CLOCK();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&cx.bars->acc_empty[tl]);
            float hv[8];
#pragma unroll
            for (int u = 0; u < 8; u += 2) {
                const ulonglong2 b01 = *reinterpret_cast<const ulonglong2*>(cx.bias_s + (u0 + u) * 4);
                const ulonglong2 b23 = *reinterpret_cast<const ulonglong2*>(cx.bias_s + (u0 + u) * 4 + 4);
                const f32x2 B4[4] = {b01.x, b01.y, b23.x, b23.y};
                lstm_cell_bmul2(__uint_as_float(g0[u]), __uint_as_float(g0[u + 1]), __uint_as_float(g1[u]), __uint_as_float(g1[u + 1]),
                                __uint_as_float(g2[u]), __uint_as_float(g2[u + 1]), __uint_as_float(g3[u]), __uint_as_float(g3[u + 1]), B4,
                                cst[0][u], cst[0][u + 1], hv[u], hv[u + 1]);
            }
            const long long q3 = TC_CLOCK();
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_f16x2(hv[2 * j], hv[2 * j + 1], hi[j], lo[j]);
            const uint32_t hbuf = d2_h(cx.tbase, tl, t & 1) + lane_base;
            if (DEC && c == 1 && t > 0) {                       // columns [16,32) still carry xhat_{t-1} until it has been read
                mbar_wait(&cx.bars->xhat_empty[tl], ((tl ? px1 : px0) + (uint32_t)(t - 1)) & 1);
                tc_fence_after_sync();
            }
            tmem_st4(hbuf + (uint32_t)(u0 >> 1), hi);
            tmem_st4(hbuf + (uint32_t)(D2_H / 2 + (u0 >> 1)), lo);
            if (!DEC && t == T - 1) {
                float* hT = reinterpret_cast<float*>(cx.smem0 + d2_hT_unit(tl) * D2_IMG);      // fp32 [64][128]
#pragma unroll
                for (int u = 0; u < 8; ++u) hT[(u0 + u) * TCM + row] = hv[u];
            }
            if (c == D2_NCH - 1) {                              // this tile's h_t is complete
                tmem_st_wait();
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&cx.bars->h_full[tl]);
            }
            const long long q4 = TC_CLOCK();
            prof[0] += q1 - q0; prof[1] += q2 - q1; prof[2] += q3 - q2; prof[3] += q4 - q3;
            // rotate the cell state: the next (chunk, tile) moves to cst[0]
#pragma unroll
            for (int u = 0; u < TC_UPT; ++u) {
                const float t0 = cst[0][u];
                cst[0][u] = cst[1][u]; cst[1][u] = cst[2][u]; cst[2][u] = cst[3][u]; cst[3][u] = t0;
            }
        }
    }
    n_acc[0] += (uint32_t)(D2_NCH * T); n_acc[1] += (uint32_t)(D2_NCH * T);
    if (DEC) { n_xe[0] += (uint32_t)T; n_xe[1] += (uint32_t)T; }       // T - 1 waits + xhat_{T-1}: read before the pass-end barrier, never waited for here
}

// ------------------------------------------------------------------------------------------------ MMA issuer warp
// The pass's weights are resident, so the issuer only follows the data, in the epilogue's order (c0,A) (c0,B) (c1,A) (c1,B):
//   [accumulator drained (+ window tile staged)] -> input part  ->  [h_{t-1} complete, chunk 0] -> recurrent part, commit
//   (-> xhat_{t-1}, decoder chunk 0).
// The input part does not depend on h_{t-1}, so it is issued while the epilogue is still busy with the tile's previous
// chunk.  The control code between MMAs is kept minimal -- next to 16 busy epilogue warps a single warp's scalar code runs at
// a fraction of an instruction per cycle: descriptors are a base word plus a constant (one uniform add per operand, the
// high word is constant), and every barrier parity is a function of (t, c) and two per-pass bases (no counters in local
// memory, no polling).
__device__ __forceinline__ uint32_t d2_lo(uint64_t desc) { return (uint32_t)desc; }
__device__ __forceinline__ uint64_t d2_k(uint32_t lo) { return (TC_DESC_HI & 0xFFFFFFFF00000000ull) | lo; }
__device__ __forceinline__ uint64_t d2_o(uint32_t lo) { return (make_smem_desc(0, 256, 128) & 0xFFFFFFFF00000000ull) | lo; }

// n_w: passes seen; hb: completed phases of h_full[*] (steps) and ib[b]: of in_full[*][b] before this pass (equal for both
// tiles).  mp[]: cycles waiting for [0] weights, [1] h_full, [2] acc_empty / in_full.
template <bool DEC>
__device__ __forceinline__ void d2_mma_pass(const D2Ctx& cx, uint32_t& n_w, uint32_t& hb, uint32_t (&ib)[2], long long (&mp)[8]) {
    D2Bars* bars = cx.bars;
    constexpr int PB = DEC ? 4 : 0;
    constexpr int IN_BYTES = DEC ? TC_STAGE : TC_XSTAGE;              // one input-part image (hi or lo) of a chunk
    constexpr int CHUNK_BYTES = 2 * IN_BYTES + 2 * TC_STAGE;
    const uint32_t w_lo = d2_lo(kdesc(smem_u32(cx.smem0)));
    const uint32_t in_lo = d2_lo(kdesc(smem_u32(cx.inbuf)));
    const uint32_t wo_lo = d2_lo(make_smem_desc(smem_u32(cx.wo_img), 256, 128));
    const uint32_t tbase = cx.tbase;
    constexpr uint32_t IDESC_X = make_idesc_bf16(128, 128), IDESC_H = make_idesc_f16(128, 128), IDESC16 = make_idesc_f16(128, 16);
    const int T = cx.T;
    auto xhat = [&](int tl, int tprev) {       // xhat_{tprev} = h_{tprev} W_o^T, see d2_xhat
        if (elect_one()) {
            const uint32_t h_hi = d2_h(tbase, tl, tprev & 1), h_lo = h_hi + D2_H / 2;
            const uint32_t d = d2_xhat(tbase, tl, tprev);
#pragma unroll
            for (int k = 0; k < D2_H / 16; ++k) {
                mma_ts(d, h_hi + k * 8, d2_o(wo_lo + k * 32), IDESC16, k > 0 ? 1u : 0u);
                mma_ts(d, h_lo + k * 8, d2_o(wo_lo + k * 32), IDESC16, 1u);
                mma_ts(d, h_hi + k * 8, d2_o(wo_lo + D2_WOIMG / 16 + k * 32), IDESC16, 1u);
            }
            mma_commit(&bars->xhat_full[tl]);
        }
        __syncwarp();
    };
    {                                                       // the pass's weights have landed (once per tile pair)
        const long long m0 = TC_CLOCK();
        for (int c = 0; c < D2_NCH; ++c) mbar_wait(&bars->w_full[c], n_w & 1);
        ++n_w;
        tc_fence_after_sync();
        mp[PB + 0] += TC_CLOCK() - m0;
    }
    for (int t = 0; t < T; ++t) {
        const uint32_t hpar = (hb + (uint32_t)(t - 1)) & 1;                            // h_full[*] of step t-1
        const uint32_t ipar = (ib[t & 1] + (uint32_t)(t >> 1)) & 1;                    // in_full[*][t&1]
#pragma unroll 1
        for (int c = 0; c < D2_NCH; ++c) {                  // rolled: the issuer's code shares the instruction cache with the epilogue loop
            const uint32_t wc = w_lo + c * (CHUNK_BYTES / 16);
#pragma unroll 1
            for (int tl = 0; tl < D2_NT; ++tl) {
                const uint32_t acc = d2_acc(tbase, tl);
                // input part: needs the drained accumulator (use 2t+c of acc_empty) and, encoder chunk 0, the staged window tile
                long long m0 = TC_CLOCK();
                mbar_wait(&bars->acc_empty[tl], (uint32_t)(c ^ 1));
                if (!DEC && c == 0) mbar_wait(&bars->in_full[tl][t & 1], ipar);
                tc_fence_after_sync();
                { const long long m1 = TC_CLOCK(); mp[PB + 2] += m1 - m0; m0 = m1; }
                if (elect_one()) {
                    if (!DEC) {
                        const uint32_t a_hi = in_lo + tl * (D2_XT / 16) + (t & 1) * (2 * TC_XSTAGE / 16), a_lo = a_hi + TC_XSTAGE / 16;
                        mma_ss(acc, d2_k(a_hi), d2_k(wc), IDESC_X, 0u);
                        mma_ss(acc, d2_k(a_lo), d2_k(wc), IDESC_X, 1u);
                        mma_ss(acc, d2_k(a_hi), d2_k(wc + TC_XSTAGE / 16), IDESC_X, 1u);
                    } else {
                        const uint32_t a_hi = in_lo + tl * (D2_IMG / 16), a_lo = a_hi + D2_IMG / 32;         // u image of the tile
                        const uint32_t bh = wc, bl = wc + TC_STAGE / 16;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_ss(acc, d2_k(a_hi + k * 256), d2_k(bh + k * 256), IDESC_H, k > 0 ? 1u : 0u);
                            mma_ss(acc, d2_k(a_lo + k * 256), d2_k(bh + k * 256), IDESC_H, 1u);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ss(acc, d2_k(a_hi + k * 256), d2_k(bl + k * 256), IDESC_H, 1u);
                    }
                }
                __syncwarp();
                // recurrent part: chunk 0 needs the tile's complete h_{t-1}; chunk 1 follows its input part directly
                if (c == 0 && t > 0) {
                    mbar_wait(&bars->h_full[tl], hpar);
                    tc_fence_after_sync();
                    mp[PB + 1] += TC_CLOCK() - m0;
                }
                if (elect_one()) {
                    if (t > 0) {                            // A = h_{t-1} (hi|lo) from TMEM
                        const uint32_t h_hi = d2_h(tbase, tl, (t - 1) & 1), h_lo = h_hi + D2_H / 2;
                        const uint32_t bh = wc + 2 * IN_BYTES / 16, bl = bh + TC_STAGE / 16;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            mma_ts(acc, h_hi + k * 8, d2_k(bh + k * 256), IDESC_H, 1u);
                            mma_ts(acc, h_lo + k * 8, d2_k(bh + k * 256), IDESC_H, 1u);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) mma_ts(acc, h_hi + k * 8, d2_k(bl + k * 256), IDESC_H, 1u);
                    }
                    mma_commit(&bars->acc_full[tl]);
                    if (!DEC && c == D2_NCH - 1) mma_commit(&bars->in_empty[tl][t & 1]);     // window tile fully consumed
                }
                __syncwarp();
                if (DEC && c == 0 && t > 0) xhat(tl, t - 1);        // behind the gates' MMAs: only chunk 1's stores wait for its read
            }
        }
    }
    // the last step's h_{T-1}: phase bookkeeping, and the decoder's xhat_{T-1}
    for (int tl = 0; tl < D2_NT; ++tl) {
        mbar_wait(&bars->h_full[tl], (hb + (uint32_t)(T - 1)) & 1);
        if (DEC) {
            tc_fence_after_sync();
            xhat(tl, T - 1);
        }
    }
    hb += (uint32_t)T;
    if (!DEC) { ib[0] += (uint32_t)((T + 1) >> 1); ib[1] += (uint32_t)(T >> 1); }
}

// ------------------------------------------------------------------------------------------------ copy producer (one lane)
// Loads the pass's whole weight set once (16 KB bulk copies, all in flight together); the barrier of chunk c completes when
// its 40 KB (encoder) / 64 KB (decoder) have landed.  Runs after the pass-start barrier, i.e. after every MMA and every
// generic-proxy access (heads scratch, tile B's h_T) of the previous pass to units 0-3.
template <bool DEC>
__device__ __forceinline__ void d2_prod_pass(const D2Ctx& cx, const unsigned char* w) {
    D2Bars* bars = cx.bars;
    constexpr int IN_BYTES = DEC ? TC_STAGE : TC_XSTAGE;
    constexpr int CHUNK_BYTES = 2 * IN_BYTES + 2 * TC_STAGE;
    for (int c = 0; c < D2_NCH; ++c) {
        mbar_arrive_expect_tx(&bars->w_full[c], CHUNK_BYTES);
        for (int o = 0; o < CHUNK_BYTES; o += TC_STAGE) {
            const int bytes = CHUNK_BYTES - o < TC_STAGE ? CHUNK_BYTES - o : TC_STAGE;
            bulk_g2s(cx.smem0 + c * CHUNK_BYTES + o, w + (size_t)c * CHUNK_BYTES + o, bytes, &bars->w_full[c]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ staging / output group (4 warps)
__device__ __forceinline__ void d2_aux_stage_pass(const D2Ctx& cx, const VaeDev& P, const WinSrc& src, const VaeIO& io,
                                                  uint32_t (&n_inE)[D2_NT][2]) {
    D2Bars* bars = cx.bars;
    const int row = threadIdx.x - TC_WARP_AUX0 * 32;
    const float* wbase[D2_NT];
    bool ok[D2_NT];
#pragma unroll
    for (int tl = 0; tl < D2_NT; ++tl) {
        ok[tl] = row < cx.nvalid[tl];
        const long long n = cx.n0[tl] + (ok[tl] ? row : 0);
        const long long win = (ok[tl] && io.idx) ? (long long)io.idx[n] : n;
        wbase[tl] = src.base + win * src.win_stride;
    }
    for (int t = 0; t < cx.T; ++t) {
        const int b = t & 1;
#pragma unroll
        for (int tl = 0; tl < D2_NT; ++tl) {
            float v[16];
#pragma unroll
            for (int d = 0; d < 16; ++d)
                v[d] = (ok[tl] && d < P.D) ? win_transform_fast(src, __ldg(wbase[tl] + (long long)t * src.row_stride + src.chan[d]), d) : 0.f;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
            mbar_wait(&bars->in_empty[tl][b], (n_inE[tl][b] & 1) ^ 1);
            ++n_inE[tl][b];
            unsigned char* xhi = cx.inbuf + tl * D2_XT + b * 2 * TC_XSTAGE;
            unsigned char* xlo = xhi + TC_XSTAGE;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int off = (q * 16 + (row >> 3)) * 128 + (row & 7) * 16;
                *reinterpret_cast<uint4*>(xhi + off) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                *reinterpret_cast<uint4*>(xlo + off) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
            fence_proxy_async_smem();
            aux_bar_sync();
            if (row == 0) mbar_arrive(&bars->in_full[tl][b]);
        }
    }
}

__device__ __forceinline__ void d2_aux_out_pass(const D2Ctx& cx, const VaeDev& P, const WinSrc& src, const VaeIO& io,
                                                uint32_t (&n_x)[D2_NT]) {
    D2Bars* bars = cx.bars;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = threadIdx.x - TC_WARP_AUX0 * 32;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int T = cx.T;
    const bool exact = io.cnn_in != nullptr || io.recon != nullptr;
    const float* wbase[D2_NT];
    bool ok[D2_NT];
    long long nn[D2_NT];
    float sse[D2_NT];
#pragma unroll
    for (int tl = 0; tl < D2_NT; ++tl) {
        ok[tl] = row < cx.nvalid[tl];
        nn[tl] = cx.n0[tl] + (ok[tl] ? row : 0);
        const long long win = (ok[tl] && io.idx) ? (long long)io.idx[nn[tl]] : nn[tl];
        wbase[tl] = src.base + win * src.win_stride;
        sse[tl] = 0.f;
    }
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int tl = 0; tl < D2_NT; ++tl) {
            float x[SHM_MAX_D];
#pragma unroll
            for (int d = 0; d < SHM_MAX_D; ++d)
                x[d] = (ok[tl] && d < P.D) ? ld_nc_volatile(wbase[tl] + (long long)t * src.row_stride + src.chan[d]) : 0.f;
            mbar_wait(&bars->xhat_full[tl], n_x[tl] & 1);
            ++n_x[tl];
            tc_fence_after_sync();
            uint32_t acc[16];
            tmem_ld16(d2_xhat(cx.tbase, tl, t) + lane_off, acc);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->xhat_empty[tl]);
#pragma unroll
            for (int d = 0; d < SHM_MAX_D; ++d) {
                if (d < P.D) {
                    const float y = __uint_as_float(acc[d]) + cx.bo_s[d];
                    const float xv = !ok[tl] ? 0.f : (exact ? win_transform(src, x[d], d) : win_transform_fast(src, x[d], d));
                    const float e = xv - y;
                    sse[tl] = fmaf(e, e, sse[tl]);
                    if (ok[tl]) {
                        if (io.recon) io.recon[(nn[tl] * T + t) * P.D + d] = y;
                        if (io.cnn_in) {
                            io.cnn_in[((nn[tl] * 2 + 0) * T + t) * P.D + d] = xv;
                            io.cnn_in[((nn[tl] * 2 + 1) * T + t) * P.D + d] = e * e;
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int tl = 0; tl < D2_NT; ++tl)
        if (ok[tl] && io.score) io.score[nn[tl]] = sse[tl] / (float)(T * P.D);
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
vae_score_tc_dual_kernel(VaeDev P, TcDev TC, WinSrc src, VaeIO io) {
    using S = D2Smem;
    extern __shared__ __align__(1024) unsigned char smem[];
    D2Bars* bars = reinterpret_cast<D2Bars*>(smem + S::off_bar);
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(smem + S::off_bar + 48 * 8);
    unsigned short* wo_hi = reinterpret_cast<unsigned short*>(smem + S::off_wo);
    unsigned short* wo_lo = wo_hi + 16 * D2_H;
    float* bo_s = reinterpret_cast<float*>(smem + S::off_wo + 2 * D2_WOIMG);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = src.T;
    long long n_eff = io.n;
    if (io.n_dev) n_eff = min(n_eff, (long long)__ldg(io.n_dev));
    const int n_tiles = (int)((n_eff + TCM - 1) / TCM);
    const int n_pairs = (n_tiles + 1) / 2;

    if (tid == 0) {
        for (int i = 0; i < D2_NCH; ++i) mbar_init(&bars->w_full[i], 1);
        for (int tl = 0; tl < D2_NT; ++tl) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(&bars->in_full[tl][i], 1); mbar_init(&bars->in_empty[tl][i], 1);
                
            }
            mbar_init(&bars->h_full[tl], TC_EPI_WARPS);
            mbar_init(&bars->acc_full[tl], 1); mbar_init(&bars->acc_empty[tl], TC_EPI_WARPS);
            mbar_init(&bars->xhat_full[tl], 1); mbar_init(&bars->xhat_empty[tl], 4);
        }
        fence_mbar_init();
    }
    if (warp == TC_WARP_MMA) tmem_alloc(tmem_holder, 512);
    for (int i = tid; i < 16 * D2_H; i += TC_THREADS) {        // W_o [D,H] -> fp16 hi|lo K-major images, rows d >= D zero
        const int d = i / D2_H, k = i - d * D2_H;
        const float w = d < P.D ? __ldg(P.out_w + d * D2_H + k) : 0.f;
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
    tc_fence_after_sync();

    D2Ctx cx;
    cx.bars = bars; cx.smem0 = smem + S::off_w; cx.inbuf = smem + S::off_in;
    cx.bias_s = reinterpret_cast<float*>(smem + S::off_bias);
    cx.wo_img = smem + S::off_wo; cx.bo_s = bo_s;
    cx.tbase = *tmem_holder; cx.T = T;
    const bool encode_only_call = !io.score && !io.recon && !io.cnn_in;

    // every role keeps its own running use-counters of the barriers it waits on (all roles walk the same sequence)
    auto set_pair = [&](int pair) {
#pragma unroll
        for (int tl = 0; tl < D2_NT; ++tl) {
            const int tile = 2 * pair + tl;
            cx.n0[tl] = tile < n_tiles ? (long long)tile * TCM : 0;
            cx.nvalid[tl] = tile < n_tiles ? (int)min((long long)TCM, n_eff - cx.n0[tl]) : 0;
        }
    };
    if (warp < TC_EPI_WARPS) {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REG_EPI));
        uint32_t e_acc[D2_NT] = {0, 0}, e_xe[D2_NT] = {0, 0};
        long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            set_pair(pair);
            cta_sync();                                    // (A) encoder pass
            { const long long p0 = TC_CLOCK(); d2_epi_pass<false>(cx, TC.pass[0].bias, e_acc, e_xe, prof); prof[4] += TC_CLOCK() - p0; }
            cta_sync();                                    // (B)
            for (int tl = 0; tl < D2_NT; ++tl) {           // heads, one tile after the other
                PassCtx pc;
                pc.ring = cx.smem0; pc.inbuf = cx.smem0;           // buffers addressed in 32 KB units from unit 0
                pc.heads_scratch = reinterpret_cast<float*>(cx.smem0);          // mu | logvar | z scratch: unit 0 (the encoder weights are dead)
                pc.hT_buf = d2_hT_unit(tl); pc.u_buf = 4 + tl;
                pc.nvalid = cx.nvalid[tl]; pc.n0 = cx.n0[tl];
                heads_stage<D2_H>(pc, P, io);
                epi_bar_sync();
            }
            cta_sync();                                    // (C) decoder pass
            if (encode_only_call) continue;
            { const long long p0 = TC_CLOCK(); d2_epi_pass<true>(cx, TC.pass[1].bias, e_acc, e_xe, prof); prof[5] += TC_CLOCK() - p0; }
            cta_sync();                                    // (D)
        }
        if (TC.dbg && tid == 0) for (int i = 0; i < 8; ++i) TC.dbg[(blockIdx.x * 3 + 2) * 8 + i] = prof[i];
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REG_AUX));
        uint32_t m_w = 0;
        uint32_t m_hb = 0, m_ib[2] = {0, 0};                                                                               // issuer
        long long mprof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t a_in[D2_NT][2] = {{0, 0}, {0, 0}}, a_x[D2_NT] = {0, 0};                                                              // staging / output
        for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
            set_pair(pair);
            cta_sync();                                    // (A)
            if (warp == TC_WARP_PROD) { if (lane == 0) d2_prod_pass<false>(cx, TC.pass[0].w); }
            else if (warp == TC_WARP_MMA) d2_mma_pass<false>(cx, m_w, m_hb, m_ib, mprof);
            else if (warp >= TC_WARP_AUX0 && warp < TC_WARP_AUX0 + 4) d2_aux_stage_pass(cx, P, src, io, a_in);
            __syncwarp();
            cta_sync();                                    // (B)
            cta_sync();                                    // (C)
            if (encode_only_call) continue;
            if (warp == TC_WARP_PROD) { if (lane == 0) d2_prod_pass<true>(cx, TC.pass[1].w); }
            else if (warp == TC_WARP_MMA) d2_mma_pass<true>(cx, m_w, m_hb, m_ib, mprof);
            else if (warp >= TC_WARP_AUX0 && warp < TC_WARP_AUX0 + 4) d2_aux_out_pass(cx, P, src, io, a_x);
            __syncwarp();
            cta_sync();                                    // (D)
        }
        if (TC.dbg && warp == TC_WARP_MMA && lane == 0) for (int i = 0; i < 8; ++i) TC.dbg[(blockIdx.x * 3 + 1) * 8 + i] = mprof[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == TC_WARP_MMA) tmem_dealloc(cx.tbase, 512);
}

}  // namespace shm
