// Exact device-side percentile with NumPy's default ("linear") interpolation, reproducing
// `np.percentile(scores_fp32, q)` bit for bit: threshold calibration of
// 4DOF/Scripts/04_vae_thresholding.py:283 (P99) and openLAB 05_validate_vae.py:253 (P95).
//
// NumPy (2.x) evaluates everything in the array's dtype: qf = fp32(q)/fp32(100),
// v = fp32(N-1)*qf, lo = floor(v), g = v - lo, r = a + (b-a)*g, and for g >= 0.5 r = b - (b-a)*(1-g),
// with a, b the lo-th and (lo+1)-th order statistics.  Those two are found by a most-significant-digit
// radix select (8 bits per pass over an order-preserving key), both ranks in the same pass, so the
// scores never leave the device and are never fully sorted.
//
// ONE streaming pass over the scores (r02; the r01 select read them four times), 4 launches:
//  1. gather + sample: a strided sample of 8,192 scores gives a bracket [k_lo, k_hi] around the wanted order statistics (two order
//     statistics of the sample, +-6 sigma of the sample quantile's rank);
//  2. filter: the single full read counts the scores below the bracket, appends the ones inside it (~1.5 % of N) to a candidate
//     buffer and histograms their first digit;
//  3. select: one cooperative kernel runs the remaining digit passes over the candidates (grid barriers in between) and
//     interpolates.  Digits are taken of (key - k_lo), top-aligned to the bracket's width: they are spread over all 256 bins
//     (plain shared-memory atomics, no warp aggregation) and a bracket of 2^21 keys needs 3 digits, not 4.
// The device checks that both ranks fall inside the bracket; if not (adversarial input), the select runs on the full array
// instead -- always exact, never a host round trip.
// NaNs of either sign sort last (np.sort order): their key is 0xffffffff, and the bracket is clamped to [-inf, +inf], so the
// filter can compare in the float domain (two FSETP per score instead of the key transform + two integer compares).
#include "common.cuh"
#include <cooperative_groups.h>

namespace shm {
namespace cg = cooperative_groups;

struct PctState {
    unsigned int hist[4][2][256];     // per pass, per rank (select kernel)
    unsigned int hist_f[256];         // first digit of the candidates (filter kernel; both ranks share the empty prefix)
    unsigned int rank0[2];            // the two wanted ranks in the full array
    unsigned int lo_is_last;
    float g;
    unsigned int klo, khi;            // bracket (order-preserving keys), inclusive, inside [key(-inf), key(+inf)]
    unsigned int n_below;             // scores below the bracket
    unsigned int n_cand;              // scores inside the bracket (appended to the candidate buffer)
    unsigned int sample[8192];        // keys of the strided sample
};

constexpr int PCT_SAMPLE = 8192;
static_assert(PCT_SAMPLE == 1 << 13, "the sample index is a shift");
constexpr int PCT_SELECT_THREADS = 512, PCT_SELECT_PER_SM = 1;   // one CTA per SM: every extra CTA adds 512 same-address global atomics per pass and makes the grid barrier dearer
constexpr float PCT_SIGMAS = 6.0f;    // bracket half-width in standard deviations of the sample quantile's rank (a miss only costs the fallback)
constexpr unsigned int KEY_NEG_INF = 0x007fffffu, KEY_POS_INF = 0xff800000u, KEY_NEG_ZERO = 0x7fffffffu, KEY_POS_ZERO = 0x80000000u;

__device__ __forceinline__ unsigned int f2key(float f) {
    const unsigned int u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;             // NaN: last
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
    const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// digit layout of a select over keys in [base, base + span]: n_pass digits of 8 bits, the last one at bit 0
__device__ __forceinline__ int pct_passes(unsigned int span) { return span == 0 ? 0 : (32 - __clz(span) + 7) >> 3; }
__device__ __forceinline__ int pct_shift(int n_pass, int pass) { return 8 * (n_pass - 1 - pass); }

// One warp resolves one 8-bit digit of a radix select from a 256-bin histogram in global memory (read through L2: the bins were
// accumulated with atomics by other SMs).
__device__ __forceinline__ void pct_digit(const unsigned int* hist, unsigned int& rank, unsigned int& prefix, int shift) {
    const int lane = threadIdx.x & 31;
    unsigned int c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = __ldcg(hist + lane * 8 + j); sum += c[j]; }
    unsigned int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    const unsigned int excl = incl - sum;
    const bool here = rank >= excl && rank < incl;
    const unsigned int vote = __ballot_sync(0xffffffffu, here);
    const int src = vote ? (__ffs(vote) - 1) : 31;
    unsigned int d = 255, rem = 0;
    if (lane == src) {
        unsigned int cum = excl;
        int j = 0;
        for (; j < 8; ++j) { if (rank < cum + c[j]) break; cum += c[j]; }
        if (j > 7) j = 7;
        d = (unsigned int)(lane * 8 + j);
        rem = rank - cum;
    }
    d = __shfl_sync(0xffffffffu, d, src);
    rank = __shfl_sync(0xffffffffu, rem, src);
    prefix |= d << shift;
}

// The strided sample, gathered by 32 CTAs (one SM cannot keep 8,192 scattered DRAM reads in flight: a single-CTA gather took
// 16 us of the 25 us sample kernel); also zeroes the histograms and counters.
__global__ void __launch_bounds__(256) pct_gather_kernel(const float* __restrict__ x, long long N, PctState* s) {
    const int i = blockIdx.x * 256 + threadIdx.x;                       // grid = PCT_SAMPLE / 256
    const int m = (int)(N < PCT_SAMPLE ? N : PCT_SAMPLE);
    s->sample[i] = i < m ? f2key(x[m == PCT_SAMPLE ? (long long)(((unsigned long long)i * (unsigned long long)N) >> 13) : (long long)i])
                         : 0xffffffffu;     // floor(i N / m) in integers
    if (i < 4 * 2 * 256 + 256) (&s->hist[0][0][0])[i] = 0;             // hist + hist_f (adjacent)
    if (i == 0) { s->n_below = 0; s->n_cand = 0; }
}

// ranks + interpolation weight, then the two bracket keys from the sample (8 keys per thread, kept in registers): each is the
// largest K with |{key < K}| <= rank, built bit by bit from the top; one round = 8 compares per thread and rank, a warp reduction,
// one shared-memory atomic per warp and rank (three rotating counters), one block barrier.  The bracket only has to CONTAIN the
// wanted ranks (the filter counts exactly and the select verifies), so the search stops after the top PCT_SAMPLE_BITS bits: the
// lower key is rounded down, the upper one up (2^14 keys = 0.2 % of the value, against a bracket some +-10 % wide).  While the
// two keys agree, one count serves both ranks.  One CTA of 1024 threads; measured alternatives at 32 rounds (this: 18 us): 2 bits per
// round 21 us, 4 fat warps x 64 keys 35 us, a bitonic sort of the sample 54 us.
// Thread 0's scalar prologue is fp32 only (a dependent fp64 divide + sqrt chain cost several us here).
constexpr int PCT_SAMPLE_BITS = 18;
__global__ void __launch_bounds__(1024) pct_sample_kernel(long long N, float q, PctState* s) {
    __shared__ unsigned int cnt[3][2];
    __shared__ int have[2];
    __shared__ unsigned int rk[2];
    const int m = (int)(N < PCT_SAMPLE ? N : PCT_SAMPLE);
    unsigned int key[PCT_SAMPLE / 1024];
#pragma unroll
    for (int j = 0; j < PCT_SAMPLE / 1024; ++j) key[j] = s->sample[j * 1024 + threadIdx.x];
    if (threadIdx.x < 6) (&cnt[0][0])[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        const float qf = __fdiv_rn(q, 100.0f);
        const float v = __fmul_rn((float)(N - 1), qf);
        long long lo;
        float g;
        if (v >= (float)(N - 1)) { lo = N - 1; g = 0.f; s->lo_is_last = 1; }
        else { lo = (long long)floorf(v); g = __fsub_rn(v, (float)lo); s->lo_is_last = 0; }
        s->rank0[0] = (unsigned int)lo;
        s->rank0[1] = (unsigned int)min(lo + 1, N - 1);
        s->g = g;
        // where the wanted ranks sit in the sample: +-1 rank of rounding is inside the margin's "+ 8"
        const float p = N > 1 ? fminf((float)lo / (float)(N - 1), 1.0f) : 0.0f;
        const float r = p * (float)(m - 1);
        const int margin = (int)ceilf(PCT_SIGMAS * sqrtf((float)m * p * (1.0f - p))) + 8;
        const int lo_s = (int)floorf(r) - margin, hi_s = (int)ceilf(r) + 1 + margin;
        have[0] = lo_s > 0; have[1] = hi_s < m - 1;
        rk[0] = have[0] ? (unsigned int)lo_s : 0u; rk[1] = have[1] ? (unsigned int)hi_s : (unsigned int)(m - 1);
    }
    __syncthreads();
    const unsigned int r0 = rk[0], r1 = rk[1];
    unsigned int K0 = 0, K1 = 0;
    for (int b = 31, it = 0; b >= 32 - PCT_SAMPLE_BITS; --b, ++it) {
        const unsigned int t0 = K0 | (1u << b), t1 = K1 | (1u << b);
        const bool same = K0 == K1;                    // block-uniform
        unsigned int c0 = 0, c1 = 0;
#pragma unroll
        for (int j = 0; j < PCT_SAMPLE / 1024; ++j) c0 += key[j] < t0 ? 1u : 0u;
        c0 = __reduce_add_sync(0xffffffffu, c0);
        if (!same) {
#pragma unroll
            for (int j = 0; j < PCT_SAMPLE / 1024; ++j) c1 += key[j] < t1 ? 1u : 0u;
            c1 = __reduce_add_sync(0xffffffffu, c1);
        }
        unsigned int* c = cnt[it % 3];
        if ((threadIdx.x & 31) == 0) { if (c0) atomicAdd(&c[0], c0); if (c1) atomicAdd(&c[1], c1); }
        __syncthreads();
        const unsigned int n0 = c[0], n1 = same ? n0 : c[1];
        if (n0 <= r0) K0 = t0;
        if (n1 <= r1) K1 = t1;
        // the counters of round it+2 were last read before this round's barrier and are next added to after the next one
        if (threadIdx.x < 2) cnt[(it + 2) % 3][threadIdx.x] = 0;
    }
    K1 |= (1u << (32 - PCT_SAMPLE_BITS)) - 1u;         // upper key rounded up, lower key (low bits 0) rounded down
    if (threadIdx.x == 0) {
        // clamp to the finite/infinite range (NaN keys stay outside: "above") and take both zeros when a bound is a zero, so
        // that float compares against key2f(klo), key2f(khi) select exactly the keys in [klo, khi]
        unsigned int klo = have[0] ? K0 : KEY_NEG_INF, khi = have[1] ? K1 : KEY_POS_INF;
        klo = max(klo, KEY_NEG_INF); khi = min(khi, KEY_POS_INF);
        if (klo == KEY_POS_ZERO) klo = KEY_NEG_ZERO;
        if (khi == KEY_NEG_ZERO) khi = KEY_POS_ZERO;
        if (klo > khi) klo = khi;                       // a sample of NaNs only
        s->klo = klo; s->khi = khi;
    }
}

// The one full read: count scores below the bracket, append scores inside it.  Warps are independent (no block barrier in the loop):
// each warp streams 512-score chunks, the next chunk's four 16-byte loads per lane in flight while the current one is compared
// (two float compares per score); the few candidates (~8 per chunk) are re-read through L1/L2 by position, compacted into the
// warp's shared-memory buffer by ballot rounds and flushed with one global atomic when the buffer runs full.  The flush also
// histograms the candidates' first digit (of key - klo, top-aligned to the bracket width: spread over the bins) in shared memory.
constexpr int PCT_STAGE = 1024;           // per-warp staging capacity (flushed when fewer than 512 slots are free)
__global__ void __launch_bounds__(256, 4) pct_filter_kernel(const float* __restrict__ x, long long N, PctState* s, float* __restrict__ cand) {
    __shared__ float stage[8][PCT_STAGE];
    __shared__ unsigned int hf[256];
    __shared__ unsigned int wsum[8];
    const unsigned int klo = s->klo, khi = s->khi;
    const float flo = key2f(klo), fhi = key2f(khi);
    const int n_pass = pct_passes(khi - klo), shift0 = pct_shift(n_pass, 0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    float* st = stage[warp];
    unsigned int below = 0, staged = 0;
    hf[threadIdx.x] = 0;
    __syncthreads();
    auto flush = [&]() {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(&s->n_cand, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned int i = lane; i < staged; i += 32) {
            const float v = st[i];
            cand[base + i] = v;
            if (n_pass > 0) atomicAdd(&hf[((f2key(v) - klo) >> shift0) & 255u], 1u);
        }
        __syncwarp();
        staged = 0;
    };
    // candidates of a chunk: lane element e sits at xc[(e >> 2) * 128 + lane * 4 + (e & 3)]
    auto emit = [&](const float* xc, unsigned int mask16) {
        if (!__any_sync(0xffffffffu, mask16 != 0)) return;
        if (staged > PCT_STAGE - 512) flush();
        for (;;) {
            const bool has = mask16 != 0;
            const unsigned int b = __ballot_sync(0xffffffffu, has);
            if (!b) break;
            if (has) {
                const int e = __ffs(mask16) - 1;
                st[staged + __popc(b & lt_mask)] = xc[(e >> 2) * 128 + lane * 4 + (e & 3)];
                mask16 &= mask16 - 1;
            }
            staged += __popc(b);
        }
        __syncwarp();
    };
    auto cmp4 = [&](const float4 f, int q, unsigned int& mask16) {
        const float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool lt = v[e] < flo;
            below += lt ? 1u : 0u;
            if (v[e] <= fhi && !lt) mask16 |= 1u << (4 * q + e);
        }
    };
    // aligned body in whole chunks; the unaligned head (< 4 scores) and the tail (< 512) go to one warp's scalar loop
    const long long head = min(N, (long long)(((16 - (reinterpret_cast<uintptr_t>(x) & 15)) & 15) >> 2));
    const float* xb = x + head;
    const long long n_chunks = (N - head) / 512;
    const long long wid = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
    if (wid < n_chunks) {
        float4 cur[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) cur[q] = __ldcs(reinterpret_cast<const float4*>(xb + wid * 512 + q * 128 + lane * 4));
        for (long long c = wid; c < n_chunks; c += nw) {
            const long long cn = c + nw < n_chunks ? c + nw : c;          // the last round re-reads its own chunk (L2 hit), unused
            float4 nxt[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) nxt[q] = __ldcs(reinterpret_cast<const float4*>(xb + cn * 512 + q * 128 + lane * 4));
            unsigned int mask16 = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) cmp4(cur[q], q, mask16);
            emit(xb + c * 512, mask16);
#pragma unroll
            for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
        }
    }
    if (wid == nw - 1) {                                                  // the last warp: least loaded by the chunk loop
        const long long t0 = head + n_chunks * 512;
        for (long long b0 = -32; b0 < N - t0; b0 += 32) {                 // round -32: the head
            const long long i = b0 < 0 ? (long long)lane : t0 + b0 + lane;
            const bool ok = b0 < 0 ? (long long)lane < head : i < N;
            const float v = ok ? x[i] : 0.f;
            const bool lt = ok && v < flo;
            const bool in = ok && v <= fhi && !lt;
            below += lt ? 1u : 0u;
            const unsigned int b = __ballot_sync(0xffffffffu, in);
            if (b) {
                if (staged > PCT_STAGE - 512) flush();
                if (in) st[staged + __popc(b & lt_mask)] = v;
                staged += __popc(b);
                __syncwarp();
            }
        }
    }
    if (staged) flush();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) wsum[warp] = below;
    __syncthreads();
    if (hf[threadIdx.x]) atomicAdd(&s->hist_f[threadIdx.x], hf[threadIdx.x]);
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
        for (int i = 0; i < 8; ++i) tot += wsum[i];
        if (tot) atomicAdd(&s->n_below, tot);
    }
}

// Are both wanted ranks inside the bracket?  Then the select runs on the candidates (keys relative to klo, ranks shifted by
// n_below, first digit already histogrammed by the filter); otherwise on the full array (base 0, all four digits).
struct PctSel { bool use_cand; unsigned int n; unsigned int rank[2]; unsigned int base; int n_pass; };
__device__ __forceinline__ PctSel pct_decide(const PctState* s, long long N) {
    PctSel d;
    const unsigned int r0 = s->rank0[0], r1 = s->rank0[1];
    d.use_cand = s->n_below <= r0 && (unsigned long long)r1 < (unsigned long long)s->n_below + s->n_cand;
    d.n = d.use_cand ? s->n_cand : (unsigned int)N;
    d.rank[0] = d.use_cand ? r0 - s->n_below : r0;
    d.rank[1] = d.use_cand ? r1 - s->n_below : r1;
    d.base = d.use_cand ? s->klo : 0u;
    d.n_pass = d.use_cand ? pct_passes(s->khi - s->klo) : 4;
    return d;
}

// The radix select on the candidates (or, in the fallback, on the full array): most-significant-digit first, 8 bits per pass, both
// ranks in the same pass.  ONE cooperative launch: every CTA histograms its share of a pass in shared memory, adds it to the global
// bins, and after the grid barrier resolves the digit itself (all CTAs compute the same prefix and remaining rank); CTA 0
// interpolates the result.  AGG: warp-aggregated atomics for the fallback, whose leading digits (sign + exponent) are concentrated.
template <bool AGG>
__device__ __forceinline__ void pct_count(unsigned int* h, unsigned int bin, int lane) {
    if (AGG) {
        const unsigned int peers = __match_any_sync(0xffffffffu, bin);
        if (bin < 256u && lane == __ffs(peers) - 1) atomicAdd(&h[bin], (unsigned int)__popc(peers));
    } else if (bin < 256u) {
        atomicAdd(&h[bin], 1u);
    }
}

template <bool AGG>
__device__ __forceinline__ void pct_hist_pass(const float* __restrict__ x, long long N, unsigned int base, int shift, unsigned int himask,
                                              unsigned int p0, unsigned int p1, unsigned int (*h)[256]) {
    const bool same = p0 == p1;
    const int lane = threadIdx.x & 31;
    for (long long b0 = (long long)blockIdx.x * (4 * PCT_SELECT_THREADS); b0 < N; b0 += (long long)gridDim.x * (4 * PCT_SELECT_THREADS)) {
        unsigned int k[4];
        bool valid[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {                         // four independent loads in flight per thread
            const long long i = b0 + e * PCT_SELECT_THREADS + threadIdx.x;
            valid[e] = i < N;
            k[e] = valid[e] ? f2key(x[i]) - base : 0u;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int d = (k[e] >> shift) & 255u;
            pct_count<AGG>(h[0], (valid[e] && ((k[e] ^ p0) & himask) == 0) ? d : 256u, lane);
            if (!same) pct_count<AGG>(h[1], (valid[e] && ((k[e] ^ p1) & himask) == 0) ? d : 256u, lane);
        }
    }
}

__global__ void __launch_bounds__(PCT_SELECT_THREADS)
pct_select_kernel(const float* __restrict__ x_full, const float* __restrict__ cand, long long N_full, PctState* s, double* result) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int h[2][256];
    __shared__ unsigned int pre[2], rk[2];
    const PctSel sel = pct_decide(s, N_full);
    const float* __restrict__ x = sel.use_cand ? cand : x_full;
    const long long N = sel.n;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 2) { pre[threadIdx.x] = 0; rk[threadIdx.x] = sel.rank[threadIdx.x]; }
    __syncthreads();
    int pass = 0;
    if (sel.use_cand && sel.n_pass > 0) {                         // first digit: the filter's histogram
        if (threadIdx.x < 64) {
            const int r = threadIdx.x >> 5;
            unsigned int rank = rk[r], prefix = 0;
            pct_digit(s->hist_f, rank, prefix, pct_shift(sel.n_pass, 0));
            if (lane == 0) { rk[r] = rank; pre[r] = prefix; }
        }
        __syncthreads();
        pass = 1;
    }
    for (; pass < sel.n_pass; ++pass) {
        if (threadIdx.x < 256) { h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0; }
        __syncthreads();
        const int shift = pct_shift(sel.n_pass, pass);
        const unsigned int himask = shift + 8 >= 32 ? 0u : (0xffffffffu << (shift + 8));
        const unsigned int p0 = pre[0], p1 = pre[1];
        if (sel.use_cand) pct_hist_pass<false>(x, N, sel.base, shift, himask, p0, p1, h);
        else pct_hist_pass<true>(x, N, sel.base, shift, himask, p0, p1, h);
        __syncthreads();
        if (threadIdx.x < 256) {
            const unsigned int c0 = h[0][threadIdx.x], c1 = p0 == p1 ? c0 : h[1][threadIdx.x];
            if (c0) atomicAdd(&s->hist[pass][0][threadIdx.x], c0);
            if (c1) atomicAdd(&s->hist[pass][1][threadIdx.x], c1);
        }
        grid.sync();
        if (threadIdx.x < 64) {                                   // warp r resolves rank slot r
            const int r = threadIdx.x >> 5;
            unsigned int rank = rk[r], prefix = pre[r];
            pct_digit(s->hist[pass][r], rank, prefix, shift);
            if (lane == 0) { rk[r] = rank; pre[r] = prefix; }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float a = key2f(sel.base + pre[0]), b = key2f(sel.base + pre[1]);
        const float g = s->g;
        const float diff = __fsub_rn(b, a);
        float res = __fadd_rn(a, __fmul_rn(diff, g));
        if (g >= 0.5f) res = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, g)));
        *result = (double)res;
    }
}

}  // namespace shm

static inline size_t pct_state_bytes() { return (sizeof(shm::PctState) + 255) / 256 * 256; }

// state + candidate buffer (worst case -- e.g. a constant array -- every score lies inside the bracket)
extern "C" int64_t shm_percentile_workspace_bytes(int64_t N) {
    return (int64_t)(pct_state_bytes() + (size_t)(N > 0 ? N : 0) * sizeof(float));
}

extern "C" int shm_percentile(const float* scores, int64_t N, double q, double* result, void* workspace, void* stream) {
    using namespace shm;
    if (!scores || !result || !workspace || N <= 0 || N > 0xffffffffLL || !(q >= 0.0 && q <= 100.0)) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PctState* s = static_cast<PctState*>(workspace);
    float* cand = reinterpret_cast<float*>(static_cast<char*>(workspace) + pct_state_bytes());
    pct_gather_kernel<<<PCT_SAMPLE / 256, 256, 0, st>>>(scores, N, s);
    SHM_LAUNCH_CHECK();
    pct_sample_kernel<<<1, 1024, 0, st>>>(N, (float)q, s);
    SHM_LAUNCH_CHECK();
    const int nsm = device_sm_count(dev);
    const int fgrid = (int)min((long long)nsm * 4, (long long)((N + 4095) / 4096));     // one resident wave (4 CTAs per SM)
    pct_filter_kernel<<<fgrid, 256, 0, st>>>(scores, N, s, cand);
    SHM_LAUNCH_CHECK();
    // the select reads a few per cent of N in the common case, so a modest co-resident grid suffices (grid barriers get dearer
    // with the CTA count: 8 per SM measured slower than 4); the grid-stride loop still covers the full array in the fallback
    static int sel_per_sm[64] = {0};
    if (dev < 64 && sel_per_sm[dev] == 0) {
        int nb = 0;
        SHM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pct_select_kernel, PCT_SELECT_THREADS, 0));
        sel_per_sm[dev] = nb < 1 ? 1 : (nb > PCT_SELECT_PER_SM ? PCT_SELECT_PER_SM : nb);
    }
    const int per_sm = dev < 64 ? sel_per_sm[dev] : 1;
    const int grid = (int)min((long long)nsm * per_sm, (long long)((N + 4 * PCT_SELECT_THREADS - 1) / (4 * PCT_SELECT_THREADS)));
    const float* cand_c = cand;
    long long n_ll = (long long)N;
    void* args[] = {(void*)&scores, (void*)&cand_c, (void*)&n_ll, (void*)&s, (void*)&result};
    SHM_CUDA(cudaLaunchCooperativeKernel((const void*)pct_select_kernel, dim3(grid), dim3(PCT_SELECT_THREADS), args, 0, st));
    return SHM_OK;
}
