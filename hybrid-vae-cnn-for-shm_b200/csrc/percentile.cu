// Exact device-side percentile with NumPy's default ("linear") interpolation, reproducing
// `np.percentile(scores_fp32, q)` bit for bit: threshold calibration of
// 4DOF/Scripts/04_vae_thresholding.py:283 (P99) and openLAB 05_validate_vae.py:253 (P95).
//
// NumPy (2.x) evaluates everything in the array's dtype: qf = fp32(q)/fp32(100),
// v = fp32(N-1)*qf, lo = floor(v), g = v - lo, r = a + (b-a)*g, and for g >= 0.5 r = b - (b-a)*(1-g),
// with a, b the lo-th and (lo+1)-th order statistics.  Those two are found by a most-significant-digit
// radix select (4 passes of 8 bits over an order-preserving key), both ranks in the same pass, so the
// scores never leave the device and are never fully sorted.
//
// ONE streaming pass over the scores (r02; the r01 select read them four times): a strided sample of 8,192 scores gives a
// bracket [k_lo, k_hi] around the wanted order statistics (two order statistics of the sample, +-8 sigma of the sample
// quantile's rank, found by a radix select in shared memory); the single full read then only counts the scores below the
// bracket and appends the ones inside it (~2 % of N) to a candidate buffer, and the radix select runs on the candidates,
// skipping the leading digits the bracket already fixes.  The device checks that both ranks fall inside the bracket; if not
// (adversarial input), the same select kernels run on the full array instead -- always exact, never a host round trip.
// 7 launches: sample, filter, 4 histogram passes (each starts by resolving the previous digits), result.
#include "common.cuh"

namespace shm {

struct PctState {
    unsigned int hist[4][2][256];     // per pass, per rank
    unsigned int rank0[2];            // the two wanted ranks in the full array
    unsigned int lo_is_last;
    float g;
    unsigned int klo, khi;            // bracket (order-preserving keys), inclusive
    unsigned int n_below;             // scores with key < klo
    unsigned int n_cand;              // scores inside the bracket (appended to the candidate buffer)
};

constexpr int PCT_SAMPLE = 8192;
static_assert(PCT_SAMPLE == 1 << 13, "the sample index is a shift");

__device__ __forceinline__ unsigned int f2key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
    const unsigned int u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// One warp resolves one 8-bit digit of a radix select from a 256-bin histogram (shared or global memory).
__device__ __forceinline__ void pct_digit(const unsigned int* hist, unsigned int& rank, unsigned int& prefix, int shift) {
    const int lane = threadIdx.x & 31;
    unsigned int c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = hist[lane * 8 + j]; sum += c[j]; }
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

// ranks + interpolation weight, zeroed histograms, then: strided sample -> two order statistics of the sample by a radix select in
// shared memory (both ranks per pass, warp-aggregated atomics) -> bracket keys.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) pct_sample_kernel(const float* __restrict__ x, long long N, float q, PctState* s) {
    __shared__ unsigned int keys[PCT_SAMPLE];
    __shared__ unsigned int h[2][256];
    __shared__ unsigned int pre[2], rk[2];
    __shared__ int have[2];
    for (int i = threadIdx.x; i < 4 * 2 * 256; i += 1024) (&s->hist[0][0][0])[i] = 0;
    const int m = (int)(N < PCT_SAMPLE ? N : PCT_SAMPLE);
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
        s->n_below = 0; s->n_cand = 0;
        const double p = N > 1 ? (double)lo / (double)(N - 1) : 0.0;
        const double r = p * (double)(m - 1);
        const int margin = (int)ceil(8.0 * sqrt((double)m * p * (1.0 - p))) + 8;
        const int lo_s = (int)floor(r) - margin, hi_s = (int)ceil(r) + 1 + margin;
        have[0] = lo_s > 0; have[1] = hi_s < m - 1;
        rk[0] = have[0] ? (unsigned int)lo_s : 0u; rk[1] = have[1] ? (unsigned int)hi_s : (unsigned int)(m - 1);
        pre[0] = 0; pre[1] = 0;
    }
    for (int i = threadIdx.x; i < PCT_SAMPLE; i += 1024)
        keys[i] = i < m ? f2key(x[m == PCT_SAMPLE ? (long long)(((unsigned long long)i * (unsigned long long)N) >> 13) : (long long)i])
                        : 0xffffffffu;      // floor(i N / m) in integers: the fp64 division it replaces cost ~20 us on this part's FP64 pipe
    const int lane = threadIdx.x & 31;
    for (int pass = 0; pass < 4; ++pass) {
        if (threadIdx.x < 512) (&h[0][0])[threadIdx.x] = 0;
        __syncthreads();
        const int shift = 24 - 8 * pass;
        const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
        const unsigned int p0 = pre[0], p1 = pre[1];
        for (int i = threadIdx.x; i < PCT_SAMPLE; i += 1024) {
            const unsigned int k = keys[i];
            const bool valid = i < m;
            const unsigned int d = (k >> shift) & 255u;
            const unsigned int b0v = (valid && ((k ^ p0) & himask) == 0) ? d : 256u;
            const unsigned int peers0 = __match_any_sync(0xffffffffu, b0v);
            if (b0v < 256u && lane == __ffs(peers0) - 1) atomicAdd(&h[0][b0v], (unsigned int)__popc(peers0));
            const unsigned int b1v = (valid && ((k ^ p1) & himask) == 0) ? d : 256u;
            const unsigned int peers1 = __match_any_sync(0xffffffffu, b1v);
            if (b1v < 256u && lane == __ffs(peers1) - 1) atomicAdd(&h[1][b1v], (unsigned int)__popc(peers1));
        }
        __syncthreads();
        if (threadIdx.x < 64) {
            const int r = threadIdx.x >> 5;
            unsigned int rank = rk[r], prefix = pre[r];
            pct_digit(h[r], rank, prefix, shift);
            if (lane == 0) { rk[r] = rank; pre[r] = prefix; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s->klo = have[0] ? pre[0] : 0u;
        s->khi = have[1] ? pre[1] : 0xffffffffu;
    }
}

// The one full read: count keys below the bracket, append keys inside it.  Warps are independent (no block barrier in the loop):
// each warp streams 512-score chunks (4 x float4 per lane in flight), stages its candidates in its own shared-memory buffer and
// flushes it with one global atomic when it runs full.
constexpr int PCT_STAGE = 768;            // per-warp staging capacity (>= 512 + flush threshold slack)
__global__ void __launch_bounds__(256, 5) pct_filter_kernel(const float* __restrict__ x, long long N, PctState* s, float* __restrict__ cand) {
    __shared__ float stage[8][PCT_STAGE];
    __shared__ unsigned int wsum[8];
    const unsigned int klo = s->klo, khi = s->khi;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* st = stage[warp];
    unsigned int below = 0, staged = 0;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    const long long n_chunks = (N + 511) / 512;
    const long long wid = (long long)blockIdx.x * 8 + warp, nw = (long long)gridDim.x * 8;
    auto flush = [&]() {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(&s->n_cand, staged);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (unsigned int i = lane; i < staged; i += 32) cand[base + i] = st[i];
        __syncwarp();
        staged = 0;
    };
    const unsigned int span = khi - klo;
    for (long long c = wid; c < n_chunks; c += nw) {
        const long long base = c * 512;
        float v[16];
        unsigned int mask16 = 0;
        if (vec && base + 512 <= N) {                         // full chunk: no validity bookkeeping in the hot loop
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 f = __ldcs(reinterpret_cast<const float4*>(x + base + q * 128 + lane * 4));
                v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const unsigned int k = f2key(v[e]);
                below += k < klo ? 1u : 0u;
                if (k - klo <= span && k >= klo) mask16 |= 1u << e;
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const long long i = base + (e >> 2) * 128 + lane * 4 + (e & 3);
                const bool ok = i < N;
                v[e] = ok ? x[i] : 0.f;
                const unsigned int k = f2key(v[e]);
                if (ok) {
                    below += k < klo ? 1u : 0u;
                    if (k >= klo && k <= khi) mask16 |= 1u << e;
                }
            }
        }
        const unsigned int mine = __popc(mask16);
        unsigned int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
        const unsigned int wtot = __shfl_sync(0xffffffffu, incl, 31);
        if (wtot) {
            if (staged + wtot > PCT_STAGE) flush();
            unsigned int o = staged + incl - mine;
#pragma unroll
            for (int e = 0; e < 16; ++e) if (mask16 & (1u << e)) st[o++] = v[e];
            staged += wtot;
            __syncwarp();
        }
    }
    if (staged) flush();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0) wsum[warp] = below;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tot = 0;
        for (int i = 0; i < 8; ++i) tot += wsum[i];
        if (tot) atomicAdd(&s->n_below, tot);
    }
}

// Are both wanted ranks inside the bracket?  Then the select runs on the candidates with ranks shifted by n_below.
struct PctSel { bool use_cand; unsigned int n; unsigned int rank[2]; int skip; unsigned int known; };
__device__ __forceinline__ PctSel pct_decide(const PctState* s, long long N) {
    PctSel d;
    const unsigned int r0 = s->rank0[0], r1 = s->rank0[1];
    d.use_cand = s->n_below <= r0 && (unsigned long long)r1 < (unsigned long long)s->n_below + s->n_cand;
    d.n = d.use_cand ? s->n_cand : (unsigned int)N;
    d.rank[0] = d.use_cand ? r0 - s->n_below : r0;
    d.rank[1] = d.use_cand ? r1 - s->n_below : r1;
    // every candidate lies in [klo, khi]: the leading bytes the two keys share are the leading digits of both order statistics,
    // so those passes need no histogram
    d.skip = 0; d.known = 0;
    if (d.use_cand) {
        const unsigned int diff = s->klo ^ s->khi;
        d.skip = diff == 0 ? 4 : (__clz(diff) >> 3);
        d.known = d.skip == 0 ? 0u : (s->klo & (0xffffffffu << (32 - 8 * d.skip)));
    }
    return d;
}

// One warp resolves the digits of passes [0, n_pass) of rank slot r from the finished histograms: prefix and remaining rank.
__device__ __forceinline__ void pct_chain(const PctState* s, const PctSel& sel, int n_pass, int r, unsigned int& prefix_out) {
    unsigned int prefix = sel.known, rank = sel.rank[r];
    for (int p = sel.skip; p < n_pass; ++p) pct_digit(s->hist[p][r], rank, prefix, 24 - 8 * p);
    prefix_out = n_pass >= 4 ? prefix : (prefix & (n_pass == 0 ? 0u : (0xffffffffu << (32 - 8 * n_pass))));
}

__global__ void __launch_bounds__(256)
pct_hist_kernel(const float* __restrict__ x_full, const float* __restrict__ cand, long long N_full, PctState* s, int pass) {
    __shared__ unsigned int h[2][256];
    __shared__ unsigned int pre[2];
    const PctSel sel = pct_decide(s, N_full);
    if (pass < sel.skip) return;                              // digit already fixed by the bracket
    h[0][threadIdx.x] = 0; h[1][threadIdx.x] = 0;
    if (threadIdx.x < 64) {                                   // warp r resolves rank slot r up to the previous pass
        const int r = threadIdx.x >> 5;
        unsigned int pf;
        pct_chain(s, sel, pass, r, pf);
        if ((threadIdx.x & 31) == 0) pre[r] = pf;
    }
    __syncthreads();
    const float* __restrict__ x = sel.use_cand ? cand : x_full;
    const long long N = sel.n;
    const int shift = 24 - 8 * pass;
    const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    const unsigned int p0 = pre[0], p1 = pre[1];
    const bool same = p0 == p1;
    const int lane = threadIdx.x & 31;
    for (long long b0 = (long long)blockIdx.x * 1024; b0 < N; b0 += (long long)gridDim.x * 1024) {
        unsigned int k[4];
        bool valid[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {                         // four independent loads in flight per thread
            const long long i = b0 + e * 256 + threadIdx.x;
            valid[e] = i < N;
            k[e] = valid[e] ? f2key(x[i]) : 0u;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int d = (k[e] >> shift) & 255u;
            // warp-aggregated shared-memory atomics: inside the bracket most keys share their leading digits
            const unsigned int b0v = (valid[e] && ((k[e] ^ p0) & himask) == 0) ? d : 256u;
            const unsigned int peers0 = __match_any_sync(0xffffffffu, b0v);
            if (b0v < 256u && lane == __ffs(peers0) - 1) atomicAdd(&h[0][b0v], (unsigned int)__popc(peers0));
            if (!same) {
                const unsigned int b1v = (valid[e] && ((k[e] ^ p1) & himask) == 0) ? d : 256u;
                const unsigned int peers1 = __match_any_sync(0xffffffffu, b1v);
                if (b1v < 256u && lane == __ffs(peers1) - 1) atomicAdd(&h[1][b1v], (unsigned int)__popc(peers1));
            }
        }
    }
    __syncthreads();
    const unsigned int c0 = h[0][threadIdx.x], c1 = same ? c0 : h[1][threadIdx.x];
    if (c0) atomicAdd(&s->hist[pass][0][threadIdx.x], c0);
    if (c1) atomicAdd(&s->hist[pass][1][threadIdx.x], c1);
}

__global__ void pct_result_kernel(PctState* s, long long N_full, double* result) {
    const PctSel sel = pct_decide(s, N_full);
    __shared__ unsigned int pre[2];
    const int r = threadIdx.x >> 5;                           // 64 threads
    unsigned int pf;
    pct_chain(s, sel, 4, r, pf);
    if ((threadIdx.x & 31) == 0) pre[r] = pf;
    __syncthreads();
    if (threadIdx.x == 0) {
        const float a = key2f(pre[0]), b = key2f(pre[1]);
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
    pct_sample_kernel<<<1, 1024, 0, st>>>(scores, N, (float)q, s);
    SHM_LAUNCH_CHECK();
    const int nsm = device_sm_count(dev);
    const int fgrid = (int)min((long long)nsm * 5, (long long)((N + 4095) / 4096));     // one resident wave (5 CTAs per SM)
    pct_filter_kernel<<<fgrid, 256, 0, st>>>(scores, N, s, cand);
    SHM_LAUNCH_CHECK();
    // the select reads a few per cent of N in the common case, so a modest grid suffices; the grid-stride loop still covers
    // the full array in the fallback
    const int grid = (int)min((long long)nsm * 4, (long long)((N + 1023) / 1024));
    for (int pass = 0; pass < 4; ++pass) {
        pct_hist_kernel<<<grid, 256, 0, st>>>(scores, cand, N, s, pass);
        SHM_LAUNCH_CHECK();
    }
    pct_result_kernel<<<1, 64, 0, st>>>(s, N, result);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
