// Threshold + routing: flag = score > thr (strict, fp32) and the ascending list of flagged window
// indices -- `anom_mask = mse_all > thr; idx = np.where(anom_mask)[0]`
// (4DOF/Scripts/06_test_full_pipeline.py:350-351, openLAB 10_test_hybrid_pipeline.py:367).
//
// Single pass, HBM-bound (4 B read, 1 B flag + 4 B per flagged window written): tiles are claimed in
// order through an atomic ticket, each CTA ballots its flags, publishes its tile aggregate and
// resolves its exclusive prefix with a warp-wide decoupled look-back, so the scores are read exactly
// once and the output order equals np.where's.  The tile's indices are staged in shared memory while
// the look-back runs and leave as coalesced stores.
#include "common.cuh"

namespace shm {

constexpr int CP_THREADS = 512;
constexpr int CP_ITEMS = 32;          // 16384 scores (64 KB) per tile, 4 tiles per SM: 128 KB of loads in flight per SM in each of the two
                                      // load batches, and one tile's look-back / write-out phase runs under the other tiles' load phase
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

constexpr unsigned long long ST_AGG = 1ull << 62;
constexpr unsigned long long ST_INC = 2ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(CP_THREADS, 4)
compact_kernel(const float* __restrict__ score, float thr, long long N, unsigned char* __restrict__ flag,
               int* __restrict__ idx, int* __restrict__ count, int* ticket, unsigned long long* status, int n_tiles,
               int vec_ok) {
    __shared__ int s_tile;
    __shared__ int s_warp[CP_THREADS / 32];
    __shared__ int s_excl;
    __shared__ __align__(16) unsigned char s_nib[CP_TILE / 4];
    __shared__ unsigned short s_idx[CP_TILE];                       // the tile's flagged offsets, in order: written out coalesced
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tile = blockIdx.x: CTAs are dispatched in index order, so every predecessor a look-back waits for is resident or done
    // (the forward-progress assumption of every single-pass decoupled look-back scan); a ticket counter would serialise
    // one same-address atomic per tile
    const int tile = blockIdx.x;
    (void)ticket; (void)s_tile;
    const long long tile_base = (long long)tile * CP_TILE;
    const long long base = tile_base + (long long)tid * CP_ITEMS;
    const bool full = vec_ok && tile_base + CP_TILE <= N;
    unsigned bits = 0;
    if (full) {
        // coalesced 16-byte loads (consecutive lanes -> consecutive float4s), flags parked as nibbles in shared memory,
        // then every thread picks up the 16 consecutive scores it owns for the scan
        const float4* s4 = reinterpret_cast<const float4*>(score + tile_base);
#pragma unroll
        for (int h = 0; h < 2; ++h) {                                   // two batches of 4 loads: 32 registers per thread at 4 CTAs per SM
            float4 a[CP_ITEMS / 8];
#pragma unroll
            for (int q = 0; q < CP_ITEMS / 8; ++q) a[q] = __ldcs(s4 + (h * (CP_ITEMS / 8) + q) * CP_THREADS + tid);
#pragma unroll
            for (int q = 0; q < CP_ITEMS / 8; ++q)
                s_nib[(h * (CP_ITEMS / 8) + q) * CP_THREADS + tid] =
                    (unsigned char)((a[q].x > thr ? 1u : 0u) | (a[q].y > thr ? 2u : 0u) | (a[q].z > thr ? 4u : 0u) |
                                    (a[q].w > thr ? 8u : 0u));                              // NaN > thr is false, as NumPy
        }
        __syncthreads();
        const unsigned long long nb = *reinterpret_cast<const unsigned long long*>(s_nib + tid * (CP_ITEMS / 4));
#pragma unroll
        for (int j = 0; j < CP_ITEMS / 4; ++j) bits |= (unsigned)((nb >> (8 * j)) & 0xFull) << (4 * j);
    } else {
#pragma unroll 4
        for (int i = 0; i < CP_ITEMS; ++i) bits |= (base + i < N && score[base + i] > thr) ? (1u << i) : 0u;
    }
    const int cnt = __popc(bits);

    if (flag) {
        if (full) {
#pragma unroll
            for (int h = 0; h < CP_ITEMS / 16; ++h) {
                unsigned long long p0 = 0, p1 = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    p0 |= (unsigned long long)((bits >> (16 * h + i)) & 1u) << (8 * i);
                    p1 |= (unsigned long long)((bits >> (16 * h + 8 + i)) & 1u) << (8 * i);
                }
                reinterpret_cast<ulonglong2*>(flag + base)[h] = make_ulonglong2(p0, p1);
            }
        } else {
            for (int i = 0; i < CP_ITEMS; ++i) if (base + i < N) flag[base + i] = (bits >> i) & 1u;
        }
    }

    // block-wide exclusive scan of per-thread counts
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int warp_off = 0, agg = 0;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) {
        const int c = s_warp[w];
        if (w < warp) warp_off += c;
        agg += c;
    }
    // publish the aggregate first, then stage this tile's indices while the predecessors resolve
    if (tid == 0) st_status(status + tile, (tile == 0 ? ST_INC : ST_AGG) | (unsigned)agg);
    {
        int loc = warp_off + (incl - cnt);
        unsigned b = bits;
        while (b) {
            const int i = __ffs(b) - 1;
            b &= b - 1;
            s_idx[loc++] = (unsigned short)(tid * CP_ITEMS + i);
        }
    }

    // decoupled look-back (warp 0): 32 predecessors per round.  (A 512-wide round -- 16 statuses per lane -- was measured and is
    // slower, 27 % instead of 53 % of the HBM roof: every spin re-reads 4 KB of statuses and the window almost always holds an
    // unpublished tile.)
    if (warp == 0) {
        int excl = 0;
        int look = tile - 1;
        while (look >= 0) {
            const int t = look - lane;
            unsigned long long st = (t >= 0) ? ld_status(status + t) : ST_INC;
            while (__any_sync(0xffffffffu, (st >> 62) == 0)) st = (t >= 0) ? ld_status(status + t) : ST_INC;
            const unsigned inc_mask = __ballot_sync(0xffffffffu, (st >> 62) == 2);
            const int val = (int)(st & 0xffffffffu);
            if (inc_mask) {
                const int first = __ffs(inc_mask) - 1;           // nearest predecessor with an inclusive prefix
                int part = (lane <= first) ? val : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                excl += part;
                break;
            }
            int part = val;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            excl += part;
            look -= 32;
        }
        if (lane == 0) {
            if (tile != 0) st_status(status + tile, ST_INC | (unsigned)(excl + agg));
            s_excl = excl;
            if (tile == n_tiles - 1) *count = excl + agg;
        }
    }
    __syncthreads();
    const int excl = s_excl;
    for (int j = tid; j < agg; j += CP_THREADS) idx[excl + j] = (int)(tile_base + s_idx[j]);
}

}  // namespace shm

extern "C" int64_t shm_compact_workspace_bytes(int64_t N) {
    if (N < 0) return 0;
    const int64_t tiles = (N + shm::CP_TILE - 1) / shm::CP_TILE;
    return 16 + 8 * (tiles > 0 ? tiles : 1);
}

extern "C" int shm_compact(const float* score, float thr, int64_t N, uint8_t* flag, int32_t* idx, int32_t* count,
                           void* workspace, void* stream) {
    using namespace shm;
    if (N < 0 || N > 0x7fffffffLL || !count || !workspace || (N > 0 && (!score || !idx))) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (N == 0) {
        SHM_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
        return SHM_OK;
    }
    const int n_tiles = (int)((N + CP_TILE - 1) / CP_TILE);
    SHM_CUDA(cudaMemsetAsync(workspace, 0, (size_t)shm_compact_workspace_bytes(N), st));
    int* ticket = static_cast<int*>(workspace);
    unsigned long long* status = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + 16);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(score) & 15) == 0) && ((reinterpret_cast<uintptr_t>(flag) & 15) == 0);
    compact_kernel<<<n_tiles, CP_THREADS, 0, st>>>(score, thr, N, flag, idx, count, ticket, status, n_tiles, vec_ok);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
