// Threshold + routing: flag = score > thr (strict, fp32) and the ascending list of flagged window
// indices -- `anom_mask = mse_all > thr; idx = np.where(anom_mask)[0]`
// (4DOF/Scripts/06_test_full_pipeline.py:350-351, openLAB 10_test_hybrid_pipeline.py:367).
//
// HBM-bound (4 B read, 1 B flag + 4 B per flagged window written).  Three dependency-free launches instead of a single-pass
// decoupled look-back scan (measured in r01: with ~600 tiles resident the look-back walk held 65 % of the warps at a barrier
// and the kernel reached 53 % of the copy bandwidth at 1 % flagged):
//   1. compact_flag_kernel   streams the scores once: flag bytes, a 1-bit-per-score mask (0.125 B) and the tile's count
//   2. compact_scan_kernel   exclusive scan of the tile counts (one CTA; 8 B per 16,384 scores)
//   3. compact_write_kernel  re-reads only the bit mask and writes each tile's indices as coalesced stores
// Nothing waits on another CTA, so there is no forward-progress assumption and the output order equals np.where's.
#include "common.cuh"

namespace shm {

constexpr int CP_THREADS = 512;
constexpr int CP_ITEMS = 32;          // 16384 scores (64 KB) per tile, 4 tiles per SM: 128 KB of loads in flight per SM in each of the two
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;      // load batches

__global__ void __launch_bounds__(CP_THREADS, 4)
compact_flag_kernel(const float* __restrict__ score, float thr, long long N, unsigned char* __restrict__ flag,
                    unsigned* __restrict__ mask, int* __restrict__ tile_count, int vec_ok) {
    __shared__ int s_warp[CP_THREADS / 32];
    __shared__ __align__(16) unsigned char s_nib[CP_TILE / 4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const long long tile_base = (long long)tile * CP_TILE;
    const long long base = tile_base + (long long)tid * CP_ITEMS;
    const bool full = vec_ok && tile_base + CP_TILE <= N;
    unsigned bits = 0;
    if (full) {
        // coalesced 16-byte loads (consecutive lanes -> consecutive float4s), flags parked as nibbles in shared memory,
        // then every thread picks up the 32 consecutive scores it owns
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
    mask[(long long)tile * CP_THREADS + tid] = bits;
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
    int cnt = __popc(bits);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) s_warp[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int agg = 0;
#pragma unroll
        for (int w = 0; w < CP_THREADS / 32; ++w) agg += s_warp[w];
        tile_count[tile] = agg;
    }
}

// exclusive scan of the tile counts (one CTA, 8 consecutive tiles per thread = 8192 tiles per round, running carry) and the total
constexpr int CS_ITEMS = 8;
__global__ void __launch_bounds__(1024)
compact_scan_kernel(const int* __restrict__ tile_count, int n_tiles, int* __restrict__ tile_off, int* __restrict__ count) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < n_tiles; b0 += 1024 * CS_ITEMS) {
        const int t0 = b0 + tid * CS_ITEMS;
        int v[CS_ITEMS];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < CS_ITEMS; ++k) {
            v[k] = (t0 + k < n_tiles) ? tile_count[t0 + k] : 0;
            sum += v[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;                                          // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        int excl = carry + (warp ? s_warp[warp - 1] : 0) + incl - sum;
#pragma unroll
        for (int k = 0; k < CS_ITEMS; ++k) {
            if (t0 + k < n_tiles) tile_off[t0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) *count = s_carry;
}

__global__ void __launch_bounds__(CP_THREADS, 4)
compact_write_kernel(const unsigned* __restrict__ mask, const int* __restrict__ tile_count, const int* __restrict__ tile_off,
                     int* __restrict__ idx) {
    __shared__ int s_warp[CP_THREADS / 32];
    __shared__ unsigned short s_idx[CP_TILE];                       // the tile's flagged offsets, in order: written out coalesced
    const int tile = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int agg = tile_count[tile];                               // the three loads are independent: one round trip
    const int excl = tile_off[tile];
    unsigned b = mask[(long long)tile * CP_THREADS + tid];
    if (agg == 0) return;
    const int cnt = __popc(b);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int loc = incl - cnt;
#pragma unroll
    for (int w = 0; w < CP_THREADS / 32; ++w) loc += (w < warp) ? s_warp[w] : 0;
    while (b) {
        const int i = __ffs(b) - 1;
        b &= b - 1;
        s_idx[loc++] = (unsigned short)(tid * CP_ITEMS + i);
    }
    __syncthreads();
    const long long tile_base = (long long)tile * CP_TILE;
    for (int j = tid; j < agg; j += CP_THREADS) idx[excl + j] = (int)(tile_base + s_idx[j]);
}

}  // namespace shm

// header (16 B) | tile_count int[tiles] | tile_off int[tiles] | mask uint32[tiles * 512]
extern "C" int64_t shm_compact_workspace_bytes(int64_t N) {
    if (N < 0) return 0;
    int64_t tiles = (N + shm::CP_TILE - 1) / shm::CP_TILE;
    if (tiles < 1) tiles = 1;
    return 16 + 8 * tiles + 4 * (int64_t)shm::CP_THREADS * tiles;
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
    int* tile_count = reinterpret_cast<int*>(static_cast<char*>(workspace) + 16);
    int* tile_off = tile_count + n_tiles;
    unsigned* mask = reinterpret_cast<unsigned*>(tile_off + n_tiles);
    const int vec_ok = ((reinterpret_cast<uintptr_t>(score) & 15) == 0) && ((reinterpret_cast<uintptr_t>(flag) & 15) == 0);
    compact_flag_kernel<<<n_tiles, CP_THREADS, 0, st>>>(score, thr, N, flag, mask, tile_count, vec_ok);
    SHM_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(tile_count, n_tiles, tile_off, count);
    SHM_LAUNCH_CHECK();
    compact_write_kernel<<<n_tiles, CP_THREADS, 0, st>>>(mask, tile_count, tile_off, idx);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
