// fp32-grade strided contraction on the 5th-generation tensor cores (sm_100a) for the training steps:
//   C[m*ldc + n] (+)= sum_k A[m*a_ms + k*a_ks] * B[k*b_ks + n*b_ns] (+ bias1[n] + bias2[n])
// The LSTM input projections over all timesteps, their dX twins and the split-K dW contractions of
// 4DOF/Scripts/03_train_vae.py:262-268 (forward + loss.backward()) are dense contractions with 25,600 rows or a 25,600-long
// reduction; north_star asks for them on tcgen05.  Operands stay fp32 in HBM: loader threads split every value into a
// 16-bit hi and lo half on the fly (fp16 halves for the forward activations: 2^-22 relative, absolute floor 2^-25; bf16 halves
// for gradients, whose magnitudes need fp32's exponent range: 2^-16 relative) and write them as K-major core-matrix images;
// D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with fp32 accumulation in TMEM (3 tcgen05.mma.kind::f16 passes).
//
// One CTA = one 128 x 128 output tile; K runs in chunks of 32 through two shared-memory stages: while the tensor core
// consumes chunk c (6 MMAs, released by tcgen05.commit -> mbarrier), all 256 threads fetch chunk c+1 from global memory into
// registers.  Either operand may be K-contiguous or M/N-contiguous in memory (float4 loads along the contiguous dimension).
// Split-K (blockIdx.z) with a vector atomicAdd epilogue for the dW contractions.
#include "gemm.cuh"
#include "tcgen05.cuh"

namespace shm {
using namespace tc;

constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 32;
constexpr int GT_IMG = 128 * GT_BK * 2;          // one 128 x 32 16-bit image: 8 KB (4 k-core blocks of 2 KB)
constexpr int GT_STAGE = 4 * GT_IMG;             // A_hi, A_lo, B_hi, B_lo
constexpr int GT_SMEM = 2 * GT_STAGE + 64;

struct TcGemmArgs {
    const float* A; long long a_ms, a_ks;
    const float* B; long long b_ks, b_ns;
    float* C; long long ldc;
    int M, N, K;
    const float* bias1; const float* bias2;
    int kchunk, atomic, a_kcontig, b_kcontig, bf16;
};

// 128 rows x 32 k of an operand whose element (r, k) sits at P[r*rs + k*ks]; 4 float4 per thread, each ending up as 4 consecutive
// k of ONE row (what the 8-byte core-matrix store wants).
//   kcontig (ks == 1): item W = e*8 + warp -> row group g = W>>1 (8 rows), k half W&1; lane -> row g*8 + (lane&7), k quad (W&1)*4 + (lane>>3):
//     a warp reads 8 rows x 64 B and writes 2 x 128 contiguous bytes per image.
//   otherwise (rs == 1): lanes run along the rows -- lane = (row quad mq = lane&3, kk = (lane>>2)&3, k-quad parity kh = lane>>4) reads
//     rows rb*16 + 4mq..+3 at k = ko*8 + kh*4 + kk (W -> rb = W&7, ko = W>>3), then a 4x4 transpose over the 4 lanes that differ in kk
//     (4 shuffles) leaves row rb*16 + 4mq + kk, k = ko*8 + kh*4..+3 in the lane: the warp then writes 256 contiguous bytes per image.
//     (r02 first version: 2-byte scattered stores, 5.7 M bank conflicts per dW launch -- profiles/r02_tc_gemm_raw.csv.)
__device__ __forceinline__ void gt_load(const float* __restrict__ P, long long rs, long long ks, int rows_valid, int k0, int kend,
                                        bool kcontig, float4 (&v)[4]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int W = e * 8 + warp;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kcontig) {
            const int row = (W >> 1) * 8 + (lane & 7);
            const int k = k0 + ((W & 1) * 4 + (lane >> 3)) * 4;
            if (row < rows_valid && k < kend) f = __ldg(reinterpret_cast<const float4*>(P + (long long)row * rs + k));
        } else {
            const int row0 = (W & 7) * 16 + (lane & 3) * 4;
            const int k = k0 + (W >> 3) * 8 + (lane >> 4) * 4 + ((lane >> 2) & 3);
            if (row0 < rows_valid && k < kend) f = __ldg(reinterpret_cast<const float4*>(P + (long long)k * ks + row0));
        }
        v[e] = f;
    }
}

// 4x4 transpose over the 4 lanes that differ in lane bits 2..3: in: lane kk holds X[row 0..3][k = kk]; out: X[row = kk][k 0..3]
__device__ __forceinline__ float4 gt_transpose4(float4 a, int kk) {
    const bool hiA = (kk & 2) != 0, hiB = (kk & 1) != 0;
    const float r0 = __shfl_xor_sync(0xffffffffu, hiA ? a.x : a.z, 8);
    const float r1 = __shfl_xor_sync(0xffffffffu, hiA ? a.y : a.w, 8);
    const float t0 = hiA ? r0 : a.x, t1 = hiA ? r1 : a.y, t2 = hiA ? a.z : r0, t3 = hiA ? a.w : r1;
    const float s0 = __shfl_xor_sync(0xffffffffu, hiB ? t0 : t1, 4);
    const float s1 = __shfl_xor_sync(0xffffffffu, hiB ? t2 : t3, 4);
    return make_float4(hiB ? s0 : t0, hiB ? t1 : s0, hiB ? s1 : t2, hiB ? t3 : s1);
}

// registers -> hi / lo K-major core-matrix images: byte offset(r, k) = (k/8)*2048 + (r/8)*128 + (r%8)*16 + (k%8)*2
__device__ __forceinline__ void gt_store(const float4 (&v)[4], unsigned char* hi_img, unsigned char* lo_img, bool kcontig, bool bf16) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int W = e * 8 + warp;
        float4 x = v[e];
        int row, kq;                                  // this lane's row and k quad (4 consecutive k) of the 128 x 32 tile
        if (kcontig) {
            row = (W >> 1) * 8 + (lane & 7);
            kq = (W & 1) * 4 + (lane >> 3);
        } else {
            const int kk = (lane >> 2) & 3;
            x = gt_transpose4(x, kk);
            row = (W & 7) * 16 + (lane & 3) * 4 + kk;
            kq = (W >> 3) * 2 + (lane >> 4);
        }
        uint32_t h0, l0, h1, l1;
        if (bf16) { split_bf16x2(x.x, x.y, h0, l0); split_bf16x2(x.z, x.w, h1, l1); }
        else { split_f16x2(x.x, x.y, h0, l0); split_f16x2(x.z, x.w, h1, l1); }
        const uint32_t off = (uint32_t)(kq >> 1) * 2048u + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u + (uint32_t)(kq & 1) * 8u;
        *reinterpret_cast<uint2*>(hi_img + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(lo_img + off) = make_uint2(l0, l1);
    }
}

__global__ void __launch_bounds__(256, 2) tc_gemm_kernel(const TcGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem + 2 * GT_STAGE);       // [2]
    uint64_t* bar_done = bar_free + 2;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bar_free + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * GT_BM, n0 = blockIdx.x * GT_BN;
    const int kbeg = blockIdx.z * g.kchunk;
    const int kend = min(g.K, kbeg + g.kchunk);
    const int n_chunks = (kend - kbeg + GT_BK - 1) / GT_BK;
    if (tid == 0) {
        mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(bar_done, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(tmem_holder, GT_BN);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tbase = *tmem_holder;

    const float* Ap = g.A + (long long)m0 * g.a_ms;
    const float* Bp = g.B + (long long)n0 * g.b_ns;
    const int a_rows = g.M - m0, b_rows = g.N - n0;
    const bool akc = g.a_kcontig != 0, bkc = g.b_kcontig != 0, bf = g.bf16 != 0;
    const uint32_t idesc = bf ? make_idesc_bf16(GT_BM, GT_BN) : make_idesc_f16(GT_BM, GT_BN);
    float4 va[4], vb[4];
    if (n_chunks > 0) {
        gt_load(Ap, g.a_ms, g.a_ks, a_rows, kbeg, kend, akc, va);
        gt_load(Bp, g.b_ns, g.b_ks, b_rows, kbeg, kend, bkc, vb);
    }
    for (int c = 0; c < n_chunks; ++c) {
        const int s = c & 1;
        unsigned char* st = smem + s * GT_STAGE;
        if (c >= 2) {                                     // the MMAs that read this stage two chunks ago have completed
            mbar_wait(&bar_free[s], ((c >> 1) - 1) & 1);
            tc_fence_after_sync();
        }
        gt_store(va, st, st + GT_IMG, akc, bf);
        gt_store(vb, st + 2 * GT_IMG, st + 3 * GT_IMG, bkc, bf);
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after_sync();
            if (elect_one()) {
                const uint32_t base = smem_u32(st);
#pragma unroll
                for (int j = 0; j < GT_BK / 16; ++j) {
                    const uint32_t koff = (uint32_t)j * 4096u;                  // two k-core blocks per 16-wide MMA step
                    const uint64_t a_hi = make_smem_desc(base + koff, 2048, 128);
                    const uint64_t a_lo = make_smem_desc(base + GT_IMG + koff, 2048, 128);
                    const uint64_t b_hi = make_smem_desc(base + 2 * GT_IMG + koff, 2048, 128);
                    const uint64_t b_lo = make_smem_desc(base + 3 * GT_IMG + koff, 2048, 128);
                    mma_ss(tbase, a_hi, b_hi, idesc, (c | j) ? 1u : 0u);
                    mma_ss(tbase, a_lo, b_hi, idesc, 1u);
                    mma_ss(tbase, a_hi, b_lo, idesc, 1u);
                }
                mma_commit(&bar_free[s]);
                if (c == n_chunks - 1) mma_commit(bar_done);
            }
            __syncwarp();
        }
        if (c + 1 < n_chunks) {                           // next chunk's global loads fly while the tensor core works
            const int k0 = kbeg + (c + 1) * GT_BK;
            gt_load(Ap, g.a_ms, g.a_ks, a_rows, k0, kend, akc, va);
            gt_load(Bp, g.b_ns, g.b_ks, b_rows, k0, kend, bkc, vb);
        }
    }
    // epilogue: warp w reads TMEM lanes 32*(w%4).. (its quarter), columns 64*(w/4)..+63
    if (n_chunks > 0) {
        mbar_wait(bar_done, 0);
        tc_fence_after_sync();
        const int row = m0 + (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool add_bias = blockIdx.z == 0;
        const bool vec_ok = (g.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(g.C) & 15) == 0);
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            const int c0 = (warp >> 2) * 64 + cb * 32;
            uint32_t v[32];
            tmem_ld32(tbase + lane_base + (uint32_t)c0, v);
            tmem_ld_wait();
            if (row < g.M) {
                float* crow = g.C + (long long)row * g.ldc + n0 + c0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = n0 + c0 + 4 * q;
                    float f[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f[j] = __uint_as_float(v[4 * q + j]);
                        if (add_bias && n + j < g.N) {
                            if (g.bias1) f[j] += __ldg(g.bias1 + n + j);
                            if (g.bias2) f[j] += __ldg(g.bias2 + n + j);
                        }
                    }
                    if (n + 3 < g.N && vec_ok) {
                        if (g.atomic) atomicAdd(reinterpret_cast<float4*>(crow + 4 * q), make_float4(f[0], f[1], f[2], f[3]));
                        else *reinterpret_cast<float4*>(crow + 4 * q) = make_float4(f[0], f[1], f[2], f[3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < g.N) { if (g.atomic) atomicAdd(crow + 4 * q + j, f[j]); else crow[4 * q + j] = f[j]; }
                    }
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, GT_BN);
}

// Is the tensor-core path applicable?  Large enough to pay, and float4-loadable along a contiguous dimension of each operand.
static bool tc_gemm_ok(const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, int M, int N, int K) {
    if (M < 128 || N < 32 || K < 4) return false;
    if ((double)M * N * K < 3.0e7) return false;
    if ((K & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
    const bool a_k = a_ks == 1 && (a_ms & 3) == 0;
    const bool a_m = a_ms == 1 && (a_ks & 3) == 0 && (M & 3) == 0;
    const bool b_k = b_ks == 1 && (b_ns & 3) == 0;
    const bool b_n = b_ns == 1 && (b_ks & 3) == 0 && (N & 3) == 0;
    return (a_k || a_m) && (b_k || b_n);
}

int tc_gemm(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
            long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk, int mode) {
    if (mode == SHM_GEMM_SIMT || !tc_gemm_ok(A, a_ms, a_ks, B, b_ks, b_ns, M, N, K)) return 1;      // caller falls back to the SIMT kernel
    static bool attr_done[64] = {false};
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_done[dev]) {
        SHM_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
        attr_done[dev] = true;
    }
    TcGemmArgs g{A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, K, splitk ? 1 : 0,
                 (a_ks == 1 && (a_ms & 3) == 0) ? 1 : 0, (b_ks == 1 && (b_ns & 3) == 0) ? 1 : 0, mode == SHM_GEMM_TC_BF16X3 ? 1 : 0};
    const int tiles = ((M + GT_BM - 1) / GT_BM) * ((N + GT_BN - 1) / GT_BN);
    int splits = 1;
    if (splitk) {
        const int nsm = device_sm_count(dev);
        splits = max(1, min((K + 255) / 256, (2 * nsm + tiles - 1) / tiles));
        g.kchunk = ((K + splits - 1) / splits + GT_BK - 1) / GT_BK * GT_BK;
        splits = (K + g.kchunk - 1) / g.kchunk;
    }
    dim3 grid((N + GT_BN - 1) / GT_BN, (M + GT_BM - 1) / GT_BM, splits);
    tc_gemm_kernel<<<grid, 256, GT_SMEM, st>>>(g);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

}  // namespace shm

extern "C" int shm_gemm_f32(const float* A, int64_t a_ms, int64_t a_ks, const float* B, int64_t b_ks, int64_t b_ns, float* C, int64_t ldc,
                            int32_t M, int32_t N, int32_t K, const float* bias, int32_t splitk, int32_t mode, void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K < 0) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = shm::check_device(dev);
    if (rc != SHM_OK) return rc;
    if (mode != SHM_GEMM_SIMT && mode != SHM_GEMM_TC_F16X3 && mode != SHM_GEMM_TC_BF16X3) return SHM_ERR_ARG;
    return shm::sgemm_mode(static_cast<cudaStream_t>(stream), A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias, nullptr, splitk != 0, mode);
}
