// 4DOF sensor-fault vs structural-fault CNN, eval mode: 4DOF/Scripts/Models/cnn_model.py:16-34,45-51.
//   x [n,2,100,12] -> conv3x3(2->16,pad1)+BN+ReLU+MaxPool2 -> [16,50,6]
//                  -> conv3x3(16->32,pad1)+BN+ReLU+MaxPool2 -> [32,25,3] -> flatten (C,H,W) = 2400
//                  -> Linear 2400->128 + ReLU (+Dropout = identity) -> Linear 128->2
// plus label = argmax+1 (tie -> 1) and p_struct = softmax[:,1] (06_test_full_pipeline.py:366-372).
//
// Two kernels: (1) one CTA per window keeps both conv stages on chip (padded planes in shared
// memory, BatchNorm folded to one FMA, ReLU and the 2x2 max fused into the accumulation epilogue)
// and writes only the 2400 pooled features; (2) a register-tiled fp32 GEMM over 64-window tiles for
// fc1 with ReLU, fc2, argmax and softmax fused behind it.
#include <new>
#include "common.cuh"

struct shm_cnn4dof {
    int device;
    float* buf;            // all parameters, repacked
    float* feat;           // [cap, 2400] pooled conv features
    int64_t feat_cap;
    // offsets (floats) inside buf
    size_t o_w1, o_a1, o_b1, o_w2, o_a2, o_b2, o_fc1t, o_fc1b, o_fc2w, o_fc2b, total;
    float* raw;            // staging for caller tensors
    size_t raw_total;
};

namespace shm {

constexpr int C4_T = 100, C4_F = 12, C4_C1 = 16, C4_C2 = 32;
constexpr int C4_H1 = 50, C4_W1 = 6, C4_H2 = 25, C4_W2 = 3;
constexpr int C4_FEAT = C4_C2 * C4_H2 * C4_W2;     // 2400
constexpr int C4_THREADS = 320;

// padded planes: in0 [2][102][14], p1 [16][52][8]
constexpr int IN0_W = 14, IN0_H = 102, P1_W = 8, P1_H = 52;

struct Cnn4Dev {
    const float *w1, *a1, *b1;      // conv1 weights [16][2][9], folded BN scale/shift (conv bias included)
    const float *w2, *a2, *b2;      // conv2 weights [32][16][9]
    const float *fc1t, *fc1b;       // fc1 weight transposed [2400][128], bias
    const float *fc2w, *fc2b;
};

// packed fp32 pairs (sm_100 fma.rn.f32x2 -> FFMA2: two lanes per issue slot, same flops per clock; scripts/ffma2_probe.cu)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

__global__ void __launch_bounds__(C4_THREADS)
cnn4dof_conv_kernel(Cnn4Dev P, const float* __restrict__ x, const int* __restrict__ n_dev, long long n,
                    float* __restrict__ feat) {
    extern __shared__ __align__(16) float sm[];
    float* in0 = sm;                                   // 2*102*14 = 2856
    float* p1 = in0 + 2 * IN0_H * IN0_W;               // 16*52*8 = 6656
    float* sw1 = p1 + C4_C1 * P1_H * P1_W;             // 288
    float* sw2 = sw1 + C4_C1 * 2 * 9;                  // 4608
    const int tid = threadIdx.x;
    long long n_eff = n;
    if (n_dev) n_eff = min(n_eff, (long long)__ldg(n_dev));

    // weights staged as [input channel][tap][output channel]: the 8 output channels of an item are two 16-byte loads per tap, and
    // adjacent channels are the two halves of a packed-FMA operand (fma.rn.f32x2 -> FFMA2, two channels per issue slot)
    for (int i = tid; i < C4_C1 * 2 * 9; i += C4_THREADS) {
        const int co = i / 18, r = i - co * 18;                      // r = c * 9 + tap
        sw1[r * C4_C1 + co] = __ldg(P.w1 + i);
    }
    for (int i = tid; i < C4_C2 * C4_C1 * 9; i += C4_THREADS) {
        const int co = i / (C4_C1 * 9), r = i - co * (C4_C1 * 9);    // r = ci * 9 + tap
        sw2[r * C4_C2 + co] = __ldg(P.w2 + i);
    }

    // zero padding of both planes, once: every window overwrites the interiors completely and never touches the borders
    for (int i = tid; i < 2 * IN0_H * IN0_W + C4_C1 * P1_H * P1_W; i += C4_THREADS) sm[i] = 0.f;
    for (long long win = blockIdx.x; win < n_eff; win += gridDim.x) {
        __syncthreads();
        const float* xw = x + win * (2 * C4_T * C4_F);
        for (int i = tid; i < 2 * C4_T * C4_F; i += C4_THREADS) {
            const int c = i / (C4_T * C4_F);
            const int r = i - c * (C4_T * C4_F);
            const int t = r / C4_F, f = r - t * C4_F;
            in0[(c * IN0_H + t + 1) * IN0_W + f + 1] = __ldg(xw + i);
        }
        __syncthreads();

        // conv1 + BN + ReLU + 2x2 max: item = (channel group of 8, pooled position)
        for (int item = tid; item < 2 * C4_H1 * C4_W1; item += C4_THREADS) {
            // channel group fastest: the 2 lanes of a position share (broadcast) the patch loads, and a warp's 16 positions
            // touch few shared-memory banks twice (position-major items cost 35 % of all wavefronts in bank conflicts)
            const int cg = item & 1;
            const int pos = item >> 1;
            const int ph = pos / C4_W1, pw = pos - ph * C4_W1;
            float patch[2][4][4];
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) patch[c][r][q] = in0[(c * IN0_H + 2 * ph + r) * IN0_W + 2 * pw + q];
            f32x2 acc2[4][4];                                        // [channel pair][2x2 pool position]
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) { acc2[cp][0] = 0ull; acc2[cp][1] = 0ull; acc2[cp][2] = 0ull; acc2[cp][3] = 0ull; }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                f32x2 pd[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) pd[r][q] = pk2(patch[c][r][q], patch[c][r][q]);
#pragma unroll
                for (int kr = 0; kr < 3; ++kr)
#pragma unroll
                    for (int kq = 0; kq < 3; ++kq) {
                        const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(sw1 + (c * 9 + kr * 3 + kq) * C4_C1 + cg * 8);
                        const ulonglong2 wa = wp[0], wb = wp[1];
                        const f32x2 w[4] = {wa.x, wa.y, wb.x, wb.y};
#pragma unroll
                        for (int cp = 0; cp < 4; ++cp) {
                            acc2[cp][0] = fma2(w[cp], pd[kr][kq], acc2[cp][0]);
                            acc2[cp][1] = fma2(w[cp], pd[kr][kq + 1], acc2[cp][1]);
                            acc2[cp][2] = fma2(w[cp], pd[kr + 1][kq], acc2[cp][2]);
                            acc2[cp][3] = fma2(w[cp], pd[kr + 1][kq + 1], acc2[cp][3]);
                        }
                    }
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int co = cg * 8 + cc;
                float a[4], o;
#pragma unroll
                for (int j = 0; j < 4; ++j) { if (cc & 1) unpk2(acc2[cc >> 1][j], o, a[j]); else unpk2(acc2[cc >> 1][j], a[j], o); }
                const float s = __ldg(P.a1 + co), b = __ldg(P.b1 + co);
                const float m = fmaxf(fmaxf(fmaf(a[0], s, b), fmaf(a[1], s, b)), fmaxf(fmaf(a[2], s, b), fmaf(a[3], s, b)));
                p1[(co * P1_H + ph + 1) * P1_W + pw + 1] = fmaxf(m, 0.f);
            }
        }
        __syncthreads();

        // conv2 + BN + ReLU + 2x2 max: item = (channel group of 8, pooled position)
        for (int item = tid; item < 4 * C4_H2 * C4_W2; item += C4_THREADS) {
            const int cg = item & 3;                                 // channel group fastest (see conv1)
            const int pos = item >> 2;
            const int ph = pos / C4_W2, pw = pos - ph * C4_W2;
            f32x2 acc2[4][4];                                        // [channel pair][2x2 pool position]
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) { acc2[cp][0] = 0ull; acc2[cp][1] = 0ull; acc2[cp][2] = 0ull; acc2[cp][3] = 0ull; }
            for (int ci = 0; ci < C4_C1; ++ci) {
                f32x2 pd[4][4];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const float v = p1[(ci * P1_H + 2 * ph + r) * P1_W + 2 * pw + q]; pd[r][q] = pk2(v, v); }
#pragma unroll
                for (int kr = 0; kr < 3; ++kr)
#pragma unroll
                    for (int kq = 0; kq < 3; ++kq) {
                        const ulonglong2* wp = reinterpret_cast<const ulonglong2*>(sw2 + (ci * 9 + kr * 3 + kq) * C4_C2 + cg * 8);
                        const ulonglong2 wa = wp[0], wb = wp[1];
                        const f32x2 w[4] = {wa.x, wa.y, wb.x, wb.y};
#pragma unroll
                        for (int cp = 0; cp < 4; ++cp) {
                            acc2[cp][0] = fma2(w[cp], pd[kr][kq], acc2[cp][0]);
                            acc2[cp][1] = fma2(w[cp], pd[kr][kq + 1], acc2[cp][1]);
                            acc2[cp][2] = fma2(w[cp], pd[kr + 1][kq], acc2[cp][2]);
                            acc2[cp][3] = fma2(w[cp], pd[kr + 1][kq + 1], acc2[cp][3]);
                        }
                    }
            }
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) {
                const int co = cg * 8 + cc;
                float a[4], o;
#pragma unroll
                for (int j = 0; j < 4; ++j) { if (cc & 1) unpk2(acc2[cc >> 1][j], o, a[j]); else unpk2(acc2[cc >> 1][j], a[j], o); }
                const float s = __ldg(P.a2 + co), b = __ldg(P.b2 + co);
                const float m = fmaxf(fmaxf(fmaf(a[0], s, b), fmaf(a[1], s, b)), fmaxf(fmaf(a[2], s, b), fmaf(a[3], s, b)));
                feat[win * C4_FEAT + (co * C4_H2 + ph) * C4_W2 + pw] = fmaxf(m, 0.f);
            }
        }
    }
}

// fc1 (2400->128) + ReLU + fc2 (128->2) + argmax/softmax.  CTA tile: 64 windows x 128 outputs,
// thread tile 8 windows x 4 outputs, K streamed in 32-wide slabs.
constexpr int FC_BM = 64, FC_BK = 32, FC_THREADS = 256;

__global__ void __launch_bounds__(FC_THREADS)
cnn4dof_fc_kernel(Cnn4Dev P, const float* __restrict__ feat, const int* __restrict__ n_dev, long long n,
                  float* __restrict__ logits, long long* __restrict__ label, float* __restrict__ p_struct) {
    __shared__ __align__(16) float raw[FC_BM * 129];          // K-loop slabs, then reused for the fc1 activations
    float (*sA)[FC_BM + 4] = reinterpret_cast<float (*)[FC_BM + 4]>(raw);                       // [k][window]
    float (*sB)[128] = reinterpret_cast<float (*)[128]>(raw + FC_BK * (FC_BM + 4));            // [k][output]
    float (*sH)[129] = reinterpret_cast<float (*)[129]>(raw);
    const int tid = threadIdx.x;
    long long n_eff = n;
    if (n_dev) n_eff = min(n_eff, (long long)__ldg(n_dev));
    const long long n0 = (long long)blockIdx.x * FC_BM;
    if (n0 >= n_eff) return;
    const int nvalid = (int)min((long long)FC_BM, n_eff - n0);
    const int wg = tid >> 5;          // 8 window groups of 8
    const int og = tid & 31;          // 32 output groups of 4
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

    for (int k0 = 0; k0 < C4_FEAT; k0 += FC_BK) {
        __syncthreads();
        for (int i = tid; i < FC_BM * FC_BK; i += FC_THREADS) {
            const int w = i / FC_BK, k = i - w * FC_BK;
            sA[k][w] = (w < nvalid) ? __ldg(feat + (n0 + w) * C4_FEAT + k0 + k) : 0.f;
        }
        for (int i = tid; i < FC_BK * 32; i += FC_THREADS) {
            const int k = i >> 5, o4 = i & 31;
            *reinterpret_cast<float4*>(&sB[k][o4 * 4]) = __ldg(reinterpret_cast<const float4*>(P.fc1t + (size_t)(k0 + k) * 128) + o4);
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < FC_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[k][wg * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[k][wg * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&sB[k][og * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(a[i], b.x, acc[i][0]); acc[i][1] = fmaf(a[i], b.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], b.z, acc[i][2]); acc[i][3] = fmaf(a[i], b.w, acc[i][3]);
            }
        }
    }
    const float4 bb = __ldg(reinterpret_cast<const float4*>(P.fc1b) + og);
    __syncthreads();        // slabs are dead, sH aliases them
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sH[wg * 8 + i][og * 4 + 0] = fmaxf(acc[i][0] + bb.x, 0.f);
        sH[wg * 8 + i][og * 4 + 1] = fmaxf(acc[i][1] + bb.y, 0.f);
        sH[wg * 8 + i][og * 4 + 2] = fmaxf(acc[i][2] + bb.z, 0.f);
        sH[wg * 8 + i][og * 4 + 3] = fmaxf(acc[i][3] + bb.w, 0.f);
    }
    __syncthreads();
    if (tid < FC_BM && tid < nvalid) {
        float l0 = __ldg(P.fc2b + 0), l1 = __ldg(P.fc2b + 1);
        for (int o = 0; o < 128; ++o) {
            const float hv = sH[tid][o];
            l0 = fmaf(hv, __ldg(P.fc2w + o), l0);
            l1 = fmaf(hv, __ldg(P.fc2w + 128 + o), l1);
        }
        const long long w = n0 + tid;
        logits[w * 2 + 0] = l0;
        logits[w * 2 + 1] = l1;
        if (label) label[w] = (l1 > l0) ? 2 : 1;                 // argmax+1, tie -> class 0 -> label 1
        if (p_struct) {
            const float m = fmaxf(l0, l1);
            const float e0 = expf(l0 - m), e1 = expf(l1 - m);
            p_struct[w] = e1 / (e0 + e1);
        }
    }
}

// fold BatchNorm(eval) into scale/shift applied to the bias-free conv sum; transpose fc1
__global__ void cnn4dof_pack_kernel(const float* conv_b, const float* bn_w, const float* bn_b, const float* bn_m,
                                    const float* bn_v, float eps, int C, float* a, float* b) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        const float s = bn_w[c] / sqrtf(bn_v[c] + eps);
        a[c] = s;
        b[c] = fmaf(conv_b[c] - bn_m[c], s, bn_b[c]);
    }
}
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * cols; i += gridDim.x * blockDim.x) {
        const int r = i / cols, c = i - r * cols;
        out[(size_t)c * rows + r] = in[i];
    }
}

}  // namespace shm

using namespace shm;

static int cnn4_upload(shm_cnn4dof* h, const shm_cnn4dof_weights* w, cudaStream_t st) {
    // raw staging layout: conv_w1, conv_w2, fc1_w, fc1_b, fc2_w, fc2_b, then per block conv_b,bn_w,bn_b,bn_m,bn_v
    size_t off = 0;
    auto cp = [&](const float* src, size_t n, float** dst) -> int {
        if (!src) return SHM_ERR_ARG;
        *dst = h->raw + off;
        off += (n + 3) / 4 * 4;
        SHM_CUDA(cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDefault, st));
        return SHM_OK;
    };
    int rc;
    float *fc1w, *cb[2], *bw[2], *bb[2], *bm[2], *bv[2];
    const int C[2] = {C4_C1, C4_C2};
    float* d;
    if ((rc = cp(w->conv_w[0], C4_C1 * 2 * 9, &d))) return rc;
    SHM_CUDA(cudaMemcpyAsync(h->buf + h->o_w1, d, C4_C1 * 2 * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = cp(w->conv_w[1], C4_C2 * C4_C1 * 9, &d))) return rc;
    SHM_CUDA(cudaMemcpyAsync(h->buf + h->o_w2, d, C4_C2 * C4_C1 * 9 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = cp(w->fc1_w, (size_t)128 * C4_FEAT, &fc1w))) return rc;
    if ((rc = cp(w->fc1_b, 128, &d))) return rc;
    SHM_CUDA(cudaMemcpyAsync(h->buf + h->o_fc1b, d, 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = cp(w->fc2_w, 256, &d))) return rc;
    SHM_CUDA(cudaMemcpyAsync(h->buf + h->o_fc2w, d, 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if ((rc = cp(w->fc2_b, 2, &d))) return rc;
    SHM_CUDA(cudaMemcpyAsync(h->buf + h->o_fc2b, d, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int b = 0; b < 2; ++b) {
        if ((rc = cp(w->conv_b[b], C[b], &cb[b]))) return rc;
        if ((rc = cp(w->bn_w[b], C[b], &bw[b]))) return rc;
        if ((rc = cp(w->bn_b[b], C[b], &bb[b]))) return rc;
        if ((rc = cp(w->bn_mean[b], C[b], &bm[b]))) return rc;
        if ((rc = cp(w->bn_var[b], C[b], &bv[b]))) return rc;
    }
    const float eps = w->bn_eps > 0.f ? w->bn_eps : 1e-5f;
    cnn4dof_pack_kernel<<<1, 32, 0, st>>>(cb[0], bw[0], bb[0], bm[0], bv[0], eps, C4_C1, h->buf + h->o_a1, h->buf + h->o_b1);
    SHM_LAUNCH_CHECK();
    cnn4dof_pack_kernel<<<1, 32, 0, st>>>(cb[1], bw[1], bb[1], bm[1], bv[1], eps, C4_C2, h->buf + h->o_a2, h->buf + h->o_b2);
    SHM_LAUNCH_CHECK();
    transpose_kernel<<<148, 256, 0, st>>>(fc1w, 128, C4_FEAT, h->buf + h->o_fc1t);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

extern "C" int shm_cnn4dof_create(shm_cnn4dof** out, const shm_cnn4dof_weights* w, int device) {
    if (!out || !w) return SHM_ERR_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != SHM_OK) return rc;
    int prev = 0;
    SHM_CUDA(cudaGetDevice(&prev));
    SHM_CUDA(cudaSetDevice(device));
    shm_cnn4dof* h = new (std::nothrow) shm_cnn4dof();
    if (!h) return SHM_ERR_NOMEM;
    memset(h, 0, sizeof(*h));
    h->device = device;
    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return o; };
    h->o_w1 = take(C4_C1 * 2 * 9); h->o_a1 = take(C4_C1); h->o_b1 = take(C4_C1);
    h->o_w2 = take(C4_C2 * C4_C1 * 9); h->o_a2 = take(C4_C2); h->o_b2 = take(C4_C2);
    h->o_fc1t = take((size_t)C4_FEAT * 128); h->o_fc1b = take(128); h->o_fc2w = take(256); h->o_fc2b = take(4);
    h->total = off;
    h->raw_total = h->total + 1024;
    if (cudaMalloc(&h->buf, h->total * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&h->raw, h->raw_total * sizeof(float)) != cudaSuccess) {
        set_cuda_error(cudaGetLastError(), "cudaMalloc(cnn4dof)");
        shm_cnn4dof_destroy(h);
        cudaSetDevice(prev);
        return SHM_ERR_NOMEM;
    }
    rc = cnn4_upload(h, w, 0);
    if (rc == SHM_OK && cudaStreamSynchronize(0) != cudaSuccess) { set_cuda_error(cudaGetLastError(), "cnn4dof create"); rc = SHM_ERR_CUDA; }
    if (rc == SHM_OK) {
        const int smem = (2 * IN0_H * IN0_W + C4_C1 * P1_H * P1_W + C4_C1 * 2 * 9 + C4_C2 * C4_C1 * 9) * sizeof(float);
        if (cudaFuncSetAttribute(cnn4dof_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            set_cuda_error(cudaGetLastError(), "cudaFuncSetAttribute(cnn4dof_conv)");
            rc = SHM_ERR_CUDA;
        }
    }
    cudaSetDevice(prev);
    if (rc != SHM_OK) { shm_cnn4dof_destroy(h); return rc; }
    *out = h;
    return SHM_OK;
}

extern "C" int shm_cnn4dof_update_weights(shm_cnn4dof* h, const shm_cnn4dof_weights* w, void* stream) {
    if (!h || !w) return SHM_ERR_ARG;
    return cnn4_upload(h, w, static_cast<cudaStream_t>(stream));
}

extern "C" int shm_cnn4dof_destroy(shm_cnn4dof* h) {
    if (!h) return SHM_OK;
    if (h->buf) cudaFree(h->buf);
    if (h->raw) cudaFree(h->raw);
    if (h->feat) cudaFree(h->feat);
    delete h;
    return SHM_OK;
}

extern "C" int shm_cnn4dof_forward(shm_cnn4dof* h, const float* x, const int32_t* n_dev, int64_t n, float* logits,
                                   int64_t* label, float* p_struct, void* stream) {
    if (!h || n < 0 || (n > 0 && (!x || !logits))) return SHM_ERR_ARG;
    if (n == 0) return SHM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n > h->feat_cap) {
        // grow the feature scratch (stream-ordered free of the old one would need a sync; sizes only grow)
        SHM_CUDA(cudaStreamSynchronize(st));
        if (h->feat) cudaFree(h->feat);
        h->feat = nullptr; h->feat_cap = 0;
        if (cudaMalloc(&h->feat, (size_t)n * C4_FEAT * sizeof(float)) != cudaSuccess) {
            set_cuda_error(cudaGetLastError(), "cudaMalloc(cnn4dof features)");
            return SHM_ERR_NOMEM;
        }
        h->feat_cap = n;
    }
    Cnn4Dev P;
    P.w1 = h->buf + h->o_w1; P.a1 = h->buf + h->o_a1; P.b1 = h->buf + h->o_b1;
    P.w2 = h->buf + h->o_w2; P.a2 = h->buf + h->o_a2; P.b2 = h->buf + h->o_b2;
    P.fc1t = h->buf + h->o_fc1t; P.fc1b = h->buf + h->o_fc1b; P.fc2w = h->buf + h->o_fc2w; P.fc2b = h->buf + h->o_fc2b;
    const int smem = (2 * IN0_H * IN0_W + C4_C1 * P1_H * P1_W + C4_C1 * 2 * 9 + C4_C2 * C4_C1 * 9) * sizeof(float);
    const int sms = device_sm_count(h->device);
    const int grid1 = (int)min((long long)n, (long long)sms * 4 * 8);
    cnn4dof_conv_kernel<<<grid1, C4_THREADS, smem, st>>>(P, x, n_dev, n, h->feat);
    SHM_LAUNCH_CHECK();
    const long long grid2 = (n + FC_BM - 1) / FC_BM;
    cnn4dof_fc_kernel<<<(unsigned)grid2, FC_THREADS, 0, st>>>(P, h->feat, n_dev, n, logits, reinterpret_cast<long long*>(label), p_struct);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
