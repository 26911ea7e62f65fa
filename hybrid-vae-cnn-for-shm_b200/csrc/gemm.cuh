// fp32-grade strided contraction shared by the training paths:
//   C[m*ldc + n] (+)= sum_k A[m*a_ms + k*a_ks] * B[k*b_ks + n*b_ns] (+ bias1[n] + bias2[n])
// splitk = true: split along K with an atomicAdd epilogue (C must be zeroed by the caller).
// mode: SHM_GEMM_SIMT (fp32 FMA pipe, train.cu), SHM_GEMM_TC_F16X3 / SHM_GEMM_TC_BF16X3 (tcgen05, 3-pass hi/lo split, gemm_tc.cu;
// shapes that do not qualify fall back to the SIMT kernel).
#pragma once
#include "common.cuh"

namespace shm {
int sgemm_simt(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
               long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk);
// returns 1 when the tensor-core path does not apply (the caller then uses sgemm_simt)
int tc_gemm(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
            long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk, int mode);
inline int sgemm_mode(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns,
                      float* C, long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk, int mode) {
    if (M <= 0 || N <= 0 || K <= 0) return SHM_OK;
    const int r = tc_gemm(st, A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, splitk, mode);
    return r == 1 ? sgemm_simt(st, A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, splitk) : r;
}
inline int sgemm(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
                 long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk) {
    return sgemm_simt(st, A, a_ms, a_ks, B, b_ks, b_ns, C, ldc, M, N, K, bias1, bias2, splitk);
}
int adam_step(cudaStream_t st, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, int step, float lr,
              float beta1, float beta2, float eps, float weight_decay, int decoupled, float max_norm, float grad_scale, float* norm2);
}  // namespace shm
