// fp32 strided contraction shared by the training paths (defined in train.cu):
//   C[m*ldc + n] (+)= sum_k A[m*a_ms + k*a_ks] * B[k*b_ks + n*b_ns] (+ bias1[n] + bias2[n])
// splitk = true: split along K with an atomicAdd epilogue (C must be zeroed by the caller).
#pragma once
#include "common.cuh"

namespace shm {
int sgemm(cudaStream_t st, const float* A, long long a_ms, long long a_ks, const float* B, long long b_ks, long long b_ns, float* C,
          long long ldc, int M, int N, int K, const float* bias1, const float* bias2, bool splitk);
int adam_step(cudaStream_t st, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, int step, float lr,
              float beta1, float beta2, float eps, float weight_decay, int decoupled, float max_norm, float grad_scale, float* norm2);
}  // namespace shm
