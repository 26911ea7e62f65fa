// Tensor-core path of the openLAB attribution CNN (see cnnol_tc.cu).
#pragma once
#include "common.cuh"

namespace shm {

struct CnnOlDev;   // cnnol.cu

struct CnnOlTc {
    unsigned char* wimg[3];   // fp16 hi|lo UMMA K-major images of conv blocks 2..4, one stage per (ci-chunk, dt, df)
    float* wscale;            // device float[3][2]: power-of-two weight scale and its inverse per block
    int chunk;                // windows the workspace is sized for (0 = not allocated)
    float* raw[2];            // ping-pong raw convolution outputs, channels-last fp32 [chunk][H][4][C] (25,600 floats / window)
    unsigned char* staged;    // activated + pooled + split A-operand images of the current block
    double* stats;            // [4][chunk][8][2] GroupNorm sum / sum of squares
    float2* scsh;             // [chunk][256] per (window, channel) scale / shift of the fused GroupNorm affine
    int nsm;
};

int cnnol_tc_init(CnnOlTc* t, int device);
void cnnol_tc_free(CnnOlTc* t);
// raw_w[b]: device pointers to conv{b}.weight in the reference layout [Cout][Cin][kt][3], b = 1..3 used
int cnnol_tc_pack(CnnOlTc* t, const float* const raw_w[4], cudaStream_t st);
int cnnol_tc_forward(CnnOlTc* t, const float* const conv1_w /*reference layout [32][1][7][3]*/, const float* const bias[4],
                     const float* const gn_w[4], const float* const gn_b[4], const float* fc1t, const float* fc1b,
                     const float* fc2w, const float* fc2b, float gn_eps, const WinSrc& src, const int* idx, const int* n_dev,
                     long long n, float* logits, double* prob, cudaStream_t st);

}  // namespace shm
