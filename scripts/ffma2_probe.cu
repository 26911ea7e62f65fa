// Micro-probe: packed fp32 FMA (fma.rn.f32x2 -> FFMA2) issue rate per SM sub-partition on sm_100a, against scalar FFMA,
// alone and mixed with MUFU.EX2 (the LSTM cell update's mix: ~7 MUFU next to ~24 FP32 ops per cell).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma2_probe scripts/ffma2_probe.cu && /tmp/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ float ffma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
// MODE 0: 16 scalar FFMA   1: 8 FFMA2 (same flops)   2: 16 FFMA + 4 EX2   3: 8 FFMA2 + 4 EX2   4: 4 EX2
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
    float f[16]; unsigned long long p[8]; float a[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = 1.0f + 0.01f * i + threadIdx.x * 1e-6f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = make_float2(f[2 * i], f[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&v); }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = 0.001f * (threadIdx.x + i);
    const float2 m2 = make_float2(1.0001f, 0.9999f), c2 = make_float2(0.5f, 0.25f);
    const unsigned long long M = *reinterpret_cast<const unsigned long long*>(&m2), C = *reinterpret_cast<const unsigned long long*>(&c2);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = ffma1(f[i], 1.0001f, 0.5f);
        }
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], M, C);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = ex2(a[i]);
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&p[i]); s += v.x + v.y; }
#pragma unroll
    for (int i = 0; i < 4; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const char* names[5] = {"16 FFMA", "8 FFMA2", "16 FFMA + 4 EX2", "8 FFMA2 + 4 EX2", "4 EX2"};
    for (int threads : {128, 512}) {
        for (int mode = 0; mode < 5; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, threads>>>(out, cyc, iters);
                if (mode == 1) probe<1><<<148, threads>>>(out, cyc, iters);
                if (mode == 2) probe<2><<<148, threads>>>(out, cyc, iters);
                if (mode == 3) probe<3><<<148, threads>>>(out, cyc, iters);
                if (mode == 4) probe<4><<<148, threads>>>(out, cyc, iters);
                cudaDeviceSynchronize();
            }
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            const double warps_per_smsp = threads / 32.0 / 4.0;
            printf("threads %3d  %-18s : %8lld clk -> %.2f clk per iteration per warp-slot of one SMSP\n", threads, names[mode], h[0],
                   (double)h[0] / ((double)iters * warps_per_smsp));
        }
    }
    return 0;
}
