// Micro-probe: MUFU.EX2 / MUFU.RCP issue rate per SM sub-partition on sm_100a, alone and mixed with FFMA.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/mufu_probe scripts/mufu_probe.cu && /tmp/mufu_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp(float x) { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void probe(float* out, long long* cyc, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = 1.0f + 0.01f * i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]);
            if (MODE == 1) a[i] = rcp(a[i]);
            if (MODE == 2) { a[i] = ex2(a[i]);
#pragma unroll
                for (int k = 0; k < 4; ++k) f[i] = fmaf(f[i], 1.0001f, 0.5f); }
            if (MODE == 3) {
#pragma unroll
                for (int k = 0; k < 4; ++k) f[i] = fmaf(f[i], 1.0001f, 0.5f); }
            if (MODE == 4) { a[i] = ex2(a[i]);
#pragma unroll
                for (int k = 0; k < 8; ++k) f[i] = fmaf(f[i], 1.0001f, 0.5f); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const char* names[5] = {"ex2 only", "rcp only", "ex2 + 4 ffma", "4 ffma only", "ex2 + 8 ffma"};
    for (int threads : {128, 256, 512}) {
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
            const double per_group = (double)h[0] / ((double)iters * 8 * warps_per_smsp);   // clk per (one MUFU [+k FFMA]) warp-instr group per SMSP
            printf("threads %3d  %-14s : %8lld clk  -> %.2f clk per warp-level group per SMSP\n", threads, names[mode], h[0], per_group);
        }
    }
    return 0;
}
