// Shared helpers for libshmfast (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/shmfast.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libshmfast is written for sm_100a (B200) only"
#endif

namespace shm {

void set_cuda_error(cudaError_t e, const char* where);

#define SHM_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) {                                    \
            shm::set_cuda_error(_e, #call);                         \
            return SHM_ERR_CUDA;                                    \
        }                                                           \
    } while (0)

#define SHM_LAUNCH_CHECK()                                          \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) {                                    \
            shm::set_cuda_error(_e, "kernel launch");               \
            return SHM_ERR_CUDA;                                    \
        }                                                           \
    } while (0)

int device_sm_count(int device);
int check_device(int device);    // SHM_OK iff cc 10.x

// Device-side copy of shm_window_src (passed by value as a kernel parameter).
struct WinSrc {
    const float* base;
    long long win_stride;
    long long row_stride;
    int T, D;
    int chan[SHM_MAX_D];
    int normalize, nan_to_zero;
    float clip;
    float mean[SHM_MAX_D];
    float std[SHM_MAX_D];
    float rstd[SHM_MAX_D];      // 1/std, for the division-free transform of the tensor-core scorer
};

inline int make_winsrc(const shm_window_src* s, WinSrc* w) {
    if (!s || !s->base || s->T <= 0 || s->D <= 0 || s->D > SHM_MAX_D) return SHM_ERR_ARG;
    w->base = s->base; w->win_stride = s->win_stride; w->row_stride = s->row_stride;
    w->T = s->T; w->D = s->D; w->normalize = s->normalize; w->nan_to_zero = s->nan_to_zero; w->clip = s->clip;
    for (int i = 0; i < SHM_MAX_D; ++i) {
        w->chan[i] = i < s->D ? s->chan[i] : 0;
        w->mean[i] = i < s->D ? s->mean[i] : 0.f;
        w->std[i] = i < s->D ? s->std[i] : 1.f;
        w->rstd[i] = (float)(1.0 / (double)w->std[i]);
        if (i < s->D && s->chan[i] < 0) return SHM_ERR_ARG;
    }
    return SHM_OK;
}

#ifdef __CUDACC__
// x[win][t][d] with the reference's exact normalisation arithmetic (see shmfast.h).
__device__ __forceinline__ float win_transform(const WinSrc& s, float raw, int d) {
    float x = raw;
    if (s.normalize) x = __fdiv_rn(__fsub_rn(raw, s.mean[d]), s.std[d]);
    if (s.clip > 0.f) x = (x < -s.clip) ? -s.clip : ((x > s.clip) ? s.clip : x);   // NaN passes, as np.clip
    if (s.nan_to_zero && !isfinite(x)) x = 0.f;
    return x;
}

// Same transform without the IEEE-divide subroutine: q = (x-m)*rstd refined by one Newton step (the result is
// the correctly rounded quotient except in rare last-bit cases).  Used where the value only feeds the
// tolerance-checked score, never where windows are handed back to the caller.
__device__ __forceinline__ float win_transform_fast(const WinSrc& s, float raw, int d) {
    float x = raw;
    if (s.normalize) {
        const float num = raw - s.mean[d];
        const float q = num * s.rstd[d];
        const float r = fmaf(-q, s.std[d], num);
        const float q2 = fmaf(r, s.rstd[d], q);
        x = isfinite(q2) ? q2 : q;
    }
    if (s.clip > 0.f) x = (x < -s.clip) ? -s.clip : ((x > s.clip) ? s.clip : x);
    if (s.nan_to_zero && !isfinite(x)) x = 0.f;
    return x;
}

__device__ __forceinline__ float win_fetch(const WinSrc& s, long long win, int t, int d) {
    const float raw = __ldg(s.base + win * s.win_stride + (long long)t * s.row_stride + s.chan[d]);
    return win_transform(s, raw, d);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif

}  // namespace shm
