// 1_DOF post-processing: stitch_windows (overlap-average) -> destandardize -> segment_rmse
// (1_DOF/Scripts/datasets.py:21-22,38-71; call site 04_test_seen_variants.py:296-311), fp64 like the
// reference.  One CTA per RMSE segment; each (row, channel) sums its covering windows in ascending
// window order -- the same order as the reference's `out[start:end] += windows[n]` loop, so the
// stitched series is bit-identical in fp64 -- and the squared error is reduced in the block.
#include "common.cuh"

namespace shm {

constexpr int ST_THREADS = 256;

__global__ void __launch_bounds__(ST_THREADS)
stitch_rmse_kernel(const float* __restrict__ recon, long long N, int T, int F, int stride, long long full_len,
                   const double* __restrict__ mean, const double* __restrict__ stdv, const float* __restrict__ y_true,
                   int segment_len, double* __restrict__ series_out, double* __restrict__ rmse_out) {
    __shared__ double s_part[ST_THREADS / 32];
    const long long i0 = (long long)blockIdx.x * segment_len;
    const long long i1 = min(i0 + segment_len, full_len);
    const long long items = (i1 - i0) * F;
    double sq = 0.0;
    for (long long it = threadIdx.x; it < items; it += ST_THREADS) {
        const long long r = i0 + it / F;
        const int f = (int)(it % F);
        // windows n with n*stride <= r < n*stride + T
        long long n_hi = r / stride;
        if (n_hi > N - 1) n_hi = N - 1;
        long long n_lo = (r - T + 1 + stride - 1) / stride;
        if (r - T + 1 <= 0) n_lo = 0;
        double acc = 0.0;
        long long cnt = 0;
        for (long long n = n_lo; n <= n_hi; ++n) {
            acc = __dadd_rn(acc, (double)recon[(n * T + (r - n * stride)) * F + f]);
            ++cnt;
        }
        const double c = cnt == 0 ? 1.0 : (double)cnt;
        const double xn = acc / c;
        const double x = __dadd_rn(__dmul_rn(xn, stdv[f]), mean[f]);
        if (series_out) series_out[r * F + f] = x;
        if (y_true) {
            const double e = x - (double)y_true[r * F + f];
            sq = fma(e, e, sq);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0 && rmse_out) {
        double s = 0.0;
        for (int w = 0; w < ST_THREADS / 32; ++w) s += s_part[w];
        rmse_out[blockIdx.x] = sqrt(s / (double)items);
    }
}

}  // namespace shm

extern "C" int shm_stitch_segment_rmse(const float* recon, int64_t N, int32_t T, int32_t F, int32_t stride,
                                       int64_t full_len, const double* mean, const double* std, const float* y_true,
                                       int32_t segment_len, double* series_out, double* rmse_out, void* stream) {
    using namespace shm;
    if (!recon || !mean || !std || N <= 0 || T <= 0 || F <= 0 || stride <= 0 || full_len <= 0 || segment_len <= 0)
        return SHM_ERR_ARG;
    if (rmse_out && !y_true) return SHM_ERR_ARG;
    int dev = 0;
    SHM_CUDA(cudaGetDevice(&dev));
    int rc = check_device(dev);
    if (rc != SHM_OK) return rc;
    const long long segs = (full_len + segment_len - 1) / segment_len;
    stitch_rmse_kernel<<<(unsigned)segs, ST_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        recon, N, T, F, stride, full_len, mean, std, y_true, segment_len, series_out, rmse_out);
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}
