// Tensor-core engine (SHM_ENGINE_TC_BF16X3) of the fused LSTM-VAE scorer: interface.
// Implementation: vae_tc.cu (tcgen05.mma, 3-pass bf16 hi/lo split, fp32 accumulation in TMEM).
#pragma once
#include "vae_fp32.cuh"

// Role counters (cycles each warp role spends waiting / working; shm_vae_debug_counters) are compiled in only with -DSHM_TC_PROF
// (SHMFAST_PROF=1 python -m shmfast.build --force): clock64() is volatile, so the always-on form kept 8 x 64-bit counters in every
// thread's 80-register budget and ~20 extra instructions per chunk in the epilogue's hot loop.
#ifdef SHM_TC_PROF
#define TC_CLOCK() clock64()
#else
#define TC_CLOCK() 0LL
#endif

namespace shm {

struct VaeTcRaw {
    const float *enc_wih[SHM_MAX_L], *enc_whh[SHM_MAX_L], *enc_bih[SHM_MAX_L], *enc_bhh[SHM_MAX_L];
    const float *dec_wih[SHM_MAX_L], *dec_whh[SHM_MAX_L], *dec_bih[SHM_MAX_L], *dec_bhh[SHM_MAX_L];
};

struct VaeTc {
    void* wpack;       // bf16 hi/lo weight tiles in the UMMA smem layout, one block per (stack, layer) pass
    float* bias;       // pre-scaled packed biases [2L][H][4]
    void* scratch;     // per-CTA inter-layer h_t stream (grown on demand)
    size_t wpack_bytes, scratch_bytes;
    size_t pass_off[2 * SHM_MAX_L];
    int H, L, D;
    long long* dbg;    // optional profiling counters [#SM][8] (shm_vae_debug_counters)
};

bool vae_tc_supported(const shm_vae_cfg& cfg);
int vae_tc_alloc(VaeTc* tc, const shm_vae_cfg& cfg);
int vae_tc_pack(VaeTc* tc, const shm_vae_cfg& cfg, const VaeTcRaw& raw, cudaStream_t st);
void vae_tc_free(VaeTc* tc);
int vae_tc_score(VaeTc* tc, const VaeDev& P, const WinSrc& src, const VaeIO& io, cudaStream_t st);
bool vae_tc_can_rescore(const VaeTc* tc);      // io.mu_in / io.logvar_in (encoder skipped) is implemented by the single-tile kernels only

}  // namespace shm
