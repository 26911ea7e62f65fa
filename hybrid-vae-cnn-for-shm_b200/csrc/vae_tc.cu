// Tensor-core engine of the fused LSTM-VAE scorer (work in progress: not yet selectable).
#include "vae_tc.cuh"

namespace shm {

bool vae_tc_supported(const shm_vae_cfg&) { return false; }
int vae_tc_alloc(VaeTc*, const shm_vae_cfg&) { return SHM_ERR_UNSUPPORTED; }
int vae_tc_pack(VaeTc*, const shm_vae_cfg&, const VaeTcRaw&, cudaStream_t) { return SHM_ERR_UNSUPPORTED; }
void vae_tc_free(VaeTc* tc) {
    if (!tc) return;
    if (tc->wpack) cudaFree(tc->wpack);
    if (tc->bias) cudaFree(tc->bias);
    if (tc->scratch) cudaFree(tc->scratch);
    tc->wpack = nullptr; tc->bias = nullptr; tc->scratch = nullptr;
}
int vae_tc_score(VaeTc*, const VaeDev&, const WinSrc&, const VaeIO&, cudaStream_t) { return SHM_ERR_UNSUPPORTED; }

}  // namespace shm
