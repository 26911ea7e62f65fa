// C ABI for the LSTM-VAE scorer: handle management, device-side weight repacking, engine dispatch.
#include <new>
#include <vector>
#include "vae_fp32.cuh"
#include "vae_tc.cuh"

struct shm_vae {
    shm_vae_cfg cfg;
    int device;
    int engine;
    float* arena;          // raw state_dict tensors, copied from the caller
    float* packed;         // repacked LSTM weights / biases for the fp32 engine
    size_t arena_floats, packed_floats;
    shm::VaeDev dev;
    // raw tensor offsets inside the arena
    size_t o_enc_wih[SHM_MAX_L], o_enc_whh[SHM_MAX_L], o_enc_bih[SHM_MAX_L], o_enc_bhh[SHM_MAX_L];
    size_t o_dec_wih[SHM_MAX_L], o_dec_whh[SHM_MAX_L], o_dec_bih[SHM_MAX_L], o_dec_bhh[SHM_MAX_L];
    size_t o_ln_w, o_ln_b, o_mu_w, o_mu_b, o_lv_w, o_lv_b, o_l2h_w, o_l2h_b, o_out_w, o_out_b;
    size_t p_encW, p_decW, p_encB[SHM_MAX_L], p_decB[SHM_MAX_L];
    shm::VaeTc tc;         // tensor-core engine state (allocated only when selected)
};

namespace shm {

// Wp[k][ (u%4)*H + (u/4)*4 + g ] = [W_ih | W_hh][g*H+u][k], rows k in [Kin, Kin_pad) zero.
__global__ void pack_lstm_layer_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                       const float* __restrict__ b_ih, const float* __restrict__ b_hh, int Kin,
                                       int Kin_pad, int H, float* __restrict__ Wp, float* __restrict__ bp) {
    const int G4 = 4 * H;
    const int K = Kin_pad + H;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * G4; i += gridDim.x * blockDim.x) {
        const int k = i / G4;
        const int col = i - k * G4;
        const int j = col / H;
        const int r = col - j * H;
        const int ug = r >> 2, g = r & 3;
        const int u = ug * 4 + j;
        const int row = g * H + u;
        float v = 0.f;
        if (k < Kin) v = w_ih[(size_t)row * Kin + k];
        else if (k >= Kin_pad) v = w_hh[(size_t)row * H + (k - Kin_pad)];
        Wp[i] = v;
        if (k == 0) bp[col] = b_ih[row] + b_hh[row];
    }
}

static inline int pad16(int x) { return (x + 15) / 16 * 16; }

}  // namespace shm

using namespace shm;

static int vae_upload(shm_vae* h, const shm_vae_weights* w, cudaStream_t st) {
    const int D = h->cfg.D, H = h->cfg.H, Z = h->cfg.Z, L = h->cfg.L;
    auto cp = [&](size_t off, const float* src, size_t n) -> int {
        if (!src) return SHM_ERR_ARG;
        SHM_CUDA(cudaMemcpyAsync(h->arena + off, src, n * sizeof(float), cudaMemcpyDefault, st));
        return SHM_OK;
    };
    int rc;
    for (int l = 0; l < L; ++l) {
        const size_t ein = (l == 0) ? D : H;
        if ((rc = cp(h->o_enc_wih[l], w->enc_w_ih[l], (size_t)4 * H * ein))) return rc;
        if ((rc = cp(h->o_enc_whh[l], w->enc_w_hh[l], (size_t)4 * H * H))) return rc;
        if ((rc = cp(h->o_enc_bih[l], w->enc_b_ih[l], (size_t)4 * H))) return rc;
        if ((rc = cp(h->o_enc_bhh[l], w->enc_b_hh[l], (size_t)4 * H))) return rc;
        if ((rc = cp(h->o_dec_wih[l], w->dec_w_ih[l], (size_t)4 * H * H))) return rc;
        if ((rc = cp(h->o_dec_whh[l], w->dec_w_hh[l], (size_t)4 * H * H))) return rc;
        if ((rc = cp(h->o_dec_bih[l], w->dec_b_ih[l], (size_t)4 * H))) return rc;
        if ((rc = cp(h->o_dec_bhh[l], w->dec_b_hh[l], (size_t)4 * H))) return rc;
    }
    if (h->cfg.has_ln) {
        if ((rc = cp(h->o_ln_w, w->ln_w, H))) return rc;
        if ((rc = cp(h->o_ln_b, w->ln_b, H))) return rc;
    }
    if ((rc = cp(h->o_mu_w, w->fc_mu_w, (size_t)Z * H))) return rc;
    if ((rc = cp(h->o_mu_b, w->fc_mu_b, Z))) return rc;
    if ((rc = cp(h->o_lv_w, w->fc_lv_w, (size_t)Z * H))) return rc;
    if ((rc = cp(h->o_lv_b, w->fc_lv_b, Z))) return rc;
    if ((rc = cp(h->o_l2h_w, w->l2h_w, (size_t)H * Z))) return rc;
    if ((rc = cp(h->o_l2h_b, w->l2h_b, H))) return rc;
    if ((rc = cp(h->o_out_w, w->out_w, (size_t)D * H))) return rc;
    if ((rc = cp(h->o_out_b, w->out_b, D))) return rc;

    size_t erow = 0, drow = 0;
    for (int l = 0; l < L; ++l) {
        const int ekin = (l == 0) ? D : H;
        const int ekp = (l == 0) ? 16 : H;
        pack_lstm_layer_kernel<<<64, 256, 0, st>>>(h->arena + h->o_enc_wih[l], h->arena + h->o_enc_whh[l],
                                                   h->arena + h->o_enc_bih[l], h->arena + h->o_enc_bhh[l], ekin, ekp, H,
                                                   h->packed + h->p_encW + erow * 4 * H, h->packed + h->p_encB[l]);
        SHM_LAUNCH_CHECK();
        erow += ekp + H;
        pack_lstm_layer_kernel<<<64, 256, 0, st>>>(h->arena + h->o_dec_wih[l], h->arena + h->o_dec_whh[l],
                                                   h->arena + h->o_dec_bih[l], h->arena + h->o_dec_bhh[l], H, H, H,
                                                   h->packed + h->p_decW + drow * 4 * H, h->packed + h->p_decB[l]);
        SHM_LAUNCH_CHECK();
        drow += 2 * H;
    }
    if (h->engine == SHM_ENGINE_TC_BF16X3) {
        VaeTcRaw raw;
        for (int l = 0; l < L; ++l) {
            raw.enc_wih[l] = h->arena + h->o_enc_wih[l]; raw.enc_whh[l] = h->arena + h->o_enc_whh[l];
            raw.enc_bih[l] = h->arena + h->o_enc_bih[l]; raw.enc_bhh[l] = h->arena + h->o_enc_bhh[l];
            raw.dec_wih[l] = h->arena + h->o_dec_wih[l]; raw.dec_whh[l] = h->arena + h->o_dec_whh[l];
            raw.dec_bih[l] = h->arena + h->o_dec_bih[l]; raw.dec_bhh[l] = h->arena + h->o_dec_bhh[l];
        }
        if ((rc = vae_tc_pack(&h->tc, h->cfg, raw, st))) return rc;
    }
    return SHM_OK;
}

extern "C" int shm_vae_create(shm_vae** out, const shm_vae_cfg* cfg, const shm_vae_weights* w, int device) {
    if (!out || !cfg || !w) return SHM_ERR_ARG;
    *out = nullptr;
    int rc = check_device(device);
    if (rc != SHM_OK) return rc;
    const int D = cfg->D, H = cfg->H, Z = cfg->Z, L = cfg->L;
    if (D < 1 || Z < 1 || L < 1) return SHM_ERR_ARG;
    if (!(H == 32 || H == 64 || H == 128) || L > SHM_MAX_L || D > SHM_MAX_D || Z > VAE_MAX_Z) return SHM_ERR_UNSUPPORTED;
    int engine = cfg->engine;
    if (engine == SHM_ENGINE_AUTO) engine = vae_tc_supported(*cfg) ? SHM_ENGINE_TC_BF16X3 : SHM_ENGINE_FP32;
    if (engine == SHM_ENGINE_TC_BF16X3 && !vae_tc_supported(*cfg)) return SHM_ERR_UNSUPPORTED;
    if (engine != SHM_ENGINE_FP32 && engine != SHM_ENGINE_TC_BF16X3) return SHM_ERR_ARG;

    int prev = 0;
    SHM_CUDA(cudaGetDevice(&prev));
    SHM_CUDA(cudaSetDevice(device));
    shm_vae* h = new (std::nothrow) shm_vae();
    if (!h) { cudaSetDevice(prev); return SHM_ERR_NOMEM; }
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg; h->device = device; h->engine = engine;
    if (h->cfg.ln_eps <= 0.f) h->cfg.ln_eps = 1e-5f;

    size_t off = 0;
    auto take = [&](size_t n) { size_t o = off; off += (n + 3) / 4 * 4; return o; };   // keep 16-byte alignment
    for (int l = 0; l < L; ++l) {
        h->o_enc_wih[l] = take((size_t)4 * H * (l == 0 ? D : H)); h->o_enc_whh[l] = take((size_t)4 * H * H);
        h->o_enc_bih[l] = take(4 * H); h->o_enc_bhh[l] = take(4 * H);
        h->o_dec_wih[l] = take((size_t)4 * H * H); h->o_dec_whh[l] = take((size_t)4 * H * H);
        h->o_dec_bih[l] = take(4 * H); h->o_dec_bhh[l] = take(4 * H);
    }
    h->o_ln_w = take(H); h->o_ln_b = take(H);
    h->o_mu_w = take((size_t)Z * H); h->o_mu_b = take(Z); h->o_lv_w = take((size_t)Z * H); h->o_lv_b = take(Z);
    h->o_l2h_w = take((size_t)H * Z); h->o_l2h_b = take(H); h->o_out_w = take((size_t)D * H); h->o_out_b = take(D);
    h->arena_floats = off;

    size_t poff = 0;
    auto ptake = [&](size_t n) { size_t o = poff; poff += (n + 3) / 4 * 4; return o; };
    size_t encK = 0, decK = 0;
    for (int l = 0; l < L; ++l) { encK += (l == 0 ? 16 : H) + H; decK += 2 * H; }
    h->p_encW = ptake(encK * 4 * H);
    h->p_decW = ptake(decK * 4 * H);
    for (int l = 0; l < L; ++l) { h->p_encB[l] = ptake(4 * H); h->p_decB[l] = ptake(4 * H); }
    h->packed_floats = poff;

    cudaError_t e1 = cudaMalloc(&h->arena, h->arena_floats * sizeof(float));
    cudaError_t e2 = cudaMalloc(&h->packed, h->packed_floats * sizeof(float));
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        set_cuda_error(e1 != cudaSuccess ? e1 : e2, "cudaMalloc(vae weights)");
        shm_vae_destroy(h);
        cudaSetDevice(prev);
        return SHM_ERR_NOMEM;
    }
    if (engine == SHM_ENGINE_TC_BF16X3 && (rc = vae_tc_alloc(&h->tc, h->cfg)) != SHM_OK) {
        shm_vae_destroy(h);
        cudaSetDevice(prev);
        return rc;
    }

    VaeDev& d = h->dev;
    d.D = D; d.H = H; d.Z = Z; d.L = L; d.has_ln = cfg->has_ln; d.Dpad = 16; d.ln_eps = h->cfg.ln_eps;
    d.encW = h->packed + h->p_encW; d.decW = h->packed + h->p_decW;
    for (int l = 0; l < L; ++l) {
        d.encB[l] = h->packed + h->p_encB[l]; d.decB[l] = h->packed + h->p_decB[l];
        d.encKin[l] = (l == 0) ? 16 : H; d.encK[l] = d.encKin[l] + H;
        d.decKin[l] = H; d.decK[l] = 2 * H;
    }
    d.ln_w = h->arena + h->o_ln_w; d.ln_b = h->arena + h->o_ln_b;
    d.mu_w = h->arena + h->o_mu_w; d.mu_b = h->arena + h->o_mu_b;
    d.lv_w = h->arena + h->o_lv_w; d.lv_b = h->arena + h->o_lv_b;
    d.l2h_w = h->arena + h->o_l2h_w; d.l2h_b = h->arena + h->o_l2h_b;
    d.out_w = h->arena + h->o_out_w; d.out_b = h->arena + h->o_out_b;

    rc = vae_upload(h, w, 0);
    if (rc == SHM_OK) {
        cudaError_t e = cudaStreamSynchronize(0);
        if (e != cudaSuccess) { set_cuda_error(e, "vae create sync"); rc = SHM_ERR_CUDA; }
    }
    // opt in to the large dynamic shared memory footprints once per device
    if (rc == SHM_OK) {
        cudaError_t e = cudaSuccess;
        if (H == 128) e = cudaFuncSetAttribute(vae_score_fp32_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VaeSmem<128, 64>::bytes);
        if (H == 64) e = cudaFuncSetAttribute(vae_score_fp32_kernel<64, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VaeSmem<64, 128>::bytes);
        if (H == 32) e = cudaFuncSetAttribute(vae_score_fp32_kernel<32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VaeSmem<32, 128>::bytes);
        if (e != cudaSuccess) { set_cuda_error(e, "cudaFuncSetAttribute(vae_score_fp32)"); rc = SHM_ERR_CUDA; }
    }
    cudaSetDevice(prev);
    if (rc != SHM_OK) { shm_vae_destroy(h); return rc; }
    *out = h;
    return SHM_OK;
}

extern "C" int shm_vae_update_weights(shm_vae* h, const shm_vae_weights* w, void* stream) {
    if (!h || !w) return SHM_ERR_ARG;
    return vae_upload(h, w, static_cast<cudaStream_t>(stream));
}

extern "C" int shm_vae_destroy(shm_vae* h) {
    if (!h) return SHM_OK;
    if (h->arena) cudaFree(h->arena);
    if (h->packed) cudaFree(h->packed);
    vae_tc_free(&h->tc);
    delete h;
    return SHM_OK;
}

extern "C" int shm_vae_engine(const shm_vae* h) { return h ? h->engine : SHM_ERR_ARG; }

extern "C" int shm_vae_get_cfg(const shm_vae* h, shm_vae_cfg* out) {
    if (!h || !out) return SHM_ERR_ARG;
    *out = h->cfg;
    out->engine = h->engine;
    return SHM_OK;
}

extern "C" int shm_vae_debug_counters(shm_vae* h, long long* out_host, int n) {
    if (!h || n < 0) return SHM_ERR_ARG;
    if (h->engine != SHM_ENGINE_TC_BF16X3) return SHM_ERR_UNSUPPORTED;
    const size_t total = (size_t)256 * 3 * 8;
    if (!h->tc.dbg) {                       // first call switches the counters on
        SHM_CUDA(cudaMalloc(&h->tc.dbg, total * sizeof(long long)));
        SHM_CUDA(cudaMemset(h->tc.dbg, 0, total * sizeof(long long)));
    }
    if (out_host && n > 0) {
        SHM_CUDA(cudaDeviceSynchronize());
        SHM_CUDA(cudaMemcpy(out_host, h->tc.dbg, (size_t)(n < (int)total ? n : (int)total) * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    return SHM_OK;
}

static int vae_launch(shm_vae* h, const shm::WinSrc& src, const shm::VaeIO& io, cudaStream_t st) {
    if (h->engine == SHM_ENGINE_TC_BF16X3 && !io.z_in) return vae_tc_score(&h->tc, h->dev, src, io, st);
    const int H = h->cfg.H;
    const long long n = io.n;
    if (H == 128) {
        using S = VaeSmem<128, 64>;
        vae_score_fp32_kernel<128, 64><<<(unsigned)((n + 63) / 64), S::NT, S::bytes, st>>>(h->dev, src, io);
    } else if (H == 64) {
        using S = VaeSmem<64, 128>;
        vae_score_fp32_kernel<64, 128><<<(unsigned)((n + 127) / 128), S::NT, S::bytes, st>>>(h->dev, src, io);
    } else {
        using S = VaeSmem<32, 128>;
        vae_score_fp32_kernel<32, 128><<<(unsigned)((n + 127) / 128), S::NT, S::bytes, st>>>(h->dev, src, io);
    }
    SHM_LAUNCH_CHECK();
    return SHM_OK;
}

extern "C" int shm_vae_decode(shm_vae* h, const float* z, int64_t n, int32_t T, float* recon, void* stream) {
    if (!h || n < 0 || T <= 0 || (n > 0 && (!z || !recon))) return SHM_ERR_ARG;
    if (n == 0) return SHM_OK;
    WinSrc src;
    memset(&src, 0, sizeof(src));
    src.T = T; src.D = h->cfg.D;
    VaeIO io;
    memset(&io, 0, sizeof(io));
    io.z_in = z; io.n = n; io.recon = recon;
    return vae_launch(h, src, io, static_cast<cudaStream_t>(stream));
}

extern "C" int shm_vae_score(shm_vae* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev,
                             const float* eps, int64_t n, float* score, float* mu, float* logvar, float* recon,
                             float* cnn_in, void* stream) {
    if (!h || !src_host || n < 0) return SHM_ERR_ARG;
    WinSrc src;
    int rc = make_winsrc(src_host, &src);
    if (rc != SHM_OK) return rc;
    if (src.D != h->cfg.D) return SHM_ERR_ARG;
    if (n == 0) return SHM_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaeIO io;
    memset(&io, 0, sizeof(io));
    io.idx = idx; io.n_dev = n_dev; io.eps = eps; io.n = n;
    io.score = score; io.mu = mu; io.logvar = logvar; io.recon = recon; io.cnn_in = cnn_in;

    return vae_launch(h, src, io, st);
}

extern "C" int shm_vae_rescore(shm_vae* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev, const float* mu_all,
                               const float* logvar_all, const float* eps, int64_t n, float* score, float* recon, float* cnn_in,
                               void* stream) {
    if (!h || !src_host || n < 0 || !mu_all || !logvar_all) return SHM_ERR_ARG;
    if (h->engine != SHM_ENGINE_TC_BF16X3 || !vae_tc_can_rescore(&h->tc)) return SHM_ERR_UNSUPPORTED;
    WinSrc src;
    int rc = make_winsrc(src_host, &src);
    if (rc != SHM_OK) return rc;
    if (src.D != h->cfg.D) return SHM_ERR_ARG;
    if (n == 0) return SHM_OK;
    VaeIO io;
    memset(&io, 0, sizeof(io));
    io.idx = idx; io.n_dev = n_dev; io.eps = eps; io.n = n; io.mu_in = mu_all; io.logvar_in = logvar_all;
    io.score = score; io.recon = recon; io.cnn_in = cnn_in;
    return vae_tc_score(&h->tc, h->dev, src, io, static_cast<cudaStream_t>(stream));
}
