/*
 * shmfast.h -- C ABI of libshmfast.so: B200 (sm_100a) hybrid window scoring for structural-health
 * monitoring.  Drop-in boundary for ONE hot path of Ogunleyemma1/Hybrid-VAE-CNN-for-SHM:
 * window gather+normalise -> LSTM-VAE gate -> per-window reconstruction MSE -> stored-threshold
 * compare + ascending compaction -> CNN attribution on the flagged windows.
 *
 * The reference has no FFI: its seam is the Python class surface of `Models/` and the inner loops of
 * its numbered scripts.  Each entry point below cites the reference code it
 * replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding and the
 * `Models/` shim a maintainer adds.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer; the library owns only its handles (repacked weights, workspace);
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default);
 *   - return value: 0 = SHM_OK, negative = error (shm_strerror); never throws, never aborts;
 *   - there is NO CPU fallback: without a CUDA device of compute capability 10.x every compute
 *     entry point returns SHM_ERR_DEVICE.
 */
#ifndef SHMFAST_H
#define SHMFAST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHM_MAX_D 16      /* max channels per window fed to the VAE / CNN gather            */
#define SHM_MAX_L 2       /* max stacked LSTM layers (reference uses 1 and 2)                */

enum {
    SHM_OK = 0,
    SHM_ERR_ARG = -1,         /* null pointer / negative size / inconsistent shapes          */
    SHM_ERR_UNSUPPORTED = -2, /* configuration outside the supported set (see shm_vae_cfg)   */
    SHM_ERR_CUDA = -3,        /* a CUDA runtime call failed (shm_last_cuda_error)            */
    SHM_ERR_DEVICE = -4,      /* no CUDA device, or not compute capability 10.x (B200)       */
    SHM_ERR_NOMEM = -5
};

const char* shm_strerror(int code);
const char* shm_last_cuda_error(void);        /* text of the last CUDA failure on this thread */
int shm_version(void);                        /* 10000*major + 100*minor + patch              */
int shm_device_check(int device);             /* SHM_OK iff `device` is an sm_100 part        */

/* ---------------------------------------------------------------------------------------------
 * Window source: how x[n][t][d] is produced.  One description serves both the materialising
 * gather kernel and the fused scorers (which then never write the windows to HBM):
 *
 *     raw  = base[ win(n) * win_stride + t * row_stride + chan[d] ]
 *     x    = normalize ? (raw - mean[d]) / std[d] : raw          (IEEE fp32 divide, as NumPy)
 *     x    = clip > 0  ? min(max(x, -clip), clip) : x            (NaN passes through, as np.clip)
 *     x    = nan_to_zero && !isfinite(x) ? 0 : x                 (np.nan_to_num(nan=0,posinf=0,neginf=0))
 *
 * with win(n) = idx ? idx[n] : n.  Replaces make_windows + normalize_windows
 * (4DOF/Scripts/06_test_full_pipeline.py:106-126, copies in 03/04/05), windowize_2d + standardize +
 * channel select (openLAB feature_utils.py:130-152, 10_test_hybrid_pipeline.py:233-237,351) and
 * standardize + make_windows (1_DOF/Scripts/datasets.py:17-35).  The per-stage std guards
 * (std==0 -> 1e-6; sd<1e-12 -> 1; sd<1e-8 -> 1) are applied by the caller to `std` beforehand,
 * exactly where the reference applies them (load_stats / fit_mean_std).
 *
 *   series [rows, d_all], stride s : win_stride = s*d_all, row_stride = d_all, chan = column ids
 *   windows [N, T, D_all]          : win_stride = T*D_all, row_stride = D_all
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* base;
    int64_t win_stride;
    int64_t row_stride;
    int32_t T;
    int32_t D;
    int32_t chan[SHM_MAX_D];
    int32_t normalize;
    int32_t nan_to_zero;
    float clip;                 /* <= 0 : no clipping */
    float mean[SHM_MAX_D];
    float std[SHM_MAX_D];
} shm_window_src;

/* Materialise out[n, t, d] (C-contiguous [N,T,D] fp32) for n in [0,N).  `idx` (optional, device,
 * int32[N]) selects source windows (np fancy indexing Z[sel], 06_test_full_pipeline.py:362). */
int shm_window_normalize(const shm_window_src* src_host, const int32_t* idx, int64_t N, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSTM-VAE (TemporalVAE / VAE): 4DOF/Scripts/Models/temporal_vae.py:14-77,
 * 1_DOF/Scripts/Models/temporal_vae.py:8-58, openLAB Codes/Models/temporal_vae_model.py:4-66.
 * Supported: H in {32, 64, 128}, 1 <= L <= SHM_MAX_L, 1 <= D <= SHM_MAX_D, 1 <= Z <= 16.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t D, H, Z, L;
    int32_t has_ln;             /* LayerNorm on h_n[-1] (4DOF, openLAB) or not (1_DOF)         */
    float ln_eps;               /* 1e-5                                                         */
    int32_t engine;             /* SHM_ENGINE_*                                                 */
} shm_vae_cfg;

enum {
    SHM_ENGINE_AUTO = 0,        /* fastest engine that meets the 1e-4 score tolerance          */
    SHM_ENGINE_FP32 = 1,        /* fp32 FMA recurrence (any supported shape)                   */
    SHM_ENGINE_TC_BF16X3 = 2    /* tcgen05 tensor cores, 3-pass bf16 hi/lo split, fp32 accum   */
};

/* state_dict tensors, reference key names in comments; all fp32, row-major, device or host
 * pointers (copied + repacked at create/update; not referenced afterwards). */
typedef struct {
    const float* enc_w_ih[SHM_MAX_L];   /* encoder_lstm.weight_ih_l{k} [4H, D or H], gate rows i,f,g,o */
    const float* enc_w_hh[SHM_MAX_L];   /* encoder_lstm.weight_hh_l{k} [4H, H]  */
    const float* enc_b_ih[SHM_MAX_L];   /* encoder_lstm.bias_ih_l{k}   [4H]     */
    const float* enc_b_hh[SHM_MAX_L];   /* encoder_lstm.bias_hh_l{k}   [4H]     */
    const float* ln_w;                  /* layer_norm.weight [H] (NULL if !has_ln) */
    const float* ln_b;                  /* layer_norm.bias   [H] */
    const float* fc_mu_w;               /* fc_mu.weight [Z,H] */
    const float* fc_mu_b;               /* fc_mu.bias   [Z]   */
    const float* fc_lv_w;               /* fc_logvar.weight [Z,H] */
    const float* fc_lv_b;               /* fc_logvar.bias   [Z]   */
    const float* l2h_w;                 /* fc_latent_to_hidden.weight [H,Z] */
    const float* l2h_b;                 /* fc_latent_to_hidden.bias   [H]   */
    const float* dec_w_ih[SHM_MAX_L];   /* decoder_lstm.weight_ih_l{k} [4H, H] */
    const float* dec_w_hh[SHM_MAX_L];   /* decoder_lstm.weight_hh_l{k} [4H, H] */
    const float* dec_b_ih[SHM_MAX_L];   /* decoder_lstm.bias_ih_l{k}   [4H]    */
    const float* dec_b_hh[SHM_MAX_L];   /* decoder_lstm.bias_hh_l{k}   [4H]    */
    const float* out_w;                 /* output_layer.weight [D,H] */
    const float* out_b;                 /* output_layer.bias   [D]   */
} shm_vae_weights;

typedef struct shm_vae shm_vae;

int shm_vae_create(shm_vae** out, const shm_vae_cfg* cfg_host, const shm_vae_weights* w_host, int device);
/* refresh after load_state_dict()/optimizer.step(); asynchronous on `stream` */
int shm_vae_update_weights(shm_vae* h, const shm_vae_weights* w_host, void* stream);
int shm_vae_destroy(shm_vae* h);
int shm_vae_engine(const shm_vae* h);    /* the engine actually selected */
int shm_vae_get_cfg(const shm_vae* h, shm_vae_cfg* out_host);   /* the handle's shape; engine = the one selected */
/* Profiling aid (tensor-core engine): the first call enables per-CTA cycle counters of the MMA-issuer warp,
 * later calls copy them out: out_host[cta*8 + {0: wait weights, 1: wait input, 2: wait accumulator drain,
 * 3: wait h_t, 4..7: total cycles of pass 0..3}], accumulated over launches.  The counters are compiled into the kernels only
 * when the library is built with -DSHM_TC_PROF (SHMFAST_PROF=1 python -m shmfast.build --force); the default build returns zeros. */
int shm_vae_debug_counters(shm_vae* h, long long* out_host, int n);

/* Fused forward + score for n windows: encode -> z = mu + eps*exp(0.5*logvar) -> decode ->
 * score[n] = mean_{t,d} (x - xhat)^2.  Replaces TemporalVAE.forward (temporal_vae.py:72-77) and the
 * scoring loops full_mse_scores_batched (4DOF/Scripts/04_vae_thresholding.py:113-124),
 * 06_test_full_pipeline.py:338-344, recon_mse_per_window (openLAB 10_test_hybrid_pipeline.py:240-251).
 *
 *   src_host : window source (base may point at a series or at materialised windows)
 *   idx      : optional int32[n] source-window list (second pass on flagged windows, 06:358-366)
 *   n_dev    : optional device int32; if non-NULL the effective count is min(*n_dev, n) so the
 *              flagged-subset pass needs no host round trip
 *   eps      : [n,Z] fp32 host-drawn noise (the reference's torch.randn_like, temporal_vae.py:62);
 *              NULL = deterministic z = mu (extra mode, not reference behaviour)
 *   score, mu, logvar, recon, cnn_in : optional outputs ([n], [n,Z], [n,Z], [n,T,D], [n,2,T,D]);
 *              cnn_in is stack([x, (x-xhat)^2], dim=1) of 06_test_full_pipeline.py:364-365 /
 *              05_train_cnn.py:136-138.  The reconstruction stays on chip unless recon/cnn_in is given.
 */
int shm_vae_score(shm_vae* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev,
                  const float* eps, int64_t n, float* score, float* mu, float* logvar, float* recon,
                  float* cnn_in, void* stream);

/* Second pass of the hybrid loop on the flagged windows (06_test_full_pipeline.py:360-365: `recon, _, _ = vae(z_sel)` with fresh noise;
 * training-time twin 05_train_cnn.py:118-141).  In eval mode the encoder is deterministic, so its outputs of the first pass are reused:
 * mu_all / logvar_all [N_all, Z] are shm_vae_score's mu / logvar over ALL windows of `src`; entry j works on window idx[j] (or j if idx
 * is NULL), z = mu_all[idx[j]] + eps[j] * exp(0.5 * logvar_all[idx[j]]), then decode + residual exactly as shm_vae_score does.  Results are
 * bit-identical to shm_vae_score(src, idx, ..., eps) on the same engine.  SHM_ERR_UNSUPPORTED unless the handle runs the tensor-core
 * engine on a stacked (L >= 2) or H = 128 model: call shm_vae_score instead. */
int shm_vae_rescore(shm_vae* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev, const float* mu_all,
                    const float* logvar_all, const float* eps, int64_t n, float* score, float* recon, float* cnn_in, void* stream);

/* TemporalVAE.decode(z, seq_len) (temporal_vae.py:65-70): z [n,Z] -> recon [n,T,D]. */
int shm_vae_decode(shm_vae* h, const float* z, int64_t n, int32_t T, float* recon, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSTM-VAE training step (BASELINE config 5): the inner loop of 4DOF/Scripts/03_train_vae.py:260-271
 *     xhat, mu, logvar = vae(xb); recon = mse_loss(xhat, xb); kl = -0.5*mean(1+logvar-mu^2-exp(logvar));
 *     loss = recon + kl_w*kl; loss.backward(); clip_grad_norm_(params, 2.0); opt.step()   # Adam, wd=1e-5
 * as four entry points, so that either the reference's own loop drives them through an autograd
 * Function (forward / backward) or a fused trainer calls all four with one gradient all-reduce between
 * backward and the optimiser step (data parallel).
 *
 * `params` / `grads` / `exp_avg` / `exp_avg_sq` are flat fp32 device buffers of shm_vae_param_count(cfg)
 * elements in list(model.parameters()) order (temporal_vae.py:29-49): per encoder layer weight_ih, weight_hh,
 * bias_ih, bias_hh; layer_norm.{weight,bias} (if has_ln); fc_mu.{weight,bias}; fc_logvar.{weight,bias};
 * fc_latent_to_hidden.{weight,bias}; per decoder layer the same four; output_layer.{weight,bias}.
 * The caller owns them; the trainer handle owns only the activation workspace for (T, max_batch).
 * Supported shapes: as shm_vae_cfg (H in {32,64,128}, L <= 2, Z <= 16); cfg->engine is ignored (fp32).
 * ------------------------------------------------------------------------------------------- */
typedef struct shm_vae_trainer shm_vae_trainer;
int64_t shm_vae_param_count(const shm_vae_cfg* cfg_host);
int shm_vae_trainer_create(shm_vae_trainer** out, const shm_vae_cfg* cfg_host, int32_t T, int32_t max_batch, int device);
int shm_vae_trainer_destroy(shm_vae_trainer* h);

/* TemporalVAE.forward in train() mode (temporal_vae.py:72-77) for x [B,T,D]; activations are kept in the handle
 * for the backward pass.
 *   eps       : [B,Z] the reparameterisation noise (torch.randn_like, temporal_vae.py:62)
 *   drop_enc / drop_dec : optional uint8 keep-masks [L-1][B,T,H] for the inter-layer dropout of nn.LSTM
 *               (temporal_vae.py:33,47; 1 = keep, kept values are scaled by 1/(1-drop_p)); both NULL = no dropout
 *   xhat [B,T,D], mu [B,Z], logvar [B,Z] : outputs (each optional) */
int shm_vae_train_forward(shm_vae_trainer* h, const float* params, const float* x, int32_t B, const float* eps,
                          const uint8_t* drop_enc, const uint8_t* drop_dec, float drop_p, float* xhat, float* mu,
                          float* logvar, void* stream);

/* Back-propagation through time of the last shm_vae_train_forward: upstream gradients d_xhat [B,T,D] and
 * (optional) d_mu, d_logvar [B,Z] -> grads (flat, overwritten).  One backward per forward. */
int shm_vae_train_backward(shm_vae_trainer* h, const float* params, const float* d_xhat, const float* d_mu,
                           const float* d_logvar, float* grads, void* stream);

/* ELBO of 03_train_vae.py:264-266 and its gradients w.r.t. xhat / mu / logvar (each optional):
 * loss3 = device float[3] {recon + kl_w*kl, recon, kl}; n_x = B*T*D, n_z = B*Z. */
int shm_vae_elbo_grad(const float* x, const float* xhat, const float* mu, const float* logvar, int64_t n_x, int64_t n_z,
                      float kl_w, float* d_xhat, float* d_mu, float* d_logvar, float* loss3, void* stream);

/* torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step (weight_decay added to the
 * gradient, bias-corrected; 03_train_vae.py:222,269-270) over flat buffers.  `grads` is read as grads*grad_scale
 * (1/world_size after a SUM all-reduce).  step >= 1 is the Adam step count AFTER this update.  max_norm <= 0
 * disables clipping.  norm2 = device float[2]: {sum of squares of grads, total norm returned by clip_grad_norm_}. */
int shm_adam_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                       float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                       float* norm2, void* stream);

/* fp32-grade strided contraction used by the training steps (exposed for tests and callers that want it):
 *   C[m*ldc + n] (+)= sum_k A[m*a_ms + k*a_ks] * B[k*b_ks + n*b_ns] (+ bias[n]);  splitk != 0: split along K, atomicAdd into C
 *   (zero C first).  mode SHM_GEMM_SIMT: fp32 FMA pipe.  SHM_GEMM_TC_F16X3 / SHM_GEMM_TC_BF16X3: tcgen05 tensor cores, operands
 *   split on the fly into 16-bit hi/lo halves, 3 MMA passes, fp32 accumulation in TMEM (fp16 halves: 2^-22 relative with an
 *   absolute floor of 2^-25, for O(1) activations; bf16 halves: 2^-16 relative over fp32's whole range, for gradients); shapes
 *   that do not qualify (small, or not float4-loadable along a contiguous dimension) run on the FMA pipe.
 * shm_train_set_tensor_cores: 1 (default) = tensor-core contractions + the two-CTA-cluster recurrence for H = 128; 0 = the fp32 FMA
 * contractions and single-CTA recurrence kernels; 2 = cluster recurrence only; 5 = tensor-core contractions only. */
enum { SHM_GEMM_SIMT = 0, SHM_GEMM_TC_F16X3 = 1, SHM_GEMM_TC_BF16X3 = 2 };
int shm_gemm_f32(const float* A, int64_t a_ms, int64_t a_ks, const float* B, int64_t b_ks, int64_t b_ns, float* C, int64_t ldc,
                 int32_t M, int32_t N, int32_t K, const float* bias, int32_t splitk, int32_t mode, void* stream);
int shm_train_set_tensor_cores(int enable);

/* AdamW twin of shm_adam_clip_step (decoupled weight decay: param *= 1 - lr*weight_decay before the Adam update), the optimiser of
 * openLAB Codes/06_train_cnn.py:395 (AdamW lr 3e-4 wd 1e-4) after clip_grad_norm_(2.0) (:417). */
int shm_adamw_clip_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, int32_t step,
                        float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                        float* norm2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CNN training steps.
 *   SHM_CNN_4DOF    : 4DOF/Scripts/Models/cnn_model.py:16-51 in train() mode (BatchNorm batch statistics + running-stat
 *                     update, Dropout 0.5 after fc1); loop 4DOF/Scripts/05_train_cnn.py:266-281
 *                     (logits = model(xb); loss = CrossEntropyLoss; loss.backward(); Adam lr 1e-4 wd 5e-5).
 *   SHM_CNN_OPENLAB : openLAB Codes/Models/cnn_model.py:16-57 (GroupNorm(8), SiLU, Dropout 0.4); loop Codes/06_train_cnn.py:410-421
 *                     (WeightedFocalLoss gamma 2 (:195-207); clip_grad_norm_(2.0); AdamW lr 3e-4 wd 1e-4).
 * `params` / `grads`: flat fp32 device buffers of shm_cnn_param_count(arch) elements in list(model.parameters()) order
 * (per block conv.weight, conv.bias, norm.weight, norm.bias; then fc1 weight, bias; fc2 weight, bias).
 * forward: x [B,C,T,F] NCHW ([B,2,100,12] / [B,1,200,4]); it must stay alive until the backward call.
 *   bn_running : 4DOF only, optional: [running_mean1(16), running_var1(16), running_mean2(32), running_var2(32)], updated in
 *                place with `bn_momentum` (nn.BatchNorm2d default 0.1; the variance folded in is the unbiased one);
 *   drop_mask  : optional uint8 keep-mask [B,128] of the Dropout after fc1 (1 = keep, kept values scaled by 1/(1-drop_p));
 *   logits     : [B,2].
 * backward: d_logits [B,2] -> grads (overwritten).  One backward per forward.
 * shm_cnn_loss_grad: loss (device float[1]) and d loss / d logits for int64 targets [B]; alpha == NULL and gamma == 0 =
 *   nn.CrossEntropyLoss (mean); otherwise mean(alpha[t] * (1 - pt)^gamma * ce), pt = exp(-ce) (alpha: device float[2] or NULL).
 * ------------------------------------------------------------------------------------------- */
enum { SHM_CNN_4DOF = 0, SHM_CNN_OPENLAB = 1 };
typedef struct shm_cnn_trainer shm_cnn_trainer;
int64_t shm_cnn_param_count(int arch);
int shm_cnn_trainer_create(shm_cnn_trainer** out, int arch, int32_t max_batch, int device);
int shm_cnn_trainer_destroy(shm_cnn_trainer* h);
int shm_cnn_train_forward(shm_cnn_trainer* h, const float* params, const float* x, int32_t B, float* bn_running, float bn_momentum,
                          const uint8_t* drop_mask, float drop_p, float* logits, void* stream);
int shm_cnn_train_backward(shm_cnn_trainer* h, const float* params, const float* d_logits, float* grads, void* stream);
int shm_cnn_loss_grad(const float* logits, const int64_t* targets, int64_t B, const float* alpha, float gamma, float* d_logits,
                      float* loss, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Threshold + routing: mask = score > thr (strict, fp32), idx = np.where(mask)[0] ascending
 * (06_test_full_pipeline.py:350-351, 10_test_hybrid_pipeline.py:367).
 *   flag  : optional uint8[N];  idx : int32[N] (first *count entries valid);  count : device int32
 *   workspace : device scratch of shm_compact_workspace_bytes(N) bytes
 * ------------------------------------------------------------------------------------------- */
int64_t shm_compact_workspace_bytes(int64_t N);
int shm_compact(const float* score, float thr, int64_t N, uint8_t* flag, int32_t* idx, int32_t* count,
                void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 4DOF CNN (sensor-fault vs structural-fault): 4DOF/Scripts/Models/cnn_model.py:8-57, eval mode
 * (BatchNorm running stats, Dropout identity).  x [n,2,100,12] -> logits [n,2];
 * optional label = argmax+1 (int64, tie -> 1) and p_struct = softmax[:,1]
 * (06_test_full_pipeline.py:366-372).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* conv_w[2];     /* conv{1,2}.0.weight [16,2,3,3], [32,16,3,3] */
    const float* conv_b[2];     /* conv{1,2}.0.bias */
    const float* bn_w[2];       /* conv{1,2}.1.weight */
    const float* bn_b[2];       /* conv{1,2}.1.bias */
    const float* bn_mean[2];    /* conv{1,2}.1.running_mean */
    const float* bn_var[2];     /* conv{1,2}.1.running_var */
    const float* fc1_w;         /* fc1.0.weight [128,2400] */
    const float* fc1_b;
    const float* fc2_w;         /* fc2.weight [2,128] */
    const float* fc2_b;
    float bn_eps;               /* 1e-5 */
} shm_cnn4dof_weights;

typedef struct shm_cnn4dof shm_cnn4dof;
int shm_cnn4dof_create(shm_cnn4dof** out, const shm_cnn4dof_weights* w_host, int device);
int shm_cnn4dof_update_weights(shm_cnn4dof* h, const shm_cnn4dof_weights* w_host, void* stream);
int shm_cnn4dof_destroy(shm_cnn4dof* h);
int shm_cnn4dof_forward(shm_cnn4dof* h, const float* x, const int32_t* n_dev, int64_t n, float* logits,
                        int64_t* label, float* p_struct, void* stream);

/* ---------------------------------------------------------------------------------------------
 * openLAB CNN: 20250506_openLAB_tests/Codes/Models/cnn_model.py:8-57 (Conv+GroupNorm(8)+SiLU x4,
 * MaxPool(2,1) x3, GAP, 256->128 SiLU ->2), eval mode.  The input is described by a window source
 * over X_raw (T=200, D=4) so the flagged-window gather + standardise (clip 10, NaN->0) of
 * stage2_predict_cnn (10_test_hybrid_pipeline.py:272-278) is fused into the first convolution.
 * prob = softmax[:,1] widened to fp64 for the `>= thr` decision (:294-301).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const float* conv_w[4];     /* features.{0,2,4,6}.0.weight [32,1,7,3],[64,32,5,3],[128,64,5,3],[256,128,3,3] */
    const float* conv_b[4];     /* features.{0,2,4,6}.0.bias */
    const float* gn_w[4];       /* features.{0,2,4,6}.1.weight */
    const float* gn_b[4];       /* features.{0,2,4,6}.1.bias */
    const float* fc1_w;         /* classifier.1.weight [128,256] */
    const float* fc1_b;
    const float* fc2_w;         /* classifier.4.weight [2,128] */
    const float* fc2_b;
    float gn_eps;               /* 1e-5 */
} shm_cnnol_weights;

typedef struct shm_cnnol shm_cnnol;
int shm_cnnol_create(shm_cnnol** out, const shm_cnnol_weights* w_host, int device);
int shm_cnnol_update_weights(shm_cnnol* h, const shm_cnnol_weights* w_host, void* stream);
int shm_cnnol_destroy(shm_cnnol* h);
int shm_cnnol_forward(shm_cnnol* h, const shm_window_src* src_host, const int32_t* idx, const int32_t* n_dev,
                      int64_t n, float* logits, double* prob, void* stream);
/* Engine of the convolution blocks: SHM_ENGINE_TC_BF16X3 (default; blocks 2-4 as implicit GEMMs on the tensor cores,
 * 3-pass fp16 hi/lo split, fp32 accumulate; the first call allocates a workspace of ~0.21 MB per window for up to 9472
 * windows per internal chunk) or SHM_ENGINE_FP32 (everything on the CUDA cores, one CTA per window, no workspace). */
int shm_cnnol_set_engine(shm_cnnol* h, int engine);
int shm_cnnol_engine(const shm_cnnol* h);

/* ---------------------------------------------------------------------------------------------
 * Fused hybrid loops: the reference's script-level hot loops as ONE call on one stream.  No host synchronisation:
 * the flagged count stays on the device; per-flagged scratch is bounded (the 4DOF second pass runs in chunks of
 * 75,776 flagged windows), so `max_flagged` may be as large as n.
 *
 * shm_hybrid4dof_score == eval_group (4DOF/Scripts/06_test_full_pipeline.py:327-383): score pass (eps1[n,Z]) ->
 *   `mse > thr` + np.where (:350-351) -> SECOND VAE pass on the flagged windows with fresh noise (eps2[j] belongs to
 *   the j-th flagged window, :360-363) -> stack([z, (z-zhat)^2]) (:364-365) -> CNN (:366) -> cls = argmax,
 *   y_pred[sel] = cls+1, hyb_score_full[sel] = softmax[:,1] (:367-372).
 *     score[n], flag[n] (opt), idx[n] (first status[0] entries valid, ascending);
 *     status: device int32[2] = {flagged count, 1 if count > max_flagged (windows past max_flagged were NOT attributed)};
 *     logits[max_flagged,2], label[max_flagged], p_struct[max_flagged]: optional per-flagged outputs;
 *     y_pred[n] int64 / p_full[n] fp32: optional dense outputs, 0 for windows that were not flagged (:336,356).
 * shm_hybridol_score == 10_test_hybrid_pipeline.py:351-367 (gate on src_gate) + stage2_predict_cnn (:265-302, CNN on
 *   src_raw[flagged], prob_st = softmax[:,1] as fp64, pred_bin = prob_st >= cnn_thr) + the label scatter (:389-401):
 *     prob[max_flagged], pred_bin[max_flagged]: optional per-flagged outputs;
 *     y_pred[n]: 0 = not flagged, 1 = sensor fault (pred_bin 0), 2 = structural (pred_bin 1); prob_full[n] fp64.
 * shm_scatter_flagged_*: only the scatter step, for callers that ran the stages themselves (`count` may be NULL:
 *   all `cap` entries are valid).  Dense outputs are zero-filled first.
 * ------------------------------------------------------------------------------------------- */
int64_t shm_hybrid4dof_workspace_bytes(const shm_vae* vae, int64_t n, int64_t max_flagged);
int shm_hybrid4dof_score(shm_vae* vae, shm_cnn4dof* cnn, const shm_window_src* src_host, int64_t n, const float* eps1,
                         const float* eps2, float thr, int64_t max_flagged, float* score, uint8_t* flag, int32_t* idx,
                         int32_t* status, float* logits, int64_t* label, float* p_struct, int64_t* y_pred, float* p_full,
                         void* workspace, int64_t workspace_bytes, void* stream);
int64_t shm_hybridol_workspace_bytes(int64_t n, int64_t max_flagged);
int shm_hybridol_score(shm_vae* vae, shm_cnnol* cnn, const shm_window_src* src_gate_host, const shm_window_src* src_raw_host,
                       int64_t n, const float* eps, float vae_thr, double cnn_thr, int64_t max_flagged, float* score,
                       uint8_t* flag, int32_t* idx, int32_t* status, float* logits, double* prob, int64_t* pred_bin,
                       int64_t* y_pred, double* prob_full, void* workspace, int64_t workspace_bytes, void* stream);
int shm_scatter_flagged_4dof(const int32_t* idx, const int32_t* count, int64_t cap, const int64_t* label, const float* p_struct,
                             int64_t n, int64_t* y_pred, float* p_full, void* stream);
int shm_scatter_flagged_openlab(const int32_t* idx, const int32_t* count, int64_t cap, const double* prob, double cnn_thr,
                                int64_t n, int64_t* pred_bin, int64_t* y_pred, double* prob_full, void* stream);

/* ---------------------------------------------------------------------------------------------
 * openLAB extraction front-end: what feeds the hybrid path.  One run's parsed catman
 * columns raw [R,4] float32 = DMS_1, LWA_2, LWA_3, LWA_4 (as _to_float yields them,
 * 20250506_openLAB_tests/Codes/01_extract_windows_and_labels.py:58-59) -> obstruction sentinel -> NaN (:117-119),
 * provider AND-rule outlier masks (:65-83), clean_openlab_and_rule (feature_utils.py:49-99: AND-rule removal, linear
 * interpolation, moving average; fp64 -> fp32), rows with a finite DMS kept (:151-156), and per window of
 * (T, stride) (:159-214): mask ratios, structural envelope u_min/u_max over the selected clean channels, DMS range,
 * load-aware flatline flag and the label (0 Normal, 1 Sensor Fault, 2 Structural Fault; SF > ST > Normal).
 * a_clean / a_raw [R,4]: the first *rows_kept rows are valid; window w is rows [w*stride, w*stride+T) of them -- describe
 * them to shm_vae_score / shm_cnnol_forward with a shm_window_src (win_stride = stride*4, row_stride = 4).
 * Per-window outputs hold (R-T)/stride+1 entries, the first *n_windows valid.  Bit-identical to the reference.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t T, stride;                    /* config.py:27-28  SEQ_LEN 200, STRIDE 20 */
    int32_t ma_window;                    /* config.py:45     MOVING_AVG_WINDOW 5 (odd, or <= 1 for none) */
    int32_t struct_channel_mask;          /* bit k = LWA_{2+k} defines the structural envelope; 01:50 uses LWA_3 only = 2 */
    double obstruction_sentinel;          /* config.py:42     -1e5 */
    double raw_diff_th, raw_abs_th;       /* config.py:50-51  1.0, 65.0 (>=, float32 arithmetic) */
    double clean_max_jump, clean_max_abs; /* config.py:43-44  1.0, 65.0 (>, float64 arithmetic) */
    float raw_invalid_ratio_fault;        /* config.py:52     0.05 */
    float flat_var_eps;                   /* config.py:55     1e-6 */
    float force_range_for_flatline;       /* config.py:56     5.0 */
    float allow_max;                      /* config.py:37     20.0 */
} shm_openlab_extract_cfg;

int64_t shm_openlab_extract_workspace_bytes(int64_t R);
int shm_openlab_extract(const float* raw, int64_t R, const shm_openlab_extract_cfg* cfg_host, float* a_clean, float* a_raw,
                        int32_t* rows_kept, int32_t* n_windows, int32_t* label, float* u_min, float* u_max, float* dms_range,
                        float* raw_invalid_ratio, float* raw_outlier_ratio, float* removed_ratio, int32_t* flatline_loadaware,
                        int32_t* all_nan_struct, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * 1_DOF post-processing: overlap-average reconstructed windows back to a series, de-standardise,
 * RMSE per `segment_len` rows over all channels -- stitch_windows / destandardize / segment_rmse
 * (1_DOF/Scripts/datasets.py:21-22,38-71; call site 04_test_seen_variants.py:296-311).  fp64 like
 * the reference.  recon [N,T,F] fp32; y_true [full_len,F] fp32 (physical units); mean/std fp64[F];
 * series_out optional [full_len,F] fp64; rmse_out [ceil(full_len/segment_len)] fp64.
 * ------------------------------------------------------------------------------------------- */
int shm_stitch_segment_rmse(const float* recon, int64_t N, int32_t T, int32_t F, int32_t stride, int64_t full_len,
                            const double* mean, const double* std, const float* y_true, int32_t segment_len,
                            double* series_out, double* rmse_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Exact percentile on device with NumPy's default linear interpolation:
 * np.percentile(scores, q) of 04_vae_thresholding.py:283 / openLAB 05_validate_vae.py:253.
 * result: device double.  workspace: shm_percentile_workspace_bytes(N).
 * ------------------------------------------------------------------------------------------- */
int64_t shm_percentile_workspace_bytes(int64_t N);
int shm_percentile(const float* scores, int64_t N, double q, double* result, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHMFAST_H */
