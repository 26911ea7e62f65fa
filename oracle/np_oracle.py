"""NumPy restatement of the reference hot path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Every function cites the reference file:line whose behaviour it restates
(paths relative to /root/reference).  Weights are passed as a dict keyed by the
reference's own state_dict names, values np.ndarray.
`dtype` selects the arithmetic type: np.float32 mimics the reference's own
arithmetic, np.float64 is the "truth" used for tolerance studies.

Pinned against the reference modules by tests/golden/make_golden.py +
tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# a1-a4  windowing + normalisation
# --------------------------------------------------------------------------------------

def slice_frac(X: np.ndarray, frac_range) -> np.ndarray:
    """4DOF/Scripts/06_test_full_pipeline.py:98-103 (copies in 03/04/05)."""
    n = X.shape[0]
    s = int(n * float(frac_range[0]))
    e = int(n * float(frac_range[1]))
    e = max(e, s)
    return X[s:e]


def n_windows(rows: int, T: int, stride: int) -> int:
    """Window count of every make_windows/windowize_2d in the reference: range(0, rows-T+1, stride)."""
    if rows < T:
        return 0
    return (rows - T) // stride + 1


def make_windows(X: np.ndarray, T: int, stride: int) -> np.ndarray:
    """4DOF/Scripts/06_test_full_pipeline.py:106-110; openLAB feature_utils.py:130-152 (windowize_2d);
    1_DOF/Scripts/datasets.py:25-35.  Short input -> empty [0,T,D] (4DOF/openLAB behaviour;
    1_DOF raises instead, see make_windows_1dof)."""
    X = np.asarray(X)
    N = n_windows(X.shape[0], T, stride)
    if N == 0:
        return np.zeros((0, T, X.shape[1]), dtype=np.float32)
    idx = np.arange(N)[:, None] * stride + np.arange(T)[None, :]
    return X[idx].astype(np.float32)


def make_windows_1dof(X: np.ndarray, T: int, stride: int = 1) -> np.ndarray:
    """1_DOF/Scripts/datasets.py:25-35 -- raises on short input and keeps the input dtype (fp64
    after standardize); the cast to fp32 happens at 04_test_seen_variants.py:292."""
    if X.shape[0] < T:
        raise ValueError(f"Time series too short: T={X.shape[0]} < seq_len={T}")
    N = n_windows(X.shape[0], T, stride)
    idx = np.arange(N)[:, None] * stride + np.arange(T)[None, :]
    return X[idx]


def guard_std_4dof(std: np.ndarray) -> np.ndarray:
    """4DOF/Scripts/06_test_full_pipeline.py:119-120: std[std == 0] = 1e-6 (fp32)."""
    std = np.array(std, dtype=np.float32, copy=True)
    std[std == 0] = 1e-6
    return std


def normalize_windows_4dof(W: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """4DOF/Scripts/06_test_full_pipeline.py:124-126: (W-mean)/std then nan/+-inf -> 0, fp32."""
    W = np.asarray(W, dtype=np.float32)
    with np.errstate(all="ignore"):
        Z = (W - mean[None, None, :].astype(np.float32)) / std[None, None, :].astype(np.float32)
    return np.nan_to_num(Z, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)


def standardize_openlab(X: np.ndarray, mu: np.ndarray, sd: np.ndarray, clip: float = 10.0) -> np.ndarray:
    """openLAB 10_test_hybrid_pipeline.py:233-237: (X-mu)/sd, clip(+-clip) [NaN survives np.clip],
    then nan/+-inf -> 0, fp32."""
    X = np.asarray(X, dtype=np.float32)
    with np.errstate(all="ignore"):
        Xn = (X - mu[None, None, :].astype(np.float32)) / sd[None, None, :].astype(np.float32)
        Xn = np.clip(Xn, -float(clip), float(clip))
    Xn = np.nan_to_num(Xn, nan=0.0, posinf=0.0, neginf=0.0)
    return Xn.astype(np.float32)


def standardize_series_1dof(x: np.ndarray, mean: np.ndarray, std: np.ndarray) -> np.ndarray:
    """1_DOF/Scripts/datasets.py:17-18 (series standardised BEFORE windowing; numpy promotes to the
    stats' dtype, fp64 in the reference)."""
    return (x - mean) / std


def compute_standardizer_1dof(x: np.ndarray):
    """1_DOF/Scripts/datasets.py:6-14."""
    mean = x.mean(axis=0)
    std = x.std(axis=0)
    std = np.where(std == 0.0, 1e-6, std)
    return mean, std


# --------------------------------------------------------------------------------------
# a5-a8  LSTM-VAE
# --------------------------------------------------------------------------------------

def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def vae_config(sd: dict) -> dict:
    """Shapes implied by a reference VAE state_dict."""
    w = sd["encoder_lstm.weight_ih_l0"]
    H = w.shape[0] // 4
    D = w.shape[1]
    L = 0
    while f"encoder_lstm.weight_ih_l{L}" in sd:
        L += 1
    Z = sd["fc_mu.weight"].shape[0]
    return dict(D=D, H=H, Z=Z, L=L, has_ln="layer_norm.weight" in sd)


def lstm_forward(x: np.ndarray, sd: dict, prefix: str, L: int, dtype=np.float32):
    """torch.nn.LSTM(batch_first=True) in eval mode as called at
    4DOF/Scripts/Models/temporal_vae.py:29-35,53 and :42-48,68: zero initial (h,c); gate row order
    i,f,g,o; pre-activation = x W_ih^T + b_ih + h W_hh^T + b_hh; c = f*c + i*g; h = o*tanh(c).
    Returns (top-layer outputs [B,T,H], h_n [L,B,H])."""
    x = np.asarray(x, dtype=dtype)
    B, T, _ = x.shape
    inp = x
    h_n = []
    for l in range(L):
        Wih = sd[f"{prefix}.weight_ih_l{l}"].astype(dtype)
        Whh = sd[f"{prefix}.weight_hh_l{l}"].astype(dtype)
        b = (sd[f"{prefix}.bias_ih_l{l}"].astype(dtype) + sd[f"{prefix}.bias_hh_l{l}"].astype(dtype))
        H = Whh.shape[1]
        h = np.zeros((B, H), dtype=dtype)
        c = np.zeros((B, H), dtype=dtype)
        out = np.empty((B, T, H), dtype=dtype)
        pre_all = inp.reshape(B * T, -1) @ Wih.T
        pre_all = pre_all.reshape(B, T, 4 * H) + b
        for t in range(T):
            g = pre_all[:, t, :] + h @ Whh.T
            i = _sigmoid(g[:, 0 * H:1 * H])
            f = _sigmoid(g[:, 1 * H:2 * H])
            gg = np.tanh(g[:, 2 * H:3 * H])
            o = _sigmoid(g[:, 3 * H:4 * H])
            c = f * c + i * gg
            h = o * np.tanh(c)
            out[:, t, :] = h
        inp = out
        h_n.append(h)
    return inp, np.stack(h_n, axis=0)


def vae_encode(sd: dict, x: np.ndarray, dtype=np.float32):
    """TemporalVAE.encode: 4DOF/Scripts/Models/temporal_vae.py:51-58 (LayerNorm on h_n[-1], biased
    variance, eps 1e-5); 1_DOF/Scripts/Models/temporal_vae.py:41-46 (no LayerNorm);
    openLAB Codes/Models/temporal_vae_model.py:35-42."""
    cfg = vae_config(sd)
    _, h_n = lstm_forward(x, sd, "encoder_lstm", cfg["L"], dtype)
    h = h_n[-1]
    if cfg["has_ln"]:
        m = h.mean(axis=1, keepdims=True)
        v = ((h - m) ** 2).mean(axis=1, keepdims=True)
        h = (h - m) / np.sqrt(v + dtype(1e-5))
        h = h * sd["layer_norm.weight"].astype(dtype) + sd["layer_norm.bias"].astype(dtype)
    mu = h @ sd["fc_mu.weight"].astype(dtype).T + sd["fc_mu.bias"].astype(dtype)
    lv = h @ sd["fc_logvar.weight"].astype(dtype).T + sd["fc_logvar.bias"].astype(dtype)
    return mu, lv


def vae_reparam(mu: np.ndarray, logvar: np.ndarray, eps: np.ndarray | None):
    """TemporalVAE.reparameterize: temporal_vae.py:60-63 with the randn_like draw supplied by the
    caller (eps).  eps=None is the extra deterministic z=mu mode (NOT reference behaviour,
    SURVEY.md section 0 row 3)."""
    if eps is None:
        return mu
    return mu + eps.astype(mu.dtype) * np.exp(mu.dtype.type(0.5) * logvar)


def vae_decode(sd: dict, z: np.ndarray, T: int, dtype=np.float32):
    """TemporalVAE.decode: temporal_vae.py:65-70: h0 = tanh(fc(z)) repeated over T as decoder-LSTM
    input; Linear H->D on every step."""
    cfg = vae_config(sd)
    z = z.astype(dtype)
    u = np.tanh(z @ sd["fc_latent_to_hidden.weight"].astype(dtype).T + sd["fc_latent_to_hidden.bias"].astype(dtype))
    inp = np.repeat(u[:, None, :], T, axis=1)
    dec, _ = lstm_forward(inp, sd, "decoder_lstm", cfg["L"], dtype)
    return dec @ sd["output_layer.weight"].astype(dtype).T + sd["output_layer.bias"].astype(dtype)


def vae_forward(sd: dict, x: np.ndarray, eps: np.ndarray | None, dtype=np.float32):
    """TemporalVAE.forward: temporal_vae.py:72-77 -> (recon, mu, logvar)."""
    x = np.asarray(x, dtype=dtype)
    mu, lv = vae_encode(sd, x, dtype)
    z = vae_reparam(mu, lv, eps)
    recon = vae_decode(sd, z, x.shape[1], dtype)
    return recon, mu, lv


# --------------------------------------------------------------------------------------
# a9-a10  score, threshold, routing
# --------------------------------------------------------------------------------------

def mse_score(x: np.ndarray, recon: np.ndarray) -> np.ndarray:
    """4DOF/Scripts/04_vae_thresholding.py:122, 06_test_full_pipeline.py:343;
    openLAB 10_test_hybrid_pipeline.py:249: per-window mean over (T,D) of squared error."""
    d = (x - recon)
    return (d * d).mean(axis=(1, 2)).astype(np.float32)


def vae_scores_batched(sd: dict, X: np.ndarray, eps: np.ndarray | None, batch: int, dtype=np.float32):
    """full_mse_scores_batched (04_vae_thresholding.py:113-124) / recon_mse_per_window
    (10_test_hybrid_pipeline.py:240-251) with eps supplied per window."""
    out = np.zeros((X.shape[0],), dtype=np.float32)
    for i in range(0, X.shape[0], batch):
        xb = X[i:i + batch].astype(dtype)
        e = None if eps is None else eps[i:i + batch]
        recon, _, _ = vae_forward(sd, xb, e, dtype)
        out[i:i + batch] = mse_score(xb, recon)
    return out


def flag_compact(score: np.ndarray, thr: float):
    """06_test_full_pipeline.py:350-351; 10_test_hybrid_pipeline.py:367: strict `>` evaluated in fp32
    (NumPy keeps the fp32 array dtype against a Python float), np.where ascending."""
    mask = np.asarray(score, dtype=np.float32) > np.float32(thr)
    return mask, np.where(mask)[0].astype(np.int64)


def percentile_linear(scores: np.ndarray, q: float) -> float:
    """np.percentile(s, q) with the default linear interpolation as used at
    4DOF/Scripts/04_vae_thresholding.py:283 and openLAB 05_validate_vae.py:253."""
    return float(np.percentile(np.asarray(scores), q))


# --------------------------------------------------------------------------------------
# a11-a13  4DOF CNN
# --------------------------------------------------------------------------------------

def cnn4dof_inputs(z: np.ndarray, recon: np.ndarray) -> np.ndarray:
    """06_test_full_pipeline.py:364-365 / 05_train_cnn.py:136-138: stack([z, (z-zhat)^2], dim=1)."""
    resid = (z - recon) ** 2
    return np.stack([z, resid], axis=1).astype(np.float32)


def _conv2d(x, w, b, pad):
    """Conv2d stride 1, zero padding (pt, pf).  x [B,C,Hh,Ww], w [O,C,kh,kw]."""
    B, C, Hh, Ww = x.shape
    O, _, kh, kw = w.shape
    pt, pf = pad
    xp = np.zeros((B, C, Hh + 2 * pt, Ww + 2 * pf), dtype=x.dtype)
    xp[:, :, pt:pt + Hh, pf:pf + Ww] = x
    Ho = Hh + 2 * pt - kh + 1
    Wo = Ww + 2 * pf - kw + 1
    cols = np.empty((B, C, kh, kw, Ho, Wo), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            cols[:, :, i, j] = xp[:, :, i:i + Ho, j:j + Wo]
    y = np.einsum("bckl hw,ockl->bohw".replace(" ", ""), cols, w, optimize=True)
    return y + b[None, :, None, None]


def _maxpool(x, kh, kw):
    """MaxPool2d(kernel=(kh,kw)) with stride=kernel and floor output size."""
    B, C, Hh, Ww = x.shape
    Ho, Wo = Hh // kh, Ww // kw
    x = x[:, :, :Ho * kh, :Wo * kw].reshape(B, C, Ho, kh, Wo, kw)
    return x.max(axis=(3, 5))


def cnn4dof_forward(sd: dict, xin: np.ndarray, dtype=np.float32) -> np.ndarray:
    """CNN.forward in eval mode: 4DOF/Scripts/Models/cnn_model.py:16-34,45-51.
    conv3x3(pad1)+BatchNorm(running stats, eps 1e-5)+ReLU+MaxPool2 twice; flatten (C,H,W);
    Linear 2400->128 + ReLU (+Dropout=identity); Linear 128->2."""
    x = np.asarray(xin, dtype=dtype)
    for blk in ("conv1", "conv2"):
        x = _conv2d(x, sd[f"{blk}.0.weight"].astype(dtype), sd[f"{blk}.0.bias"].astype(dtype), (1, 1))
        rm = sd[f"{blk}.1.running_mean"].astype(dtype)[None, :, None, None]
        rv = sd[f"{blk}.1.running_var"].astype(dtype)[None, :, None, None]
        g = sd[f"{blk}.1.weight"].astype(dtype)[None, :, None, None]
        be = sd[f"{blk}.1.bias"].astype(dtype)[None, :, None, None]
        x = (x - rm) / np.sqrt(rv + dtype(1e-5)) * g + be
        x = np.maximum(x, 0)
        x = _maxpool(x, 2, 2)
    x = x.reshape(x.shape[0], -1)
    x = np.maximum(x @ sd["fc1.0.weight"].astype(dtype).T + sd["fc1.0.bias"].astype(dtype), 0)
    return x @ sd["fc2.weight"].astype(dtype).T + sd["fc2.bias"].astype(dtype)


def softmax(logits: np.ndarray) -> np.ndarray:
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    return e / e.sum(axis=1, keepdims=True)


def cnn4dof_labels(logits: np.ndarray):
    """06_test_full_pipeline.py:367-372: label = argmax+1 (tie -> class 0 -> label 1); p_struct = softmax[:,1]."""
    cls = np.argmax(logits, axis=1).astype(np.int64)
    return cls + 1, softmax(logits.astype(np.float32))[:, 1].astype(np.float32)


# --------------------------------------------------------------------------------------
# a14  openLAB CNN
# --------------------------------------------------------------------------------------

def _silu(x):
    return x / (1.0 + np.exp(-x))


def _groupnorm(x, w, b, groups, eps):
    B, C, Hh, Ww = x.shape
    xg = x.reshape(B, groups, -1)
    m = xg.mean(axis=2, keepdims=True)
    v = ((xg - m) ** 2).mean(axis=2, keepdims=True)
    xg = (xg - m) / np.sqrt(v + x.dtype.type(eps))
    x = xg.reshape(B, C, Hh, Ww)
    return x * w[None, :, None, None] + b[None, :, None, None]


OPENLAB_CNN_BLOCKS = ((0, 7, 3, 3, 1, True), (2, 5, 3, 2, 1, True), (4, 5, 3, 2, 1, True), (6, 3, 3, 1, 1, False))


def cnnol_forward(sd: dict, x: np.ndarray, dtype=np.float32) -> np.ndarray:
    """openLAB CNN.forward (eval): 20250506_openLAB_tests/Codes/Models/cnn_model.py:16-43,54-57.
    4 x [Conv(kt x 3, same pad) + GroupNorm(8, eps 1e-5) + SiLU], MaxPool(2,1) after blocks 1-3,
    global average pool, Linear 256->128 + SiLU (+Dropout=identity), Linear 128->2.  x [B,1,200,4]."""
    x = np.asarray(x, dtype=dtype)
    for (idx, kt, kf, pt, pf, pool) in OPENLAB_CNN_BLOCKS:
        x = _conv2d(x, sd[f"features.{idx}.0.weight"].astype(dtype), sd[f"features.{idx}.0.bias"].astype(dtype), (pt, pf))
        x = _groupnorm(x, sd[f"features.{idx}.1.weight"].astype(dtype), sd[f"features.{idx}.1.bias"].astype(dtype), 8, 1e-5)
        x = _silu(x)
        if pool:
            x = _maxpool(x, 2, 1)
    x = x.mean(axis=(2, 3))
    x = _silu(x @ sd["classifier.1.weight"].astype(dtype).T + sd["classifier.1.bias"].astype(dtype))
    return x @ sd["classifier.4.weight"].astype(dtype).T + sd["classifier.4.bias"].astype(dtype)


def cnnol_decision(logits: np.ndarray, thr: float):
    """10_test_hybrid_pipeline.py:294-301: p = softmax(logits)[:,1] (fp32) -> float64; pred = p >= thr."""
    p = softmax(logits.astype(np.float32))[:, 1].astype(np.float64)
    return (p >= float(thr)).astype(np.int64), p


# --------------------------------------------------------------------------------------
# a15  1_DOF stitch + segment RMSE
# --------------------------------------------------------------------------------------

def stitch_windows(windows: np.ndarray, full_len: int, stride: int = 1) -> np.ndarray:
    """1_DOF/Scripts/datasets.py:38-54: overlap-average windows back into a series (fp64)."""
    N, T, F = windows.shape
    out = np.zeros((full_len, F), dtype=float)
    cnt = np.zeros((full_len, 1), dtype=float)
    for n in range(N):
        s = n * stride
        out[s:s + T] += windows[n]
        cnt[s:s + T] += 1.0
    cnt[cnt == 0.0] = 1.0
    return out / cnt


def destandardize(xn, mean, std):
    """1_DOF/Scripts/datasets.py:21-22."""
    return xn * std + mean


def segment_rmse(y_true: np.ndarray, y_pred: np.ndarray, segment_len: int) -> np.ndarray:
    """1_DOF/Scripts/datasets.py:57-71: RMSE over all channels per segment of segment_len rows."""
    T = y_true.shape[0]
    S = int(np.ceil(T / segment_len))
    out = np.empty((S,), dtype=float)
    for s in range(S):
        i0, i1 = s * segment_len, min((s + 1) * segment_len, T)
        e = y_pred[i0:i1] - y_true[i0:i1]
        out[s] = float(np.sqrt(np.mean(e ** 2)))
    return out


# --------------------------------------------------------------------------------------
# whole-pipeline restatements (the script-level hot loops)
# --------------------------------------------------------------------------------------

def hybrid_4dof(vae_sd, cnn_sd, Z, eps1, eps2, thr, batch=512, dtype=np.float32):
    """eval_group of 4DOF/Scripts/06_test_full_pipeline.py:327-383 on already-normalised windows Z:
    score pass (eps1 per window) -> strict threshold -> SECOND VAE pass on the flagged subset with
    fresh eps (eps2, indexed by position in the flagged list, as the reference draws them batch by
    batch over idx_anom) -> residual stack -> CNN -> label = argmax+1, p_struct."""
    N = Z.shape[0]
    score = vae_scores_batched(vae_sd, Z, eps1, batch, dtype)
    mask, idx = flag_compact(score, thr)
    y_pred = np.zeros((N,), dtype=np.int64)
    p_struct = np.zeros((N,), dtype=np.float32)
    logits_all = np.zeros((idx.size, 2), dtype=np.float32)
    for j in range(0, idx.size, batch):
        sel = idx[j:j + batch]
        zb = Z[sel].astype(dtype)
        recon, _, _ = vae_forward(vae_sd, zb, None if eps2 is None else eps2[j:j + sel.size], dtype)
        xin = cnn4dof_inputs(zb, recon)
        logits = cnn4dof_forward(cnn_sd, xin, dtype)
        logits_all[j:j + batch] = logits
        lab, ps = cnn4dof_labels(logits)
        y_pred[sel] = lab
        p_struct[sel] = ps
    return dict(score=score, mask=mask, idx=idx, logits=logits_all, y_pred=y_pred, p_struct=p_struct)


def hybrid_openlab(vae_sd, cnn_sd, X_clean, X_raw, channels_idx, vae_mu, vae_sd_, cnn_mu, cnn_sd_, eps,
                   vae_thr, cnn_thr, batch=256, clip=10.0, dtype=np.float32):
    """openLAB 10_test_hybrid_pipeline.py:351-367 + stage2_predict_cnn :265-302: gate on
    X_clean[:,:,channels_idx] standardised with the VAE stats; CNN on X_raw[mask] standardised with
    the CNN's own stats; decision softmax[:,1] >= thr in float64."""
    Xg = standardize_openlab(X_clean[:, :, channels_idx], vae_mu, vae_sd_, clip)
    score = vae_scores_batched(vae_sd, Xg, eps, batch, dtype)
    mask, idx = flag_compact(score, vae_thr)
    Xa = standardize_openlab(X_raw[mask].astype(np.float32), cnn_mu, cnn_sd_, clip)[:, None, :, :]
    logits = np.zeros((idx.size, 2), dtype=np.float32)
    for j in range(0, idx.size, batch):
        logits[j:j + batch] = cnnol_forward(cnn_sd, Xa[j:j + batch], dtype)
    pred, p = cnnol_decision(logits, cnn_thr)
    return dict(score=score, mask=mask, idx=idx, logits=logits, pred=pred, prob=p)


def adam_clip_step(p, g, m, v, step, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8, wd=1e-5, max_norm=2.0, grad_scale=1.0):
    """torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam.step (weight_decay as L2-in-grad) over flat fp32
    arrays, 4DOF/Scripts/03_train_vae.py:222,269-270.  Returns (p, m, v, total_norm); `g` is read as g*grad_scale."""
    f = np.float32
    g = (g.astype(f) * f(grad_scale)).astype(f)
    total = f(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
    if max_norm > 0:
        coef = min(1.0, float(max_norm) / (float(total) + 1e-6))
        g = (g * f(coef)).astype(f)
    g = (g + f(wd) * p).astype(f)
    m = (f(b1) * m + f(1 - b1) * g).astype(f)
    v = (f(b2) * v + f(1 - b2) * g * g).astype(f)
    bc1 = 1.0 - b1 ** step
    bc2s = np.sqrt(1.0 - b2 ** step)
    denom = (np.sqrt(v) / f(bc2s) + f(eps)).astype(f)
    p = (p - f(lr / bc1) * (m / denom)).astype(f)
    return p, m, v, float(total)


# --------------------------------------------------------------------------------------
# f2  openLAB extraction front-end: cleaned/raw series -> windows -> rule labels
#     (20250506_openLAB_tests/Codes/01_extract_windows_and_labels.py:86-236, feature_utils.py:49-99,130-177)
# --------------------------------------------------------------------------------------
OPENLAB_EXTRACT_CFG = dict(T=200, stride=20, sentinel=-1e5, raw_diff_th=1.0, raw_abs_th=65.0, clean_max_jump=1.0,
                           clean_max_abs=65.0, ma_window=5, raw_invalid_ratio_fault=0.05, flat_var_eps=1e-6,
                           force_range_for_flatline=5.0, allow_max=20.0, struct_channels=(2,))   # config.py:27-56; 01:50 (LWA_3)
LABEL_NORMAL, LABEL_SENSOR_FAULT, LABEL_STRUCT_FAULT = 0, 1, 2


def provider_outlier_mask(u: np.ndarray, diff_th: float, abs_th: float) -> np.ndarray:
    """provider_raw_outlier_mask_AND, 01_extract_windows_and_labels.py:65-83 (float32 arithmetic)."""
    u = np.asarray(u, dtype=np.float32)
    m = ~np.isfinite(u)
    if u.size > 1:
        with np.errstate(invalid="ignore"):
            du = np.abs(np.diff(u))
            m[1:] |= (du >= np.float32(diff_th)) & (np.abs(u[1:]) >= np.float32(abs_th))
    return m.astype(np.float32)


def clean_and_rule(x: np.ndarray, max_jump: float, max_abs: float, ma_window: int):
    """clean_openlab_and_rule, feature_utils.py:49-99, restated without the Python loop.  The loop marks sample i removed
    when x2[i-1] is already NaN, so the FIRST invalid sample or AND-rule hit removes everything after it; pandas'
    interpolate(limit_direction="both") then holds the last valid value to the end (all-NaN if nothing is valid), and
    np.convolve(mode="same") with a flat kernel is a centred, zero-padded moving average.  fp64 -> float32."""
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    trig = ~np.isfinite(x)
    if n > 1:
        with np.errstate(invalid="ignore"):
            trig[1:] |= (np.abs(x[1:] - x[:-1]) > float(max_jump)) & (np.abs(x[1:]) > float(max_abs))
    hits = np.flatnonzero(trig)
    i0 = int(hits[0]) if hits.size else n
    removed = np.zeros(n, dtype=bool)
    removed[i0:] = True
    xi = x.copy()
    xi[i0:] = x[i0 - 1] if i0 > 0 else np.nan
    if ma_window and ma_window > 1:
        w = int(ma_window)
        half_l, half_r = w // 2, (w - 1) // 2           # np.convolve 'same': output i sums x[i-half_r .. i+half_l] for odd w both = w//2
        pad = np.concatenate([np.zeros(half_l), xi, np.zeros(half_l)])
        k = 1.0 / float(w)
        acc = np.zeros(n, dtype=np.float64)
        for j in range(w):                              # ascending sample index, multiply then add (cblas_ddot scalar tail)
            acc = acc + pad[j:j + n] * k
        xi = acc
        del half_r
    return xi.astype(np.float32), removed.astype(np.float32)


def openlab_extract_run(raw: np.ndarray, cfg: dict = OPENLAB_EXTRACT_CFG) -> dict:
    """One run of 01_extract_windows_and_labels.py:104-236.  raw [R,4] float32 = DMS_1, LWA_2, LWA_3, LWA_4 as parsed by
    _to_float (:58-59).  Returns the kept series A_clean/A_raw [Rk,4] float32 (windows are views: start i*stride, length T),
    per-window metadata and integer labels (0 Normal, 1 Sensor Fault, 2 Structural Fault)."""
    raw = np.asarray(raw, dtype=np.float32)
    T, S = int(cfg["T"]), int(cfg["stride"])
    dms = raw[:, 0].copy()
    u = [raw[:, c].copy() for c in (1, 2, 3)]
    for a in u:
        with np.errstate(invalid="ignore"):
            a[a <= np.float32(cfg["sentinel"])] = np.nan                                  # :117-119
    out = [provider_outlier_mask(a, cfg["raw_diff_th"], cfg["raw_abs_th"]) for a in u]     # :122-124
    inv = [(~np.isfinite(a)).astype(np.float32) for a in u]
    raw_out = np.maximum.reduce(out).astype(np.float32)
    raw_inv = np.maximum.reduce(inv).astype(np.float32)
    cl = [clean_and_rule(a, cfg["clean_max_jump"], cfg["clean_max_abs"], cfg["ma_window"]) for a in u]   # :134-142
    removed = np.maximum.reduce([c[1] for c in cl]).astype(np.float32)
    A_clean = np.stack([dms, cl[0][0], cl[1][0], cl[2][0]], axis=1).astype(np.float32)
    A_raw = np.stack([dms, u[0], u[1], u[2]], axis=1).astype(np.float32)
    keep = np.isfinite(dms)                                                                # :151-156
    A_clean, A_raw, raw_out, raw_inv, removed = A_clean[keep], A_raw[keep], raw_out[keep], raw_inv[keep], removed[keep]
    n = A_clean.shape[0]
    nW = 0 if n < T else (n - T) // S + 1
    starts = np.arange(nW, dtype=np.int64) * S
    res = dict(A_clean=A_clean, A_raw=A_raw, rows_kept=n, n_windows=nW, win_start_idx=starts)
    if nW == 0:
        return res
    win = lambda a: np.stack([a[i:i + T] for i in starts]).astype(np.float32)
    Xc = win(A_clean)
    f32 = np.float32
    raw_out_ratio = win(raw_out).mean(axis=1).astype(f32)                                  # :170-176
    raw_inv_ratio = win(raw_inv).mean(axis=1).astype(f32)
    removed_ratio = win(removed).mean(axis=1).astype(f32)
    U = np.stack([Xc[:, :, j + 1 - 1] for j in cfg["struct_channels"]], axis=2)          # :181 (indices into Xc: 1..3)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        u_min = np.nanmin(U, axis=(1, 2)).astype(f32)
        u_max = np.nanmax(U, axis=(1, 2)).astype(f32)
        all_nan = (~np.isfinite(u_min)) | (~np.isfinite(u_max))
        dms_win = Xc[:, :, 0]
        dms_rng = (np.nanmax(dms_win, axis=1) - np.nanmin(dms_win, axis=1)).astype(f32)
        u_var = np.nanvar(U, axis=(1, 2)).astype(f32)
    flat = ((u_var < cfg["flat_var_eps"]) & (dms_rng > cfg["force_range_for_flatline"])).astype(np.int32)   # :193
    sensor = ((raw_inv_ratio >= float(cfg["raw_invalid_ratio_fault"])) | (raw_out_ratio > 0.0) | (removed_ratio > 0.0) |
              (flat == 1) | all_nan)                                                       # :200-206
    struct = u_max > float(cfg["allow_max"])                                               # :209
    label = np.full(nW, LABEL_NORMAL, dtype=np.int32)
    label[struct & ~sensor] = LABEL_STRUCT_FAULT
    label[sensor] = LABEL_SENSOR_FAULT                                                     # :212-214
    res.update(label=label, u_min=u_min, u_max=u_max, dms_range=dms_rng, raw_invalid_ratio=raw_inv_ratio,
               raw_outlier_ratio=raw_out_ratio, removed_ratio=removed_ratio, flatline_loadaware=flat,
               all_nan_struct=all_nan.astype(np.int32), u_var=u_var)
    return res
