"""CPU port of the reference path on torch.nn primitives (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's arithmetic IS torch.nn (nn.LSTM / Linear / LayerNorm / Conv2d / BatchNorm2d /
GroupNorm), so the faithful CPU baseline is the same library kernels (MKL /
oneDNN) driven by a restatement of the reference's module wiring and inner loops.  The reference's
Python sources cannot travel to the GPU box (/root/reference is absent there), hence a port:
bench.py's `cpu_baseline` and `--impl reference` legs time THIS on the host cores
(`cpu_baseline.kind = "port"`).  Checked against the golden vectors in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


class VaePort(nn.Module):
    """Wiring of TemporalVAE (4DOF/Scripts/Models/temporal_vae.py:14-77; 1_DOF twin without
    LayerNorm :8-58; openLAB Codes/Models/temporal_vae_model.py:4-66), built from a state_dict."""

    def __init__(self, sd: dict):
        super().__init__()
        w = sd["encoder_lstm.weight_ih_l0"]
        H, D = w.shape[0] // 4, w.shape[1]
        L = sum(1 for k in sd if k.startswith("encoder_lstm.weight_ih_l"))
        Z = sd["fc_mu.weight"].shape[0]
        self.enc = nn.LSTM(D, H, L, batch_first=True)
        self.dec = nn.LSTM(H, H, L, batch_first=True)
        self.ln = nn.LayerNorm(H) if "layer_norm.weight" in sd else None
        self.mu, self.lv = nn.Linear(H, Z), nn.Linear(H, Z)
        self.l2h, self.out = nn.Linear(Z, H), nn.Linear(H, D)
        remap = {"encoder_lstm": "enc", "decoder_lstm": "dec", "layer_norm": "ln", "fc_mu": "mu", "fc_logvar": "lv",
                 "fc_latent_to_hidden": "l2h", "output_layer": "out"}
        self.load_state_dict({remap[k.split(".")[0]] + "." + k.split(".", 1)[1]: _t(np.asarray(v)) for k, v in sd.items()})
        self.eval()

    @torch.no_grad()
    def forward(self, x: torch.Tensor, eps: torch.Tensor | None):
        _, (h_n, _) = self.enc(x)
        h = h_n[-1]
        if self.ln is not None:
            h = self.ln(h)
        mu, lv = self.mu(h), self.lv(h)
        z = mu if eps is None else mu + eps * torch.exp(0.5 * lv)
        u = torch.tanh(self.l2h(z)).unsqueeze(1).repeat(1, x.size(1), 1)     # temporal_vae.py:67-68
        y, _ = self.dec(u)
        return self.out(y), mu, lv


@torch.no_grad()
def vae_scores_batched(vae: VaePort, X: np.ndarray, eps: np.ndarray | None, batch: int) -> np.ndarray:
    """full_mse_scores_batched (4DOF/Scripts/04_vae_thresholding.py:113-124) /
    recon_mse_per_window (openLAB 10_test_hybrid_pipeline.py:240-251); eps=None draws randn like the
    reference does."""
    out = np.zeros((X.shape[0],), dtype=np.float32)
    for i in range(0, X.shape[0], batch):
        xb = torch.tensor(X[i:i + batch], dtype=torch.float32)
        e = torch.randn((xb.shape[0], vae.mu.out_features)) if eps is None else _t(eps[i:i + batch])
        xhat, _, _ = vae(xb, e)
        out[i:i + batch] = ((xb - xhat) ** 2).mean(dim=(1, 2)).numpy().astype(np.float32)
    return out


class Cnn4dofPort(nn.Module):
    """4DOF/Scripts/Models/cnn_model.py:16-34,45-51."""

    def __init__(self, sd: dict):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(2, 16, 3, padding=1), nn.BatchNorm2d(16), nn.ReLU(), nn.MaxPool2d(2))
        self.conv2 = nn.Sequential(nn.Conv2d(16, 32, 3, padding=1), nn.BatchNorm2d(32), nn.ReLU(), nn.MaxPool2d(2))
        self.fc1 = nn.Sequential(nn.Linear(2400, 128), nn.ReLU())
        self.fc2 = nn.Linear(128, 2)
        self.load_state_dict({k: _t(np.asarray(v)) for k, v in sd.items()})
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return self.fc2(self.fc1(torch.flatten(self.conv2(self.conv1(x)), 1)))


class CnnOpenLabPort(nn.Module):
    """20250506_openLAB_tests/Codes/Models/cnn_model.py:16-43,54-57."""

    def __init__(self, sd: dict):
        super().__init__()

        def block(ci, co, kt, pt):
            return nn.Sequential(nn.Conv2d(ci, co, (kt, 3), padding=(pt, 1)), nn.GroupNorm(8, co), nn.SiLU())

        self.features = nn.Sequential(block(1, 32, 7, 3), nn.MaxPool2d((2, 1)), block(32, 64, 5, 2), nn.MaxPool2d((2, 1)),
                                      block(64, 128, 5, 2), nn.MaxPool2d((2, 1)), block(128, 256, 3, 1), nn.AdaptiveAvgPool2d((1, 1)))
        self.classifier = nn.Sequential(nn.Flatten(), nn.Linear(256, 128), nn.SiLU(), nn.Identity(), nn.Linear(128, 2))
        self.load_state_dict({k: _t(np.asarray(v)) for k, v in sd.items()})
        self.eval()

    @torch.no_grad()
    def forward(self, x):
        return self.classifier(self.features(x))


@torch.no_grad()
def hybrid_4dof(vae: VaePort, cnn: Cnn4dofPort, series: np.ndarray, mean, std, thr: float, eps1=None, eps2=None,
                T: int = 100, stride: int = 1, batch: int = 512) -> dict:
    """The whole reference hot path for one group of 4DOF windows: make_windows + normalize_windows
    (06_test_full_pipeline.py:106-126) -> score loop (:338-344) -> threshold (:350-351) -> second VAE
    pass + residual stack + CNN + labels (:358-372)."""
    N = 0 if series.shape[0] < T else (series.shape[0] - T) // stride + 1
    W = np.stack([series[i:i + T] for i in range(0, series.shape[0] - T + 1, stride)], axis=0).astype(np.float32)
    Z = (W - mean[None, None, :]) / std[None, None, :]
    Z = np.nan_to_num(Z, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    score = vae_scores_batched(vae, Z, eps1, batch)
    mask = score > thr
    idx = np.where(mask)[0]
    y_pred = np.zeros((N,), dtype=np.int64)
    p_struct = np.zeros((N,), dtype=np.float32)
    logits_all = np.zeros((idx.size, 2), dtype=np.float32)
    for j in range(0, idx.size, batch):
        sel = idx[j:j + batch]
        zb = torch.tensor(Z[sel], dtype=torch.float32)
        e = torch.randn((zb.shape[0], vae.mu.out_features)) if eps2 is None else _t(eps2[j:j + zb.shape[0]])
        xhat, _, _ = vae(zb, e)
        logits = cnn(torch.stack([zb, (zb - xhat) ** 2], dim=1))
        logits_all[j:j + batch] = logits.numpy()
        y_pred[sel] = torch.argmax(logits, dim=1).numpy().astype(np.int64) + 1
        p_struct[sel] = torch.softmax(logits, dim=1).numpy().astype(np.float32)[:, 1]
    return dict(score=score, mask=mask, idx=idx, logits=logits_all, y_pred=y_pred, p_struct=p_struct, n=N)


@torch.no_grad()
def hybrid_4dof_device(vae: VaePort, cnn: Cnn4dofPort, series: np.ndarray, mean, std, thr: float, device, T: int = 100,
                       stride: int = 1, batch: int = 512) -> dict:
    """hybrid_4dof with the models on `device` (the incumbent: stock PyTorch kernels -- cuDNN LSTM, cuDNN conv -- on the
    GPU), keeping the reference's data movement: windows are built on the host, every batch goes host -> device and its
    scores / logits come back device -> host (06_test_full_pipeline.py:338-344,358-372).  Baseline timing only."""
    W = np.stack([series[i:i + T] for i in range(0, series.shape[0] - T + 1, stride)], axis=0).astype(np.float32)
    Z = np.nan_to_num((W - mean[None, None, :]) / std[None, None, :], nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    N = Z.shape[0]
    score = np.zeros((N,), np.float32)
    Zn = vae.mu.out_features
    for i in range(0, N, batch):
        xb = torch.tensor(Z[i:i + batch], dtype=torch.float32).to(device)
        xhat, _, _ = vae(xb, torch.randn((xb.shape[0], Zn), device=device))
        score[i:i + batch] = ((xb - xhat) ** 2).mean(dim=(1, 2)).detach().cpu().numpy().astype(np.float32)
    idx = np.where(score > thr)[0]
    y_pred = np.zeros((N,), np.int64)
    for j in range(0, idx.size, batch):
        sel = idx[j:j + batch]
        zb = torch.tensor(Z[sel], dtype=torch.float32).to(device)
        xhat, _, _ = vae(zb, torch.randn((zb.shape[0], Zn), device=device))
        logits = cnn(torch.stack([zb, (zb - xhat) ** 2], dim=1))
        y_pred[sel] = torch.argmax(logits, dim=1).cpu().numpy().astype(np.int64) + 1
    return dict(score=score, idx=idx, y_pred=y_pred, n=N)


@torch.no_grad()
def hybrid_openlab(vae: VaePort, cnn: CnnOpenLabPort, series: np.ndarray, chan, vmu, vsd, cmu, csd, vae_thr: float, cnn_thr: float,
                   eps=None, T: int = 200, stride: int = 20, batch: int = 256, clip: float = 10.0) -> dict:
    """openLAB hot path for one stream: windowize_2d (feature_utils.py:130-152) -> gate on the selected clean
    channels (10_test_hybrid_pipeline.py:351-367) -> CNN on the flagged raw windows with its own stats
    (stage2_predict_cnn :265-302).  The synthetic stream serves as both X_clean and X_raw."""
    W = np.asarray([series[i:i + T] for i in range(0, series.shape[0] - T + 1, stride)], dtype=np.float32)

    def standardize(X, mu, sd):
        Xn = (X - mu[None, None, :]) / sd[None, None, :]
        Xn = np.clip(Xn, -clip, clip)
        return np.nan_to_num(Xn, nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)

    Xg = standardize(W[:, :, chan], vmu, vsd)
    score = vae_scores_batched(vae, Xg, eps, batch)
    mask = score > vae_thr
    Xa = torch.tensor(standardize(W[mask], cmu, csd)[:, None, :, :], dtype=torch.float32)
    probs = []
    for i in range(0, Xa.shape[0], batch):
        probs.append(torch.softmax(cnn(Xa[i:i + batch]), dim=1)[:, 1].numpy())
    prob = np.concatenate(probs, axis=0).astype(np.float64) if probs else np.zeros((0,), np.float64)
    return dict(score=score, mask=mask, prob=prob, pred=(prob >= cnn_thr).astype(np.int64), n=W.shape[0])


# ---------------------------------------------------------------------------------------------------------
# training step (BASELINE config 5): 4DOF/Scripts/03_train_vae.py:260-271
# ---------------------------------------------------------------------------------------------------------
PARAM_ORDER_DOC = "list(model.parameters()) order of TemporalVAE (temporal_vae.py:29-49)"


def vae_param_names(sd: dict) -> list:
    """State-dict keys in list(model.parameters()) order (the flat layout of include/shmfast.h)."""
    L = sum(1 for k in sd if k.startswith("encoder_lstm.weight_ih_l"))
    names = []
    for l in range(L):
        names += [f"encoder_lstm.{k}_l{l}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    if "layer_norm.weight" in sd:
        names += ["layer_norm.weight", "layer_norm.bias"]
    names += ["fc_mu.weight", "fc_mu.bias", "fc_logvar.weight", "fc_logvar.bias", "fc_latent_to_hidden.weight",
              "fc_latent_to_hidden.bias"]
    for l in range(L):
        names += [f"decoder_lstm.{k}_l{l}" for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    names += ["output_layer.weight", "output_layer.bias"]
    return names


class VaeTrainPort(nn.Module):
    """TemporalVAE.forward in train() mode (temporal_vae.py:51-77) with the random draws made explicit: eps is
    supplied, and nn.LSTM's inter-layer dropout (temporal_vae.py:33,47) is written out as a stack of single-layer
    nn.LSTMs with supplied keep-masks (with no mask the stack is arithmetically the reference's multi-layer nn.LSTM;
    checked against the reference module's gradients in tests/test_oracle_golden.py)."""

    def __init__(self, sd: dict):
        super().__init__()
        w = sd["encoder_lstm.weight_ih_l0"]
        H, D = w.shape[0] // 4, w.shape[1]
        L = sum(1 for k in sd if k.startswith("encoder_lstm.weight_ih_l"))
        Z = sd["fc_mu.weight"].shape[0]
        self.L = L
        self.enc = nn.ModuleList([nn.LSTM(D if l == 0 else H, H, 1, batch_first=True) for l in range(L)])
        self.dec = nn.ModuleList([nn.LSTM(H, H, 1, batch_first=True) for l in range(L)])
        self.ln = nn.LayerNorm(H) if "layer_norm.weight" in sd else None
        self.mu, self.lv = nn.Linear(H, Z), nn.Linear(H, Z)
        self.l2h, self.out = nn.Linear(Z, H), nn.Linear(H, D)
        self.names = vae_param_names(sd)
        with torch.no_grad():
            for name in self.names:
                self.param(name).copy_(_t(np.asarray(sd[name], dtype=np.float32)))

    def param(self, name: str) -> nn.Parameter:
        mod, rest = name.split(".", 1)
        if mod in ("encoder_lstm", "decoder_lstm"):
            kind, l = rest.rsplit("_l", 1)
            return getattr((self.enc if mod == "encoder_lstm" else self.dec)[int(l)], kind + "_l0")
        m = {"layer_norm": self.ln, "fc_mu": self.mu, "fc_logvar": self.lv, "fc_latent_to_hidden": self.l2h,
             "output_layer": self.out}[mod]
        return getattr(m, rest)

    def ordered_parameters(self) -> list:
        return [self.param(n) for n in self.names]

    def _stack(self, layers, x, masks, p):
        for l, lstm in enumerate(layers):
            x, (h_n, _) = lstm(x)
            if l < self.L - 1 and masks is not None:
                x = x * masks[l].to(x.dtype) / (1.0 - p)
        return x, h_n[-1]

    def forward(self, x, eps, drop_enc=None, drop_dec=None, p: float = 0.0):
        _, h = self._stack(self.enc, x, drop_enc, p)
        if self.ln is not None:
            h = self.ln(h)
        mu, lv = self.mu(h), self.lv(h)
        z = mu + eps * torch.exp(0.5 * lv)
        u = torch.tanh(self.l2h(z)).unsqueeze(1).repeat(1, x.size(1), 1)
        y, _ = self._stack(self.dec, u, drop_dec, p)
        return self.out(y), mu, lv


def train_step_port(port: VaeTrainPort, opt, x, eps, kl_w: float, drop_enc=None, drop_dec=None, p: float = 0.0,
                    max_norm: float = 2.0, world: int = 1):
    """03_train_vae.py:262-270 on the port.  Returns (loss3, flat gradient before clipping, total_norm)."""
    xhat, mu, logvar = port(x, eps, drop_enc, drop_dec, p)
    recon = F.mse_loss(xhat, x, reduction="mean")
    kl = -0.5 * torch.mean(1.0 + logvar - mu.pow(2) - logvar.exp())
    loss = recon + kl_w * kl
    opt.zero_grad(set_to_none=True)
    loss.backward()
    params = port.ordered_parameters()
    flat_g = torch.cat([q.grad.reshape(-1) for q in params]).clone()
    total = torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm) if max_norm > 0 else torch.linalg.vector_norm(flat_g)
    opt.step()
    return np.array([loss.item(), recon.item(), kl.item()]), flat_g.numpy(), float(total)


def flat_params(port: VaeTrainPort) -> np.ndarray:
    return torch.cat([q.detach().reshape(-1) for q in port.ordered_parameters()]).numpy().copy()


# ---------------------------------------------------------------------------------------------------------
# CNN training steps: 4DOF/Scripts/05_train_cnn.py:266-281 and
# 20250506_openLAB_tests/Codes/06_train_cnn.py:410-421, with nn.Dropout's mask made an explicit input.
# ---------------------------------------------------------------------------------------------------------
CNN4DOF_PARAMS = ["conv1.0.weight", "conv1.0.bias", "conv1.1.weight", "conv1.1.bias", "conv2.0.weight", "conv2.0.bias",
                  "conv2.1.weight", "conv2.1.bias", "fc1.0.weight", "fc1.0.bias", "fc2.weight", "fc2.bias"]
CNNOL_PARAMS = [f"features.{i}.{j}.{k}" for i in (0, 2, 4, 6) for j in (0, 1) for k in ("weight", "bias")] + \
               ["classifier.1.weight", "classifier.1.bias", "classifier.4.weight", "classifier.4.bias"]


class CnnTrainPort(nn.Module):
    """The two CNNs in train() mode written with torch.nn.functional on explicit parameters (list(model.parameters())
    order = CNN4DOF_PARAMS / CNNOL_PARAMS): BatchNorm2d uses batch statistics and updates the running ones (momentum 0.1),
    GroupNorm(8) + SiLU for openLAB, Dropout = x * keep_mask / (1 - p) with the mask supplied."""

    def __init__(self, arch: str, sd: dict):
        super().__init__()
        self.arch = arch
        self.names = CNN4DOF_PARAMS if arch == "4dof" else CNNOL_PARAMS
        self.p = nn.ParameterList([nn.Parameter(_t(np.asarray(sd[n], dtype=np.float32)).clone()) for n in self.names])
        if arch == "4dof":
            self.running = [_t(np.asarray(sd[f"conv{b}.1.running_{s}"], dtype=np.float32)).clone() for b in (1, 2) for s in ("mean", "var")]

    def ordered_parameters(self) -> list:
        return list(self.p)

    def forward(self, x, mask=None, p_drop: float = 0.0):
        P = list(self.p)
        if self.arch == "4dof":
            for b in range(2):
                w, bias, g, be = P[4 * b:4 * b + 4]
                x = F.conv2d(x, w, bias, padding=1)
                x = F.batch_norm(x, self.running[2 * b], self.running[2 * b + 1], g, be, training=True, momentum=0.1, eps=1e-5)
                x = F.max_pool2d(F.relu(x), 2)
            x = F.relu(F.linear(torch.flatten(x, 1), P[8], P[9]))
        else:
            pads = ((3, 1), (2, 1), (2, 1), (1, 1))
            for b in range(4):
                w, bias, g, be = P[4 * b:4 * b + 4]
                x = F.silu(F.group_norm(F.conv2d(x, w, bias, padding=pads[b]), 8, g, be, eps=1e-5))
                x = F.max_pool2d(x, (2, 1)) if b < 3 else F.adaptive_avg_pool2d(x, (1, 1))
            x = F.silu(F.linear(torch.flatten(x, 1), P[16], P[17]))
        if mask is not None and p_drop > 0:
            x = x * mask.to(x.dtype) / (1.0 - p_drop)
        return F.linear(x, P[-2], P[-1])


def focal_loss(logits, targets, alpha, gamma: float):
    """WeightedFocalLoss.forward, 06_train_cnn.py:202-207."""
    ce = F.cross_entropy(logits, targets, reduction="none")
    pt = torch.exp(-ce)
    return (alpha[targets] * ((1 - pt) ** gamma) * ce).mean()


def cnn_train_step_port(port: CnnTrainPort, opt, x, y, mask=None, p_drop: float = 0.0, alpha=None, gamma: float = 0.0,
                        max_norm: float = 0.0):
    """One optimisation step on the port.  Returns (logits, loss, flat gradient before clipping, total_norm)."""
    logits = port(x, mask, p_drop)
    loss = F.cross_entropy(logits, y) if (alpha is None and gamma == 0.0) else focal_loss(logits, y, alpha, gamma)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    params = port.ordered_parameters()
    flat_g = torch.cat([q.grad.reshape(-1) for q in params]).clone()
    total = torch.nn.utils.clip_grad_norm_(params, max_norm=max_norm) if max_norm > 0 else torch.linalg.vector_norm(flat_g)
    opt.step()
    return logits.detach().numpy(), float(loss.item()), flat_g.numpy(), float(total)
