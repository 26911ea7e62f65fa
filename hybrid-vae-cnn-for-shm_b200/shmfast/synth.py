"""Deterministic synthetic weights / windows of the reference's shapes.

Everything is drawn from numpy's PCG64 so the CPU checker, the GPU tests and bench.py see
bit-identical inputs without shipping tensors.  Distributions follow the reference modules'
initialisers (torch default U(-1/sqrt(H), 1/sqrt(H)) for LSTM/Linear/LayerNorm=1,0;
xavier_uniform for the 4DOF CNN, 4DOF/Scripts/Models/cnn_model.py:39-43; kaiming_normal(relu)
for the openLAB CNN, 20250506_openLAB_tests/Codes/Models/cnn_model.py:47-52).  State-dict key
names are the reference's.
"""
from __future__ import annotations

import math

import numpy as np

# per-stage shapes: SURVEY.md section 0 table (reference file:line cited there)
STAGES = {
    "4dof": dict(T=100, stride=1, D=12, H=128, Z=16, L=2, has_ln=True, batch=512),
    "openlab": dict(T=200, stride=20, D=3, H=64, Z=8, L=1, has_ln=True, batch=256),
    "1dof": dict(T=80, stride=1, D=12, H=32, Z=5, L=2, has_ln=False, batch=1 << 30),
}


def _u(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def vae_weights(D: int, H: int, Z: int, L: int, has_ln: bool, seed: int = 0, scale: float = 1.0) -> dict:
    """Random-init LSTM-VAE state_dict.  `scale` > 1 widens the weights to mimic trained,
    numerically harsher models (larger pre-activations, saturating gates)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = scale / math.sqrt(H)
    sd = {}
    for name, d_in0 in (("encoder_lstm", D), ("decoder_lstm", H)):
        for l in range(L):
            d_in = d_in0 if l == 0 else H
            sd[f"{name}.weight_ih_l{l}"] = _u(rng, (4 * H, d_in), k)
            sd[f"{name}.weight_hh_l{l}"] = _u(rng, (4 * H, H), k)
            sd[f"{name}.bias_ih_l{l}"] = _u(rng, (4 * H,), k)
            sd[f"{name}.bias_hh_l{l}"] = _u(rng, (4 * H,), k)
    if has_ln:
        sd["layer_norm.weight"] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
        sd["layer_norm.bias"] = (0.1 * rng.standard_normal(H)).astype(np.float32)
    sd["fc_mu.weight"] = _u(rng, (Z, H), k)
    sd["fc_mu.bias"] = _u(rng, (Z,), k)
    sd["fc_logvar.weight"] = _u(rng, (Z, H), k)
    sd["fc_logvar.bias"] = _u(rng, (Z,), k)
    kz = scale / math.sqrt(Z)
    sd["fc_latent_to_hidden.weight"] = _u(rng, (H, Z), kz)
    sd["fc_latent_to_hidden.bias"] = _u(rng, (H,), kz)
    sd["output_layer.weight"] = _u(rng, (D, H), k)
    sd["output_layer.bias"] = _u(rng, (D,), k)
    return sd


def stage_vae_weights(stage: str, seed: int = 0, scale: float = 1.0) -> dict:
    s = STAGES[stage]
    return vae_weights(s["D"], s["H"], s["Z"], s["L"], s["has_ln"], seed, scale)


def _xavier(rng, shape):
    if len(shape) == 4:
        rf = shape[2] * shape[3]
        fan_in, fan_out = shape[1] * rf, shape[0] * rf
    else:
        fan_out, fan_in = shape
    a = math.sqrt(6.0 / (fan_in + fan_out))
    return _u(rng, shape, a)


def cnn4dof_weights(seed: int = 0) -> dict:
    """4DOF CNN(2,2) state_dict; BN running stats randomised (mean~N(0,1)*0.3, var~U(0.5,2)) so
    that BN folding is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    sd = {}
    for blk, (co, ci) in (("conv1", (16, 2)), ("conv2", (32, 16))):
        sd[f"{blk}.0.weight"] = _xavier(rng, (co, ci, 3, 3))
        sd[f"{blk}.0.bias"] = (0.05 * rng.standard_normal(co)).astype(np.float32)
        sd[f"{blk}.1.weight"] = (1.0 + 0.2 * rng.standard_normal(co)).astype(np.float32)
        sd[f"{blk}.1.bias"] = (0.1 * rng.standard_normal(co)).astype(np.float32)
        sd[f"{blk}.1.running_mean"] = (0.3 * rng.standard_normal(co)).astype(np.float32)
        sd[f"{blk}.1.running_var"] = rng.uniform(0.5, 2.0, size=co).astype(np.float32)
        sd[f"{blk}.1.num_batches_tracked"] = np.array(7, dtype=np.int64)
    sd["fc1.0.weight"] = _xavier(rng, (128, 32 * 25 * 3))
    sd["fc1.0.bias"] = (0.05 * rng.standard_normal(128)).astype(np.float32)
    sd["fc2.weight"] = _xavier(rng, (2, 128))
    sd["fc2.bias"] = (0.05 * rng.standard_normal(2)).astype(np.float32)
    return sd


OPENLAB_CNN_SHAPES = ((0, (32, 1, 7, 3)), (2, (64, 32, 5, 3)), (4, (128, 64, 5, 3)), (6, (256, 128, 3, 3)))


def cnnol_weights(seed: int = 0) -> dict:
    """openLAB CNN state_dict (kaiming_normal fan_in/relu; GroupNorm affine perturbed from 1/0)."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000))
    sd = {}
    for idx, shp in OPENLAB_CNN_SHAPES:
        fan_in = shp[1] * shp[2] * shp[3]
        sd[f"features.{idx}.0.weight"] = (rng.standard_normal(shp) * math.sqrt(2.0 / fan_in)).astype(np.float32)
        sd[f"features.{idx}.0.bias"] = (0.05 * rng.standard_normal(shp[0])).astype(np.float32)
        sd[f"features.{idx}.1.weight"] = (1.0 + 0.2 * rng.standard_normal(shp[0])).astype(np.float32)
        sd[f"features.{idx}.1.bias"] = (0.1 * rng.standard_normal(shp[0])).astype(np.float32)
    sd["classifier.1.weight"] = (rng.standard_normal((128, 256)) * math.sqrt(2.0 / 256)).astype(np.float32)
    sd["classifier.1.bias"] = (0.05 * rng.standard_normal(128)).astype(np.float32)
    sd["classifier.4.weight"] = (rng.standard_normal((2, 128)) * math.sqrt(2.0 / 128)).astype(np.float32)
    sd["classifier.4.bias"] = (0.05 * rng.standard_normal(2)).astype(np.float32)
    return sd


def windows(N: int, T: int, D: int, seed: int = 0, amp: float = 1.0) -> np.ndarray:
    """[N,T,D] ~ N(0, amp^2) fp32 materialised windows."""
    rng = np.random.Generator(np.random.PCG64(seed + 3000))
    return (amp * rng.standard_normal((N, T, D), dtype=np.float32)).astype(np.float32)


def series(rows: int, d_all: int, seed: int = 0, nan_frac: float = 0.0) -> np.ndarray:
    """[rows,d_all] fp32 sensor series: smooth-ish random walk + noise, optional NaN runs
    (the openLAB X_raw carries NaN in 1.35 % of windows)."""
    rng = np.random.Generator(np.random.PCG64(seed + 4000))
    x = rng.standard_normal((rows, d_all)).astype(np.float32)
    x = (0.7 * x + 0.3 * np.cumsum(x, axis=0) / np.sqrt(np.arange(1, rows + 1, dtype=np.float32))[:, None]).astype(np.float32)
    if nan_frac > 0:
        n_runs = max(1, int(rows * nan_frac / 8))
        starts = rng.integers(0, max(1, rows - 8), size=n_runs)
        chans = rng.integers(0, d_all, size=n_runs)
        for s, c in zip(starts, chans):
            x[s:s + int(rng.integers(1, 9)), c] = np.nan
    return x


def eps(N: int, Z: int, seed: int = 0) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed + 5000))
    return rng.standard_normal((N, Z), dtype=np.float32)


def stats(D: int, seed: int = 0, zero_std_channel: int | None = None):
    """Per-channel (mean, std) fp32; optionally one exactly-zero std to exercise the guards."""
    rng = np.random.Generator(np.random.PCG64(seed + 6000))
    mean = (0.2 * rng.standard_normal(D)).astype(np.float32)
    std = rng.uniform(0.5, 1.5, size=D).astype(np.float32)
    if zero_std_channel is not None:
        std[zero_std_channel] = 0.0
    return mean, std
