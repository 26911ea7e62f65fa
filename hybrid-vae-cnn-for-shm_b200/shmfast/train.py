"""LSTM-VAE training step on libshmfast (BASELINE config 5; 4DOF/Scripts/03_train_vae.py:214-271).

Two ways in, both running the same hand-written kernels (csrc/train.cu):

* the reference's own loop, unchanged -- `vae.train(); xhat, mu, logvar = vae(xb); loss.backward();
  clip_grad_norm_(...); opt.step()` -- through `VaeTrainFunction`, the autograd bridge the `Models/` shims
  use in train() mode (forward = shm_vae_train_forward, backward = shm_vae_train_backward);
* `VaeTrainer.step(xb, kl_w)`: the fused step on flat buffers: forward -> ELBO + upstream gradients ->
  BPTT -> ONE all-reduce of the flat gradient (data parallel over torch.distributed; NCCL on GPUs) ->
  global-norm clip + Adam in one kernel.  No CPU fallback anywhere.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import ShmfastError, check
from .ops import _need_cuda, _ptr, _stream


def kl_anneal_sigmoid(epoch: int, n_epochs: int, start: float = 0.0, stop: float = 1.0, anneal_ratio: float = 0.3) -> float:
    """03_train_vae.py:120-135 (epoch is 1-based)."""
    e0 = epoch - 1
    warm = max(1, int(n_epochs * anneal_ratio))
    x = (e0 - warm) / float(max(warm, 1))
    return float(stop / (1.0 + math.exp(-x * 5.0)))


def vae_cfg_of(model) -> _lib.VaeCfg:
    has_ln = hasattr(model, "layer_norm")
    return _lib.VaeCfg(model.input_dim, model.hidden_dim, model.latent_dim, model.num_layers, 1 if has_ln else 0,
                       float(model.layer_norm.eps) if has_ln else 1e-5, 0)


class VaeTrainHandle:
    """shm_vae_trainer: activation workspace for (T, max_batch) + forward / backward entry points."""

    def __init__(self, cfg: _lib.VaeCfg, T: int, max_batch: int, device: torch.device):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ShmfastError("VaeTrainHandle needs a CUDA device: libshmfast has no CPU fallback")
        self.cfg, self.T, self.max_batch = cfg, int(T), int(max_batch)
        self.n_params = int(self._lib.shm_vae_param_count(C.byref(cfg)))
        if self.n_params <= 0:
            raise ShmfastError("bad VAE configuration")
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            check(self._lib.shm_vae_trainer_create(C.byref(h), C.byref(cfg), self.T, self.max_batch, idx), "shm_vae_trainer_create")
        self._h = h
        self.generation = 0            # id of the forward whose activations the workspace currently holds

    def forward(self, params: torch.Tensor, x: torch.Tensor, eps: torch.Tensor, drop_enc: Optional[torch.Tensor] = None,
                drop_dec: Optional[torch.Tensor] = None, drop_p: float = 0.0, want_xhat: bool = True):
        for t, name in ((params, "params"), (x, "x"), (eps, "eps")):
            _need_cuda(t, name)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ShmfastError(f"{name} must be contiguous float32")
        B, T, D = x.shape
        if T != self.T or D != self.cfg.D or B > self.max_batch or params.numel() != self.n_params:
            raise ShmfastError(f"shape mismatch: x {tuple(x.shape)} vs trainer (T={self.T}, D={self.cfg.D}, max_batch={self.max_batch})")
        if eps.numel() != B * self.cfg.Z:
            raise ShmfastError("eps must be [B, Z]")
        for m in (drop_enc, drop_dec):
            if m is not None and (m.dtype != torch.uint8 or not m.is_contiguous() or m.numel() != (self.cfg.L - 1) * B * T * self.cfg.H):
                raise ShmfastError("dropout keep-masks must be contiguous uint8 [L-1, B, T, H]")
        xhat = torch.empty_like(x) if want_xhat else None
        mu = torch.empty((B, self.cfg.Z), dtype=torch.float32, device=x.device)
        lv = torch.empty_like(mu)
        with torch.cuda.device(x.device):
            check(self._lib.shm_vae_train_forward(self._h, _ptr(params), _ptr(x), B, _ptr(eps), _ptr(drop_enc), _ptr(drop_dec),
                                                  float(drop_p), _ptr(xhat), _ptr(mu), _ptr(lv), _stream()), "shm_vae_train_forward")
        self.generation += 1
        return xhat, mu, lv

    def backward(self, params: torch.Tensor, d_xhat: torch.Tensor, d_mu: Optional[torch.Tensor], d_lv: Optional[torch.Tensor],
                 grads: Optional[torch.Tensor] = None) -> torch.Tensor:
        if grads is None:
            grads = torch.empty_like(params)
        with torch.cuda.device(params.device):
            check(self._lib.shm_vae_train_backward(self._h, _ptr(params), _ptr(d_xhat), _ptr(d_mu), _ptr(d_lv), _ptr(grads),
                                                   _stream()), "shm_vae_train_backward")
        return grads

    def close(self):
        if getattr(self, "_h", None):
            self._lib.shm_vae_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def elbo_grad(x: torch.Tensor, xhat: torch.Tensor, mu: torch.Tensor, lv: torch.Tensor, kl_w: float, want_grads: bool = True):
    """loss3 = [recon + kl_w*kl, recon, kl] (03_train_vae.py:264-266) and d loss / d(xhat, mu, logvar)."""
    lib = _lib.load()
    loss3 = torch.empty((3,), dtype=torch.float32, device=x.device)
    d_xhat = torch.empty_like(xhat) if want_grads else None
    d_mu = torch.empty_like(mu) if want_grads else None
    d_lv = torch.empty_like(lv) if want_grads else None
    with torch.cuda.device(x.device):
        check(lib.shm_vae_elbo_grad(_ptr(x), _ptr(xhat), _ptr(mu), _ptr(lv), x.numel(), mu.numel(), float(kl_w), _ptr(d_xhat),
                                    _ptr(d_mu), _ptr(d_lv), _ptr(loss3), _stream()), "shm_vae_elbo_grad")
    return loss3, d_xhat, d_mu, d_lv


def adam_clip_step(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
                   lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, max_norm: float = 0.0,
                   grad_scale: float = 1.0) -> torch.Tensor:
    """clip_grad_norm_ + Adam.step over flat buffers; returns norm2 = [sum g^2, total_norm] (device)."""
    lib = _lib.load()
    norm2 = torch.empty((2,), dtype=torch.float32, device=params.device)
    with torch.cuda.device(params.device):
        check(lib.shm_adam_clip_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), int(step), float(lr),
                                     float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(max_norm),
                                     float(grad_scale), _ptr(norm2), _stream()), "shm_adam_clip_step")
    return norm2


def draw_dropout_masks(L: int, B: int, T: int, H: int, p: float, device, generator=None):
    """Keep-masks for nn.LSTM's inter-layer dropout (temporal_vae.py:33,47): uint8 [L-1, B, T, H] x 2."""
    if L < 2 or p <= 0.0:
        return None, None
    shape = (L - 1, B, T, H)
    enc = (torch.rand(shape, device=device, generator=generator) >= p).to(torch.uint8)
    dec = (torch.rand(shape, device=device, generator=generator) >= p).to(torch.uint8)
    return enc, dec


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def broadcast_parameters(flat: torch.Tensor, group=None) -> None:
    """Identical replicas at start (what DistributedDataParallel's constructor does): rank 0 of the group wins."""
    if world_size(group) > 1:
        dist.broadcast(flat, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)


def reduce_gradients(flat_grads: torch.Tensor, group=None) -> float:
    """The ONE collective of the data-parallel step: in-place SUM all-reduce of the flat gradient.  Returns the
    scale (1/world) that turns the sum into the mean; it is applied inside the optimiser kernel."""
    world = world_size(group)
    if world > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class VaeTrainFunction(torch.autograd.Function):
    """Autograd bridge: (x, eps, masks, *parameters) -> (xhat, mu, logvar)."""

    @staticmethod
    def forward(ctx, handle: VaeTrainHandle, x, eps, drop_enc, drop_dec, drop_p, *params):
        flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        xhat, mu, lv = handle.forward(flat, x.detach().contiguous(), eps.contiguous(), drop_enc, drop_dec, drop_p)
        ctx.handle, ctx.flat = handle, flat
        ctx.generation = handle.generation          # the single workspace holds THIS forward's activations
        ctx.shapes = [p.shape for p in params]
        ctx.x_shape = x.shape
        return xhat, mu, lv

    @staticmethod
    def backward(ctx, d_xhat, d_mu, d_lv):
        if ctx.generation != ctx.handle.generation:
            raise ShmfastError("backward of a stale graph: another train-mode forward of this model overwrote the activation "
                               "workspace (one backward per forward; call backward before the next forward)")
        if d_xhat is None:
            d_xhat = ctx.flat.new_zeros(ctx.x_shape)
        g = ctx.handle.backward(ctx.flat, d_xhat.contiguous().float(), None if d_mu is None else d_mu.contiguous().float(),
                                None if d_lv is None else d_lv.contiguous().float())
        outs, o = [], 0
        for s in ctx.shapes:
            n = int(torch.Size(s).numel())
            outs.append(g[o:o + n].view(s))
            o += n
        return (None, None, None, None, None, None, *outs)


class VaeTrainer:
    """Fused data-parallel training step for a `Models/` TemporalVAE shim.

    The model's parameters are re-pointed at views of ONE flat buffer (list(model.parameters()) order), so the
    gradient all-reduce is a single collective over 4*n_params bytes (1.9 MB for the 4DOF model) and the
    optimiser is a single kernel.  Hyper-parameters default to 03_train_vae.py (Adam lr 1e-3, weight_decay 1e-5,
    clip 2.0).  Every rank must call step() with its own shard of the global batch."""

    def __init__(self, model, seq_len: int, max_batch: int, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-5, max_norm: float = 2.0, dropout: Optional[float] = None, process_group=None):
        params = list(model.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise ShmfastError("move the model to a CUDA device first (no CPU fallback)")
        self.model, self.device, self.group = model, dev, process_group
        self.handle = VaeTrainHandle(vae_cfg_of(model), seq_len, max_batch, dev)
        flat = torch.cat([p.detach().reshape(-1).float() for p in params]).contiguous()
        if flat.numel() != self.handle.n_params:
            raise ShmfastError("model parameter count does not match the TemporalVAE layout")
        o = 0
        for p in params:
            n = p.numel()
            p.data = flat[o:o + n].view_as(p)
            o += n
        self.flat = flat
        self.grads = torch.zeros_like(flat)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.steps = 0
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.dropout = float(getattr(model.encoder_lstm, "dropout", 0.0)) if dropout is None else float(dropout)
        self.last_norm = None
        broadcast_parameters(self.flat, self.group)

    def step(self, xb: torch.Tensor, kl_w: float, eps: Optional[torch.Tensor] = None, masks=None) -> torch.Tensor:
        """One optimisation step on this rank's batch shard; returns loss3 = [total, recon, kl] (device)."""
        _need_cuda(xb, "xb")
        xb = xb.to(torch.float32).contiguous()
        B, T, _ = xb.shape
        cfg = self.handle.cfg
        if eps is None:
            eps = torch.randn((B, cfg.Z), dtype=torch.float32, device=xb.device)
        if masks is None:
            masks = draw_dropout_masks(cfg.L, B, T, cfg.H, self.dropout if self.model.training else 0.0, xb.device)
        xhat, mu, lv = self.handle.forward(self.flat, xb, eps, masks[0], masks[1], self.dropout)
        loss3, d_xhat, d_mu, d_lv = elbo_grad(xb, xhat, mu, lv, kl_w)
        self.handle.backward(self.flat, d_xhat, d_mu, d_lv, self.grads)
        scale = reduce_gradients(self.grads, self.group)
        self.steps += 1
        self.last_norm = adam_clip_step(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.steps, self.lr, self.betas,
                                        self.eps, self.wd, self.max_norm, scale)
        if hasattr(self.model, "_sig"):
            self.model._sig = None                 # the scorer handle must re-pack the updated weights
        return loss3

    def close(self):
        self.handle.close()
