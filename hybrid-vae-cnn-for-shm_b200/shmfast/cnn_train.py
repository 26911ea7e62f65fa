"""CNN training steps on libshmfast (csrc/cnn_train.cu).

* 4DOF/Scripts/05_train_cnn.py:266-281 -- `logits = model(xb); loss = CrossEntropyLoss()(logits, yb); loss.backward(); optimizer.step()`
  (Adam lr 1e-4, weight_decay 5e-5) on `CNN(2, 2, 0.5)` in train() mode (BatchNorm batch statistics, running-stat update, Dropout);
* 20250506_openLAB_tests/Codes/06_train_cnn.py:410-421 -- weighted focal loss (gamma 2), `clip_grad_norm_(2.0)`, AdamW lr 3e-4 wd 1e-4.

Two ways in, as for the VAE (shmfast/train.py): the reference's own loop through `CnnTrainFunction` (what the `Models/` CNN shims use
in train() mode), or `CnnTrainer.step(xb, yb)`: forward -> loss + d logits -> backward -> ONE all-reduce of the flat gradient (data
parallel) -> clip + Adam / AdamW in one kernel.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import CNN_4DOF, CNN_OPENLAB, ShmfastError, check
from .ops import _need_cuda, _ptr, _stream
from .train import broadcast_parameters, reduce_gradients

IN_SHAPE = {CNN_4DOF: (2, 100, 12), CNN_OPENLAB: (1, 200, 4)}
HID = 128


class CnnTrainHandle:
    """shm_cnn_trainer: activation workspace for max_batch + forward / backward entry points."""

    def __init__(self, arch: int, max_batch: int, device: torch.device):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ShmfastError("CnnTrainHandle needs a CUDA device: libshmfast has no CPU fallback")
        self.arch, self.max_batch = int(arch), int(max_batch)
        self.n_params = int(self._lib.shm_cnn_param_count(self.arch))
        if self.n_params <= 0:
            raise ShmfastError("bad CNN architecture id")
        h = C.c_void_p()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(self.device):
            check(self._lib.shm_cnn_trainer_create(C.byref(h), self.arch, self.max_batch, idx), "shm_cnn_trainer_create")
        self._h = h
        self.generation = 0

    def forward(self, params: torch.Tensor, x: torch.Tensor, bn_running: Optional[torch.Tensor] = None, momentum: float = 0.1,
                drop_mask: Optional[torch.Tensor] = None, drop_p: float = 0.0) -> torch.Tensor:
        for t, name in ((params, "params"), (x, "x")):
            _need_cuda(t, name)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ShmfastError(f"{name} must be contiguous float32")
        B = x.shape[0]
        if tuple(x.shape[1:]) != IN_SHAPE[self.arch] or B > self.max_batch or params.numel() != self.n_params:
            raise ShmfastError(f"shape mismatch: x {tuple(x.shape)} vs trainer (arch {self.arch}, max_batch {self.max_batch})")
        if drop_mask is not None and (drop_mask.dtype != torch.uint8 or not drop_mask.is_contiguous() or drop_mask.numel() != B * HID):
            raise ShmfastError("dropout keep-mask must be contiguous uint8 [B, 128]")
        if bn_running is not None and (self.arch != CNN_4DOF or bn_running.dtype != torch.float32 or bn_running.numel() != 96
                                       or not bn_running.is_contiguous()):
            raise ShmfastError("bn_running must be contiguous float32 [96] (4DOF CNN only)")
        logits = torch.empty((B, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            check(self._lib.shm_cnn_train_forward(self._h, _ptr(params), _ptr(x), B, _ptr(bn_running), float(momentum), _ptr(drop_mask),
                                                  float(drop_p), _ptr(logits), _stream()), "shm_cnn_train_forward")
        self.generation += 1
        self._x = x                       # the kernels read x again in the backward pass
        return logits

    def backward(self, params: torch.Tensor, d_logits: torch.Tensor, grads: Optional[torch.Tensor] = None) -> torch.Tensor:
        if grads is None:
            grads = torch.empty_like(params)
        with torch.cuda.device(params.device):
            check(self._lib.shm_cnn_train_backward(self._h, _ptr(params), _ptr(d_logits), _ptr(grads), _stream()), "shm_cnn_train_backward")
        return grads

    def close(self):
        if getattr(self, "_h", None):
            self._lib.shm_cnn_trainer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def cnn_loss_grad(logits: torch.Tensor, targets: torch.Tensor, alpha: Optional[torch.Tensor] = None, gamma: float = 0.0,
                  want_grad: bool = True):
    """(loss [1], d loss / d logits): nn.CrossEntropyLoss (alpha None, gamma 0; 05_train_cnn.py:257) or WeightedFocalLoss
    (06_train_cnn.py:195-207)."""
    lib = _lib.load()
    _need_cuda(logits, "logits")
    logits = logits.to(torch.float32).contiguous()
    targets = targets.to(device=logits.device, dtype=torch.int64).contiguous()
    if alpha is not None:
        alpha = alpha.to(device=logits.device, dtype=torch.float32).contiguous()
    loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
    d = torch.empty_like(logits) if want_grad else None
    with torch.cuda.device(logits.device):
        check(lib.shm_cnn_loss_grad(_ptr(logits), _ptr(targets), logits.shape[0], _ptr(alpha), float(gamma), _ptr(d), _ptr(loss), _stream()),
              "shm_cnn_loss_grad")
    return loss, d


def adam_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
              max_norm: float = 0.0, grad_scale: float = 1.0, decoupled: bool = False) -> torch.Tensor:
    """clip_grad_norm_ + Adam.step (weight decay in the gradient) or AdamW.step (decoupled) over flat buffers."""
    lib = _lib.load()
    fn = lib.shm_adamw_clip_step if decoupled else lib.shm_adam_clip_step
    norm2 = torch.empty((2,), dtype=torch.float32, device=params.device)
    with torch.cuda.device(params.device):
        check(fn(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), params.numel(), int(step), float(lr), float(betas[0]),
                 float(betas[1]), float(eps), float(weight_decay), float(max_norm), float(grad_scale), _ptr(norm2), _stream()),
              "shm_adamw_clip_step" if decoupled else "shm_adam_clip_step")
    return norm2


def draw_dropout_mask(B: int, p: float, device, generator=None) -> Optional[torch.Tensor]:
    """Keep-mask of the Dropout after fc1 (cnn_model.py:31 / openLAB :40): uint8 [B, 128]."""
    if p <= 0.0:
        return None
    return (torch.rand((B, HID), device=device, generator=generator) >= p).to(torch.uint8)


class CnnTrainFunction(torch.autograd.Function):
    """Autograd bridge: (x, *parameters) -> logits, gradients from shm_cnn_train_backward."""

    @staticmethod
    def forward(ctx, handle: CnnTrainHandle, x, bn_running, momentum, mask, p, *params):
        flat = torch.cat([q.detach().reshape(-1) for q in params]).contiguous()
        logits = handle.forward(flat, x.detach().contiguous(), bn_running, momentum, mask, p)
        ctx.handle, ctx.flat, ctx.generation = handle, flat, handle.generation
        ctx.shapes = [q.shape for q in params]
        return logits

    @staticmethod
    def backward(ctx, d_logits):
        if ctx.generation != ctx.handle.generation:
            raise ShmfastError("backward of a stale graph: another train-mode forward of this model overwrote the activation workspace")
        g = ctx.handle.backward(ctx.flat, d_logits.contiguous().float())
        outs, o = [], 0
        for s in ctx.shapes:
            n = int(torch.Size(s).numel())
            outs.append(g[o:o + n].view(s))
            o += n
        return (None, None, None, None, None, None, *outs)


class CnnTrainer:
    """Fused (data-parallel) training step for a `Models/` CNN shim.  Hyper-parameters default to the reference scripts:
    arch 4DOF -> CrossEntropy + Adam(lr 1e-4, weight_decay 5e-5), no clipping (05_train_cnn.py:31-33,256-257);
    arch openLAB -> WeightedFocalLoss(alpha, gamma 2) + clip 2.0 + AdamW(lr 3e-4, weight_decay 1e-4) (06_train_cnn.py:54-58,395-396)."""

    def __init__(self, model, max_batch: int, alpha: Optional[torch.Tensor] = None, lr: Optional[float] = None,
                 weight_decay: Optional[float] = None, process_group=None):
        params = list(model.parameters())
        dev = params[0].device
        if dev.type != "cuda":
            raise ShmfastError("move the model to a CUDA device first (no CPU fallback)")
        self.model, self.device, self.group = model, dev, process_group
        self.arch = model.ARCH
        self.handle = CnnTrainHandle(self.arch, max_batch, dev)
        flat = torch.cat([q.detach().reshape(-1).float() for q in params]).contiguous()
        if flat.numel() != self.handle.n_params:
            raise ShmfastError("model parameter count does not match the CNN layout")
        o = 0
        for q in params:
            n = q.numel()
            q.data = flat[o:o + n].view_as(q)
            o += n
        self.flat = flat
        self.grads = torch.zeros_like(flat)
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.steps = 0
        if self.arch == CNN_4DOF:
            self.lr, self.wd, self.decoupled, self.max_norm, self.gamma = (1e-4 if lr is None else lr), (5e-5 if weight_decay is None else weight_decay), False, 0.0, 0.0
            run = torch.cat([model.conv1[1].running_mean, model.conv1[1].running_var, model.conv2[1].running_mean,
                             model.conv2[1].running_var]).contiguous()
            for buf, (a, b) in ((model.conv1[1].running_mean, (0, 16)), (model.conv1[1].running_var, (16, 32)),
                                (model.conv2[1].running_mean, (32, 64)), (model.conv2[1].running_var, (64, 96))):
                buf.data = run[a:b]
            self.running = run
            self.alpha = None
        else:
            self.lr, self.wd, self.decoupled, self.max_norm, self.gamma = (3e-4 if lr is None else lr), (1e-4 if weight_decay is None else weight_decay), True, 2.0, 2.0
            self.running = None
            self.alpha = None if alpha is None else alpha.to(device=dev, dtype=torch.float32).contiguous()
        self.drop_p = float(model.drop_p)
        self.last_norm = None
        broadcast_parameters(self.flat, self.group)

    def step(self, xb: torch.Tensor, yb: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step on this rank's batch shard; returns the loss (device float[1])."""
        _need_cuda(xb, "xb")
        xb = xb.to(torch.float32).contiguous()
        B = xb.shape[0]
        if mask is None and self.model.training:
            mask = draw_dropout_mask(B, self.drop_p, xb.device)
        logits = self.handle.forward(self.flat, xb, self.running, 0.1, mask, self.drop_p if mask is not None else 0.0)
        loss, d_logits = cnn_loss_grad(logits, yb, self.alpha, self.gamma)
        self.handle.backward(self.flat, d_logits, self.grads)
        scale = reduce_gradients(self.grads, self.group)
        self.steps += 1
        self.last_norm = adam_step(self.flat, self.grads, self.exp_avg, self.exp_avg_sq, self.steps, self.lr, weight_decay=self.wd,
                                   max_norm=self.max_norm, grad_scale=scale, decoupled=self.decoupled)
        if self.arch == CNN_4DOF:
            for bn in (self.model.conv1[1], self.model.conv2[1]):
                bn.num_batches_tracked += 1
        self.model._sig = None                      # the inference handle must re-pack the updated weights
        self.last_logits = logits
        return loss

    def close(self):
        self.handle.close()
