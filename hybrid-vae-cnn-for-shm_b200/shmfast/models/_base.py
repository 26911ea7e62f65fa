"""nn.Module shims with the reference's class surface whose forward runs on
libshmfast.  They hold real nn.Parameters inside the same torch containers the reference builds, in
the same construction order, so state_dict keys, shapes, default initialisation under a seed,
.parameters(), .to(), .load_state_dict() all behave identically; only the arithmetic moves to the
hand-written kernels.  CPU tensors raise: there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ShmfastError, WindowSource


def _signature(module: nn.Module):
    sig = []
    for t in list(module.parameters()) + list(module.buffers()):
        sig.append((t.data_ptr(), t._version, t.device))
    return tuple(sig)


class TemporalVAEBase(nn.Module):
    """LSTM-VAE: encode / reparameterize / decode / forward as in
    4DOF/Scripts/Models/temporal_vae.py:14-77 (layer_norm=True),
    20250506_openLAB_tests/Codes/Models/temporal_vae_model.py:4-66 (layer_norm=True),
    1_DOF/Scripts/Models/temporal_vae.py:8-58 (layer_norm=False)."""

    _layer_norm = True
    engine = ops.ENGINE_AUTO

    def __init__(self, input_dim: int, latent_dim: int, hidden_dim: int, num_layers: int, dropout: float) -> None:
        super().__init__()
        self.input_dim = input_dim
        self.latent_dim = latent_dim
        self.hidden_dim = hidden_dim
        self.num_layers = num_layers
        drop = dropout if num_layers > 1 else 0.0
        # parameter containers, constructed in the reference's order (same RNG consumption at init)
        self.encoder_lstm = nn.LSTM(input_size=input_dim, hidden_size=hidden_dim, num_layers=num_layers, batch_first=True, dropout=drop)
        if self._layer_norm:
            self.layer_norm = nn.LayerNorm(hidden_dim)
        self.fc_mu = nn.Linear(hidden_dim, latent_dim)
        self.fc_logvar = nn.Linear(hidden_dim, latent_dim)
        self.fc_latent_to_hidden = nn.Linear(latent_dim, hidden_dim)
        self.decoder_lstm = nn.LSTM(input_size=hidden_dim, hidden_size=hidden_dim, num_layers=num_layers, batch_first=True, dropout=drop)
        self.output_layer = nn.Linear(hidden_dim, input_dim)
        self._scorer = None
        self._sig = None
        self._trainer = None

    # -- handle management -------------------------------------------------------------------
    def scorer(self) -> ops.VaeScorer:
        """The libshmfast handle, (re)packed whenever a parameter changed (load_state_dict,
        optimizer.step, .to())."""
        dev = self.fc_mu.weight.device
        if dev.type != "cuda":
            raise ShmfastError("model parameters are on the CPU: move the model to a CUDA device (no CPU fallback)")
        sig = _signature(self)
        if self._scorer is None or self._scorer.device != dev:
            if self._scorer is not None:
                self._scorer.close()
            self._scorer = ops.VaeScorer(self.state_dict(), dev, engine=self.engine,
                                         ln_eps=self.layer_norm.eps if self._layer_norm else 1e-5)
            self._sig = sig
        elif sig != self._sig:
            self._scorer.update_weights(self.state_dict())
            self._sig = sig
        return self._scorer

    def _check_input(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise ShmfastError("input is a CPU tensor: libshmfast has no CPU fallback")
        if x.dim() != 3 or x.shape[2] != self.input_dim:
            raise ShmfastError(f"expected [B, T, {self.input_dim}], got {tuple(x.shape)}")
        return x.detach().to(torch.float32).contiguous()

    def _train_forward(self, x: torch.Tensor):
        """train() mode (03_train_vae.py:262): same kernels as shmfast.train.VaeTrainer, bridged into autograd so the
        reference's loss.backward() / clip_grad_norm_ / opt.step() lines run unchanged.  Inter-layer dropout masks and
        eps are drawn with torch's generator on the device and handed to the kernels."""
        from .. import train as _train
        B, T, _ = x.shape
        h = self._trainer
        if h is None or h.T != T or h.max_batch < B or h.device != x.device:
            if h is not None:
                h.close()
            h = self._trainer = _train.VaeTrainHandle(_train.vae_cfg_of(self), T, B, x.device)
        p = float(self.encoder_lstm.dropout)
        enc, dec = _train.draw_dropout_masks(self.num_layers, B, T, self.hidden_dim, p, x.device)
        eps = torch.randn((B, self.latent_dim), dtype=torch.float32, device=x.device)
        return _train.VaeTrainFunction.apply(h, x, eps, enc, dec, p, *self.parameters())

    # -- reference surface -------------------------------------------------------------------
    def encode(self, x: torch.Tensor):
        x = self._check_input(x)
        out = self.scorer().score(WindowSource(x, x.shape[1]), None, want_score=False, want_latent=True)
        return out["mu"], out["logvar"]

    @staticmethod
    def reparameterize(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
        std = torch.exp(0.5 * logvar)
        eps = torch.randn_like(std)
        return mu + eps * std

    def decode(self, z: torch.Tensor, seq_len: int) -> torch.Tensor:
        if not z.is_cuda:
            raise ShmfastError("z is a CPU tensor: libshmfast has no CPU fallback")
        return self.scorer().decode(z.detach().to(torch.float32).contiguous(), int(seq_len))

    def forward(self, x: torch.Tensor):
        x = self._check_input(x)
        if self.training:
            return self._train_forward(x)
        # the reference draws eps = torch.randn_like(std) inside forward() even in eval mode
        # (temporal_vae.py:60-63); the same call on the same generator keeps the RNG stream identical.
        eps = torch.randn((x.shape[0], self.latent_dim), dtype=torch.float32, device=x.device)
        out = self.scorer().score(WindowSource(x, x.shape[1]), eps, want_score=False, want_latent=True, want_recon=True)
        return out["recon"], out["mu"], out["logvar"]


class _HandleModule(nn.Module):
    """Shared handle refresh logic for the CNN shims."""

    _handle_cls = None

    ARCH = None          # _lib.CNN_4DOF / _lib.CNN_OPENLAB
    drop_p = 0.0

    def _init_handle(self):
        self._handle = None
        self._sig = None
        self._trainer = None

    def _bn_buffers(self):
        return []

    def _train_forward(self, x: torch.Tensor) -> torch.Tensor:
        """train() mode (05_train_cnn.py:272, 06_train_cnn.py:414): forward with batch statistics / dropout on the training kernels,
        bridged into autograd so the reference's `loss.backward()` / `optimizer.step()` lines run unchanged.  The dropout keep-mask
        is drawn with torch's generator on the device; BatchNorm running statistics are updated like nn.BatchNorm2d does."""
        from .. import cnn_train as _ct
        if not x.is_cuda:
            raise ShmfastError("input is a CPU tensor: libshmfast has no CPU fallback")
        x = x.detach().to(torch.float32).contiguous()
        B = x.shape[0]
        h = self._trainer
        if h is None or h.max_batch < B or h.device != x.device:
            if h is not None:
                h.close()
            h = self._trainer = _ct.CnnTrainHandle(self.ARCH, B, x.device)
        mask = _ct.draw_dropout_mask(B, float(self.drop_p), x.device)
        bufs = self._bn_buffers()
        running = torch.cat([b.reshape(-1) for b in bufs]).contiguous() if bufs else None
        logits = _ct.CnnTrainFunction.apply(h, x, running, 0.1, mask, float(self.drop_p) if mask is not None else 0.0, *self.parameters())
        if bufs:
            o = 0
            with torch.no_grad():
                for b in bufs:
                    b.copy_(running[o:o + b.numel()].view_as(b))
                    o += b.numel()
                for m in self.modules():
                    if isinstance(m, nn.BatchNorm2d):
                        m.num_batches_tracked += 1
        return logits

    def handle(self):
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise ShmfastError("model parameters are on the CPU: move the model to a CUDA device (no CPU fallback)")
        sig = _signature(self)
        if self._handle is None or self._handle.device != dev:
            if self._handle is not None:
                self._handle.close()
            self._handle = self._handle_cls(self.state_dict(), dev)
            self._sig = sig
        elif sig != self._sig:
            self._handle.update_weights(self.state_dict())
            self._sig = sig
        return self._handle

    def _eval_only(self, x):
        if not x.is_cuda:
            raise ShmfastError("input is a CPU tensor: libshmfast has no CPU fallback")
        return x.detach().to(torch.float32).contiguous()
