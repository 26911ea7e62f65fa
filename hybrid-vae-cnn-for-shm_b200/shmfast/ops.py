"""Host-side operators over libshmfast's C ABI.  torch is used only for device memory and streams.

Every function takes CUDA tensors and raises on CPU tensors: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ENGINE_AUTO, ENGINE_FP32, ENGINE_TC_BF16X3, ShmfastError, check  # noqa: F401

SHM_MAX_D = _lib.SHM_MAX_D


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ShmfastError(f"{name} must be a CUDA tensor: libshmfast has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _need_cuda(t, name)
    if t.dtype != torch.float32:
        raise ShmfastError(f"{name} must be float32")
    return t.contiguous()


class WindowSource:
    """Describes how windows are read (shm_window_src): a series [rows, d_all] with (T, stride), or
    materialised windows [N, T, D_all]; optional channel selection and the reference's normalisation
    (4DOF normalize_windows, 06_test_full_pipeline.py:124-126; openLAB standardize,
    10_test_hybrid_pipeline.py:233-237).  `std` must already carry the stage's zero-guard."""

    def __init__(self, data: torch.Tensor, T: int, stride: Optional[int] = None, chan: Optional[Sequence[int]] = None,
                 mean=None, std=None, clip: float = 0.0, nan_to_zero: bool = False):
        data = _f32c(data, "window data")
        self.data = data
        if data.dim() == 2:                      # series
            if stride is None or stride <= 0:
                raise ShmfastError("a series source needs a positive stride")
            rows, d_all = data.shape
            self.n_windows = 0 if rows < T else (rows - T) // stride + 1
            win_stride, row_stride = stride * d_all, d_all
        elif data.dim() == 3:                    # materialised windows
            if data.shape[1] != T:
                raise ShmfastError(f"windows have T={data.shape[1]}, expected {T}")
            self.n_windows, _, d_all = data.shape
            win_stride, row_stride = T * d_all, d_all
        else:
            raise ShmfastError("window data must be [rows, D] or [N, T, D]")
        chan = list(range(d_all)) if chan is None else [int(c) for c in chan]
        if not 1 <= len(chan) <= SHM_MAX_D or any(c < 0 or c >= d_all for c in chan):
            raise ShmfastError("bad channel selection")
        self.T, self.D = int(T), len(chan)
        s = _lib.WindowSrc()
        s.base = data.data_ptr()
        s.win_stride, s.row_stride, s.T, s.D = win_stride, row_stride, self.T, self.D
        for i, c in enumerate(chan):
            s.chan[i] = c
        s.normalize = 0
        if mean is not None or std is not None:
            m = np.zeros(self.D, np.float32) if mean is None else np.asarray(mean, dtype=np.float32).reshape(-1)
            sd = np.ones(self.D, np.float32) if std is None else np.asarray(std, dtype=np.float32).reshape(-1)
            if m.size != self.D or sd.size != self.D:
                raise ShmfastError("mean/std must have one entry per selected channel")
            s.normalize = 1
            for i in range(self.D):
                s.mean[i] = float(m[i])
                s.std[i] = float(sd[i])
        s.clip = float(clip)
        s.nan_to_zero = 1 if nan_to_zero else 0
        self.struct = s


def window_normalize(src: WindowSource, idx: Optional[torch.Tensor] = None, n: Optional[int] = None,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Materialise [n, T, D] fp32 windows (make_windows + normalize_windows of the reference)."""
    lib = _lib.load()
    if idx is not None:
        _need_cuda(idx, "idx")
        if idx.dtype != torch.int32:
            raise ShmfastError("idx must be int32")
        n = idx.numel() if n is None else n
    n = src.n_windows if n is None else int(n)
    if out is None:
        out = torch.empty((n, src.T, src.D), dtype=torch.float32, device=src.data.device)
    if n == 0:
        return out
    with torch.cuda.device(src.data.device):
        check(lib.shm_window_normalize(C.byref(src.struct), _ptr(idx), n, _ptr(out), _stream()), "shm_window_normalize")
    return out


VAE_KEYS_PER_LAYER = ("weight_ih", "weight_hh", "bias_ih", "bias_hh")


class VaeScorer:
    """Handle around shm_vae: repacked weights + fused forward/score kernel."""

    def __init__(self, state: dict, device: torch.device, engine: int = ENGINE_AUTO, ln_eps: float = 1e-5):
        lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ShmfastError("VaeScorer needs a CUDA device: libshmfast has no CPU fallback")
        w_ih0 = state["encoder_lstm.weight_ih_l0"]
        self.H = w_ih0.shape[0] // 4
        self.D = w_ih0.shape[1]
        self.Z = state["fc_mu.weight"].shape[0]
        self.L = sum(1 for k in state if k.startswith("encoder_lstm.weight_ih_l"))
        self.has_ln = "layer_norm.weight" in state
        cfg = _lib.VaeCfg(self.D, self.H, self.Z, self.L, 1 if self.has_ln else 0, ln_eps, engine)
        self._keep = None
        w = self._weights_struct(state)
        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(lib.shm_vae_create(C.byref(h), C.byref(cfg), C.byref(w), dev_index), "shm_vae_create")
        self._h = h
        self._lib = lib
        self.engine = lib.shm_vae_engine(h)

    def _weights_struct(self, state: dict) -> _lib.VaeWeights:
        keep = []

        def p(key):
            t = state[key]
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
            t = t.detach().to(dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.VaeWeights()
        for l in range(self.L):
            w.enc_w_ih[l] = p(f"encoder_lstm.weight_ih_l{l}"); w.enc_w_hh[l] = p(f"encoder_lstm.weight_hh_l{l}")
            w.enc_b_ih[l] = p(f"encoder_lstm.bias_ih_l{l}"); w.enc_b_hh[l] = p(f"encoder_lstm.bias_hh_l{l}")
            w.dec_w_ih[l] = p(f"decoder_lstm.weight_ih_l{l}"); w.dec_w_hh[l] = p(f"decoder_lstm.weight_hh_l{l}")
            w.dec_b_ih[l] = p(f"decoder_lstm.bias_ih_l{l}"); w.dec_b_hh[l] = p(f"decoder_lstm.bias_hh_l{l}")
        if self.has_ln:
            w.ln_w = p("layer_norm.weight"); w.ln_b = p("layer_norm.bias")
        w.fc_mu_w = p("fc_mu.weight"); w.fc_mu_b = p("fc_mu.bias")
        w.fc_lv_w = p("fc_logvar.weight"); w.fc_lv_b = p("fc_logvar.bias")
        w.l2h_w = p("fc_latent_to_hidden.weight"); w.l2h_b = p("fc_latent_to_hidden.bias")
        w.out_w = p("output_layer.weight"); w.out_b = p("output_layer.bias")
        self._keep = keep          # keep sources alive until the (stream-ordered) copies are done
        return w

    def update_weights(self, state: dict) -> None:
        w = self._weights_struct(state)
        with torch.cuda.device(self.device):
            check(self._lib.shm_vae_update_weights(self._h, C.byref(w), _stream()), "shm_vae_update_weights")
            if any(not t.is_cuda for t in self._keep):
                torch.cuda.current_stream().synchronize()

    def score(self, src: WindowSource, eps: Optional[torch.Tensor], n: Optional[int] = None,
              idx: Optional[torch.Tensor] = None, n_dev: Optional[torch.Tensor] = None, want_score: bool = True,
              want_latent: bool = False, want_recon: bool = False, want_cnn_in: bool = False, out: Optional[dict] = None):
        """Fused forward.  Returns dict(score, mu, logvar, recon, cnn_in) with the requested tensors."""
        if src.D != self.D:
            raise ShmfastError(f"window source has D={src.D}, model expects {self.D}")
        if idx is not None:
            _need_cuda(idx, "idx")
            if idx.dtype != torch.int32:
                raise ShmfastError("idx must be int32")
            n = idx.numel() if n is None else n
        n = src.n_windows if n is None else int(n)
        dev = src.data.device
        if eps is not None:
            eps = _f32c(eps, "eps")
            if eps.numel() < n * self.Z:
                raise ShmfastError("eps must hold [n, Z] values")
        out = {} if out is None else out

        def buf(name, want, shape):
            if not want:
                return None
            t = out.get(name)
            if t is None:
                t = torch.empty(shape, dtype=torch.float32, device=dev)
                out[name] = t
            return t

        score = buf("score", want_score, (n,))
        mu = buf("mu", want_latent, (n, self.Z))
        lv = buf("logvar", want_latent, (n, self.Z))
        recon = buf("recon", want_recon, (n, src.T, self.D))
        cnn_in = buf("cnn_in", want_cnn_in, (n, 2, src.T, self.D))
        if n == 0:
            return out
        with torch.cuda.device(dev):
            check(self._lib.shm_vae_score(self._h, C.byref(src.struct), _ptr(idx), _ptr(n_dev), _ptr(eps), n, _ptr(score),
                                          _ptr(mu), _ptr(lv), _ptr(recon), _ptr(cnn_in), _stream()), "shm_vae_score")
        return out

    def rescore(self, src: WindowSource, mu_all: torch.Tensor, logvar_all: torch.Tensor, eps: Optional[torch.Tensor],
                idx: Optional[torch.Tensor] = None, n: Optional[int] = None, n_dev: Optional[torch.Tensor] = None,
                want_score: bool = False, want_recon: bool = False, want_cnn_in: bool = True, out: Optional[dict] = None):
        """Second pass on (flagged) windows with fresh noise, reusing the first pass's encoder outputs `mu_all` / `logvar_all`
        ([all windows, Z], from score(..., want_latent=True)): the deterministic encoder is skipped (shm_vae_rescore).
        Returns None when the engine has no re-score path (call score instead)."""
        if src.D != self.D:
            raise ShmfastError(f"window source has D={src.D}, model expects {self.D}")
        mu_all, logvar_all = _f32c(mu_all, "mu_all"), _f32c(logvar_all, "logvar_all")
        if mu_all.shape != logvar_all.shape or mu_all.dim() != 2 or mu_all.shape[1] != self.Z or mu_all.shape[0] < src.n_windows:
            raise ShmfastError("mu_all / logvar_all must be [n_windows, Z]")
        if idx is not None:
            _need_cuda(idx, "idx")
            if idx.dtype != torch.int32:
                raise ShmfastError("idx must be int32")
            n = idx.numel() if n is None else n
        n = src.n_windows if n is None else int(n)
        dev = src.data.device
        if eps is not None:
            eps = _f32c(eps, "eps")
            if eps.numel() < n * self.Z:
                raise ShmfastError("eps must hold [n, Z] values")
        out = {} if out is None else out

        def buf(name, want, shape):
            if not want:
                return None
            t = out.get(name)
            if t is None:
                t = torch.empty(shape, dtype=torch.float32, device=dev)
                out[name] = t
            return t

        score = buf("score", want_score, (n,))
        recon = buf("recon", want_recon, (n, src.T, self.D))
        cnn_in = buf("cnn_in", want_cnn_in, (n, 2, src.T, self.D))
        if n == 0:
            return out
        with torch.cuda.device(dev):
            rc = self._lib.shm_vae_rescore(self._h, C.byref(src.struct), _ptr(idx), _ptr(n_dev), _ptr(mu_all), _ptr(logvar_all), _ptr(eps),
                                           n, _ptr(score), _ptr(recon), _ptr(cnn_in), _stream())
        if rc == -2:                                  # SHM_ERR_UNSUPPORTED: this engine / model has no re-score path
            return None
        check(rc, "shm_vae_rescore")
        return out

    def debug_counters(self, n_cta: int = 148):
        """Tensor-core engine profiling counters [n_cta, 3 roles, 8] (first call enables them); all zero unless the library was
        built with SHMFAST_PROF=1 (the counters cost registers and issue slots in the hot loops, so they are compiled out by default)."""
        buf = np.zeros((n_cta, 3, 8), dtype=np.int64)     # roles: MMA issuer, window-staging warp, epilogue warp 0
        check(self._lib.shm_vae_debug_counters(self._h, buf.ctypes.data_as(C.c_void_p), buf.size), "shm_vae_debug_counters")
        return buf

    def decode(self, z: torch.Tensor, T: int) -> torch.Tensor:
        """TemporalVAE.decode (temporal_vae.py:65-70): z [n,Z] -> recon [n,T,D]."""
        z = _f32c(z, "z")
        if z.dim() != 2 or z.shape[1] != self.Z:
            raise ShmfastError(f"z must be [n, {self.Z}]")
        recon = torch.empty((z.shape[0], int(T), self.D), dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            check(self._lib.shm_vae_decode(self._h, _ptr(z), z.shape[0], int(T), _ptr(recon), _stream()), "shm_vae_decode")
        return recon

    def close(self):
        if getattr(self, "_h", None):
            self._lib.shm_vae_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_WS_CACHE: dict = {}


def _workspace(dev: torch.device, nbytes: int, tag: str) -> torch.Tensor:
    """Per-(device, stream, purpose) scratch, grown on demand and reused: no allocation on the per-chunk path."""
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream, tag)
    t = _WS_CACHE.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty((max(int(nbytes), 256),), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = t
    return t


def compact(score: torch.Tensor, thr: float, want_flag: bool = True):
    """flag = score > thr (strict fp32), idx = np.where(flag)[0] ascending, count (device int32[1]).
    06_test_full_pipeline.py:350-351."""
    lib = _lib.load()
    score = _f32c(score, "score")
    N = score.numel()
    dev = score.device
    flag = torch.empty((N,), dtype=torch.uint8, device=dev) if want_flag else None
    idx = torch.empty((max(N, 1),), dtype=torch.int32, device=dev)
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    ws = _workspace(dev, int(lib.shm_compact_workspace_bytes(N)), "compact")
    with torch.cuda.device(dev):
        check(lib.shm_compact(_ptr(score), float(np.float32(thr)), N, _ptr(flag), _ptr(idx), _ptr(count), _ptr(ws), _stream()),
              "shm_compact")
    return flag, idx, count


class Cnn4dof:
    """Handle around shm_cnn4dof (4DOF/Scripts/Models/cnn_model.py, eval mode)."""

    def __init__(self, state: dict, device: torch.device, bn_eps: float = 1e-5):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ShmfastError("Cnn4dof needs a CUDA device: libshmfast has no CPU fallback")
        self.bn_eps = bn_eps
        w = self._weights_struct(state)
        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self._lib.shm_cnn4dof_create(C.byref(h), C.byref(w), dev_index), "shm_cnn4dof_create")
        self._h = h

    def _weights_struct(self, state):
        keep = []

        def p(key):
            t = state[key]
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
            t = t.detach().to(dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.Cnn4dofWeights()
        for b, blk in enumerate(("conv1", "conv2")):
            w.conv_w[b] = p(f"{blk}.0.weight"); w.conv_b[b] = p(f"{blk}.0.bias")
            w.bn_w[b] = p(f"{blk}.1.weight"); w.bn_b[b] = p(f"{blk}.1.bias")
            w.bn_mean[b] = p(f"{blk}.1.running_mean"); w.bn_var[b] = p(f"{blk}.1.running_var")
        w.fc1_w = p("fc1.0.weight"); w.fc1_b = p("fc1.0.bias"); w.fc2_w = p("fc2.weight"); w.fc2_b = p("fc2.bias")
        w.bn_eps = self.bn_eps
        self._keep = keep
        return w

    def update_weights(self, state: dict) -> None:
        w = self._weights_struct(state)
        with torch.cuda.device(self.device):
            check(self._lib.shm_cnn4dof_update_weights(self._h, C.byref(w), _stream()), "shm_cnn4dof_update_weights")
            if any(not t.is_cuda for t in self._keep):
                torch.cuda.current_stream().synchronize()

    def forward(self, x: torch.Tensor, n: Optional[int] = None, n_dev: Optional[torch.Tensor] = None,
                want_labels: bool = False):
        x = _f32c(x, "x")
        if x.dim() != 4 or tuple(x.shape[1:]) != (2, 100, 12):
            raise ShmfastError(f"4DOF CNN expects [n,2,100,12], got {tuple(x.shape)}")
        n = x.shape[0] if n is None else int(n)
        logits = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        label = torch.empty((n,), dtype=torch.int64, device=x.device) if want_labels else None
        p_struct = torch.empty((n,), dtype=torch.float32, device=x.device) if want_labels else None
        with torch.cuda.device(x.device):
            check(self._lib.shm_cnn4dof_forward(self._h, _ptr(x), _ptr(n_dev), n, _ptr(logits), _ptr(label), _ptr(p_struct),
                                                _stream()), "shm_cnn4dof_forward")
        return (logits, label, p_struct) if want_labels else logits

    def close(self):
        if getattr(self, "_h", None):
            self._lib.shm_cnn4dof_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class CnnOpenLab:
    """Handle around shm_cnnol (openLAB Codes/Models/cnn_model.py, eval mode)."""

    BLOCKS = (0, 2, 4, 6)

    def __init__(self, state: dict, device: torch.device, gn_eps: float = 1e-5, engine: int = ENGINE_AUTO):
        self._lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ShmfastError("CnnOpenLab needs a CUDA device: libshmfast has no CPU fallback")
        self.gn_eps = gn_eps
        w = self._weights_struct(state)
        h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        check(self._lib.shm_cnnol_create(C.byref(h), C.byref(w), dev_index), "shm_cnnol_create")
        self._h = h
        check(self._lib.shm_cnnol_set_engine(h, int(engine)), "shm_cnnol_set_engine")
        self.engine = self._lib.shm_cnnol_engine(h)       # ENGINE_TC_BF16X3 (default) or ENGINE_FP32

    def _weights_struct(self, state):
        keep = []

        def p(key):
            t = state[key]
            if isinstance(t, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
            t = t.detach().to(dtype=torch.float32).contiguous()
            keep.append(t)
            return t.data_ptr()

        w = _lib.CnnOlWeights()
        for b, i in enumerate(self.BLOCKS):
            w.conv_w[b] = p(f"features.{i}.0.weight"); w.conv_b[b] = p(f"features.{i}.0.bias")
            w.gn_w[b] = p(f"features.{i}.1.weight"); w.gn_b[b] = p(f"features.{i}.1.bias")
        w.fc1_w = p("classifier.1.weight"); w.fc1_b = p("classifier.1.bias")
        w.fc2_w = p("classifier.4.weight"); w.fc2_b = p("classifier.4.bias")
        w.gn_eps = self.gn_eps
        self._keep = keep
        return w

    def update_weights(self, state: dict) -> None:
        w = self._weights_struct(state)
        with torch.cuda.device(self.device):
            check(self._lib.shm_cnnol_update_weights(self._h, C.byref(w), _stream()), "shm_cnnol_update_weights")
            if any(not t.is_cuda for t in self._keep):
                torch.cuda.current_stream().synchronize()

    def forward(self, src: WindowSource, n: Optional[int] = None, idx: Optional[torch.Tensor] = None,
                n_dev: Optional[torch.Tensor] = None, want_prob: bool = False):
        if src.T != 200 or src.D != 4:
            raise ShmfastError("openLAB CNN expects windows of T=200, D=4")
        if idx is not None:
            _need_cuda(idx, "idx")
            if idx.dtype != torch.int32:
                raise ShmfastError("idx must be int32")
            n = idx.numel() if n is None else n
        n = src.n_windows if n is None else int(n)
        dev = src.data.device
        logits = torch.empty((n, 2), dtype=torch.float32, device=dev)
        prob = torch.empty((n,), dtype=torch.float64, device=dev) if want_prob else None
        with torch.cuda.device(dev):
            check(self._lib.shm_cnnol_forward(self._h, C.byref(src.struct), _ptr(idx), _ptr(n_dev), n, _ptr(logits), _ptr(prob),
                                              _stream()), "shm_cnnol_forward")
        return (logits, prob) if want_prob else logits

    def close(self):
        if getattr(self, "_h", None):
            self._lib.shm_cnnol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def stitch_segment_rmse(recon: torch.Tensor, full_len: int, stride: int, mean, std, y_true: Optional[torch.Tensor],
                        segment_len: int, want_series: bool = True):
    """1_DOF stitch_windows -> destandardize -> segment_rmse (datasets.py:21-22,38-71), fp64."""
    lib = _lib.load()
    recon = _f32c(recon, "recon")
    N, T, F = recon.shape
    dev = recon.device
    mean_d = torch.as_tensor(np.asarray(mean, dtype=np.float64), device=dev)
    std_d = torch.as_tensor(np.asarray(std, dtype=np.float64), device=dev)
    series = torch.empty((full_len, F), dtype=torch.float64, device=dev) if want_series else None
    rm = None
    if y_true is not None:
        y_true = _f32c(y_true, "y_true")
        rm = torch.empty(((full_len + segment_len - 1) // segment_len,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.shm_stitch_segment_rmse(_ptr(recon), N, T, F, stride, full_len, _ptr(mean_d), _ptr(std_d), _ptr(y_true),
                                          segment_len, _ptr(series), _ptr(rm), _stream()), "shm_stitch_segment_rmse")
    return series, rm


def percentile(scores: torch.Tensor, q: float) -> torch.Tensor:
    """np.percentile(scores, q) (linear interpolation, fp32 arithmetic like NumPy) -> device fp64[1]."""
    lib = _lib.load()
    scores = _f32c(scores, "scores")
    dev = scores.device
    res = torch.empty((1,), dtype=torch.float64, device=dev)
    ws = _workspace(dev, int(lib.shm_percentile_workspace_bytes(scores.numel())), "percentile")
    with torch.cuda.device(dev):
        check(lib.shm_percentile(_ptr(scores), scores.numel(), float(q), _ptr(res), _ptr(ws), _stream()), "shm_percentile")
    return res


def scatter_flagged_4dof(idx: torch.Tensor, count: Optional[torch.Tensor], cap: int, label: torch.Tensor, p_struct: torch.Tensor,
                         n: int):
    """y_pred[idx[j]] = label[j]; hyb_score_full[idx[j]] = p_struct[j] for j < min(count, cap); 0 elsewhere
    (06_test_full_pipeline.py:336,356,368-372).  One kernel + two memsets; `count` stays on the device."""
    lib = _lib.load()
    _need_cuda(idx, "idx")
    dev = idx.device
    y_pred = torch.empty((n,), dtype=torch.int64, device=dev)
    p_full = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.shm_scatter_flagged_4dof(_ptr(idx), _ptr(count), int(cap), _ptr(label), _ptr(p_struct), int(n), _ptr(y_pred),
                                           _ptr(p_full), _stream()), "shm_scatter_flagged_4dof")
    return y_pred, p_full


def scatter_flagged_openlab(idx: torch.Tensor, count: Optional[torch.Tensor], cap: int, prob: torch.Tensor, cnn_thr: float, n: int):
    """pred_bin = prob_st >= thr (fp64, 10_test_hybrid_pipeline.py:300-301); dense y_pred (0 not flagged / 1 sensor fault /
    2 structural, :389-401) and prob_full."""
    lib = _lib.load()
    _need_cuda(idx, "idx")
    dev = idx.device
    pred_bin = torch.empty((cap,), dtype=torch.int64, device=dev)
    y_pred = torch.empty((n,), dtype=torch.int64, device=dev)
    prob_full = torch.empty((n,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        check(lib.shm_scatter_flagged_openlab(_ptr(idx), _ptr(count), int(cap), _ptr(prob), float(cnn_thr), int(n), _ptr(pred_bin),
                                              _ptr(y_pred), _ptr(prob_full), _stream()), "shm_scatter_flagged_openlab")
    return pred_bin, y_pred, prob_full


def hybrid4dof_score(vae: VaeScorer, cnn: Cnn4dof, src: WindowSource, eps1: Optional[torch.Tensor], eps2: Optional[torch.Tensor],
                     thr: float, n: Optional[int] = None, max_flagged: Optional[int] = None, want_flag: bool = True,
                     want_compact: bool = True, want_dense: bool = True, out: Optional[dict] = None) -> dict:
    """eval_group of 06_test_full_pipeline.py:327-383 as ONE C call (shm_hybrid4dof_score): no host synchronisation.
    Returns score, flag, idx, status (device int32[2]: flagged count, overflow), logits/label/p_struct [max_flagged]
    (want_compact) and the dense y_pred / p_full [n] (want_dense)."""
    lib = _lib.load()
    n = src.n_windows if n is None else int(n)
    cap = n if max_flagged is None else min(n, int(max_flagged))
    dev = src.data.device
    Z = vae.Z
    for t, name, need in ((eps1, "eps1", n * Z), (eps2, "eps2", cap * Z)):
        if t is not None:
            _f32c(t, name)
            if t.numel() < need:
                raise ShmfastError(f"{name} must hold {need} values")
    out = {} if out is None else out

    def buf(name, want, shape, dt):
        if not want:
            return None
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dt, device=dev)
            out[name] = t
        return t

    score = buf("score", True, (n,), torch.float32)
    flag = buf("flag", want_flag, (n,), torch.uint8)
    idx = buf("idx", True, (max(n, 1),), torch.int32)
    status = buf("status", True, (2,), torch.int32)
    logits = buf("logits", want_compact, (cap, 2), torch.float32)
    label = buf("label", want_compact, (cap,), torch.int64)
    p_struct = buf("p_struct", want_compact, (cap,), torch.float32)
    y_pred = buf("y_pred", want_dense, (n,), torch.int64)
    p_full = buf("p_full", want_dense, (n,), torch.float32)
    out["count"] = status[:1]
    out["n_flagged"] = cap
    with torch.cuda.device(dev):
        nbytes = int(lib.shm_hybrid4dof_workspace_bytes(vae._h, n, cap))
        if nbytes < 0:
            check(nbytes, "shm_hybrid4dof_workspace_bytes")
        ws = _workspace(dev, nbytes, "hybrid4dof")
        check(lib.shm_hybrid4dof_score(vae._h, cnn._h, C.byref(src.struct), n, _ptr(eps1), _ptr(eps2), float(np.float32(thr)), cap,
                                       _ptr(score), _ptr(flag), _ptr(idx), _ptr(status), _ptr(logits), _ptr(label), _ptr(p_struct),
                                       _ptr(y_pred), _ptr(p_full), _ptr(ws), ws.numel(), _stream()), "shm_hybrid4dof_score")
    return out


def hybridol_score(vae: VaeScorer, cnn: CnnOpenLab, src_gate: WindowSource, src_raw: WindowSource, eps: Optional[torch.Tensor],
                   vae_thr: float, cnn_thr: float, n: Optional[int] = None, max_flagged: Optional[int] = None,
                   want_flag: bool = True, want_compact: bool = True, want_dense: bool = True, out: Optional[dict] = None) -> dict:
    """10_test_hybrid_pipeline.py:351-367 + stage2_predict_cnn (:265-302) + label scatter (:389-401) as ONE C call."""
    lib = _lib.load()
    n = src_gate.n_windows if n is None else int(n)
    cap = n if max_flagged is None else min(n, int(max_flagged))
    dev = src_gate.data.device
    if src_raw.T != 200 or src_raw.D != 4:
        raise ShmfastError("openLAB CNN expects windows of T=200, D=4")
    if eps is not None:
        _f32c(eps, "eps")
        if eps.numel() < n * vae.Z:
            raise ShmfastError("eps must hold [n, Z] values")
    out = {} if out is None else out

    def buf(name, want, shape, dt):
        if not want:
            return None
        t = out.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dt, device=dev)
            out[name] = t
        return t

    score = buf("score", True, (n,), torch.float32)
    flag = buf("flag", want_flag, (n,), torch.uint8)
    idx = buf("idx", True, (max(n, 1),), torch.int32)
    status = buf("status", True, (2,), torch.int32)
    logits = buf("logits", want_compact, (cap, 2), torch.float32)
    prob = buf("prob", want_compact, (cap,), torch.float64)
    pred = buf("pred", want_compact, (cap,), torch.int64)
    y_pred = buf("y_pred", want_dense, (n,), torch.int64)
    prob_full = buf("prob_full", want_dense, (n,), torch.float64)
    out["count"] = status[:1]
    out["n_flagged"] = cap
    with torch.cuda.device(dev):
        ws = _workspace(dev, int(lib.shm_hybridol_workspace_bytes(n, cap)), "hybridol")
        check(lib.shm_hybridol_score(vae._h, cnn._h, C.byref(src_gate.struct), C.byref(src_raw.struct), n, _ptr(eps),
                                     float(np.float32(vae_thr)), float(cnn_thr), cap, _ptr(score), _ptr(flag), _ptr(idx), _ptr(status),
                                     _ptr(logits), _ptr(prob), _ptr(pred), _ptr(y_pred), _ptr(prob_full), _ptr(ws), ws.numel(),
                                     _stream()), "shm_hybridol_score")
    return out


def gemm_f32(A: torch.Tensor, a_ms: int, a_ks: int, B: torch.Tensor, b_ks: int, b_ns: int, M: int, N: int, K: int,
             bias: Optional[torch.Tensor] = None, splitk: bool = False, mode: int = _lib.GEMM_SIMT) -> torch.Tensor:
    """C[M,N] = sum_k A[m*a_ms + k*a_ks] * B[k*b_ks + n*b_ns] (+ bias[n]) through shm_gemm_f32: the contraction the training
    steps run (mode: _lib.GEMM_SIMT / GEMM_TC_F16X3 / GEMM_TC_BF16X3).  A / B are flat float32 CUDA buffers."""
    lib = _lib.load()
    A, B = _f32c(A, "A"), _f32c(B, "B")
    C_ = torch.zeros((M, N), dtype=torch.float32, device=A.device)
    with torch.cuda.device(A.device):
        check(lib.shm_gemm_f32(_ptr(A), int(a_ms), int(a_ks), _ptr(B), int(b_ks), int(b_ns), _ptr(C_), N, int(M), int(N), int(K),
                               _ptr(bias), 1 if splitk else 0, int(mode), _stream()), "shm_gemm_f32")
    return C_
