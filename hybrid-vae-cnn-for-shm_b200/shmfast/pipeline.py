"""Script-level hot loops of the reference, device resident.

* Hybrid4dof.run  == eval_group of 4DOF/Scripts/06_test_full_pipeline.py:327-383
* HybridOpenLab.run == 10_test_hybrid_pipeline.py:351-367 + stage2_predict_cnn (:265-302)
* score_windows == full_mse_scores_batched (04_vae_thresholding.py:113-124) / recon_mse_per_window

Everything between the input tensors and the result tensors runs as libshmfast kernels on the
current stream.  Two forms per stage:

* `run(..., sync_count=True)`: the reference's control flow -- the 4-byte flagged count is read back once to size
  the flagged-subset tensors exactly like np.where does;
* `run_dense(...)`: ONE C call (shm_hybrid4dof_score / shm_hybridol_score), no host round trip; per-flagged outputs
  have `max_flagged` rows and `status[1]` reports an overflow (count > max_flagged) instead of dropping windows silently.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import ops
from .ops import WindowSource


def guard_std_4dof(std) -> np.ndarray:
    """load_stats of 4DOF/Scripts/06_test_full_pipeline.py:113-121: std[std == 0] = 1e-6."""
    s = np.array(std, dtype=np.float32, copy=True)
    s[s == 0] = 1e-6
    return s


def score_windows(vae: ops.VaeScorer, src: WindowSource, eps: Optional[torch.Tensor], n: Optional[int] = None) -> torch.Tensor:
    """Per-window reconstruction MSE for every window of `src` (one fused kernel, no batching loop)."""
    return vae.score(src, eps, n=n)["score"]


class Hybrid4dof:
    def __init__(self, vae: ops.VaeScorer, cnn: ops.Cnn4dof, thr: float):
        self.vae, self.cnn, self.thr = vae, cnn, float(thr)

    def run(self, src: WindowSource, eps1: Optional[torch.Tensor], eps2: Optional[torch.Tensor], n: Optional[int] = None,
            sync_count: bool = True, max_flagged: Optional[int] = None) -> dict:
        """score pass -> strict threshold + ascending compaction -> SECOND VAE pass on the flagged
        windows with fresh noise (eps2[j] belongs to the j-th flagged window, the order in which the
        reference draws them) -> residual stack -> CNN -> label = argmax+1, p_struct."""
        n = src.n_windows if n is None else int(n)
        if not sync_count:
            return self.run_dense(src, eps1, eps2, n=n, max_flagged=max_flagged)
        out = self.vae.score(src, eps1, n=n, want_latent=True)     # mu / logvar feed the second pass (the encoder is deterministic)
        score = out["score"]
        flag, idx, count = ops.compact(score, self.thr)
        n_f = int(count.item())
        res = dict(score=score, flag=flag, idx=idx, count=count, n_flagged=n_f)
        if n_f == 0:
            dev = score.device
            res.update(logits=torch.empty((0, 2), device=dev), label=torch.empty((0,), dtype=torch.int64, device=dev),
                       p_struct=torch.empty((0,), device=dev))
            return res
        second = self.vae.rescore(src, out["mu"], out["logvar"], eps2, idx=idx, n=n_f, n_dev=count, want_cnn_in=True)
        if second is None:                                         # engine without a re-score path: full forward, same result
            second = self.vae.score(src, eps2, n=n_f, idx=idx, n_dev=count, want_score=False, want_cnn_in=True)
        logits, label, p_struct = self.cnn.forward(second["cnn_in"], n=n_f, n_dev=count, want_labels=True)
        res.update(logits=logits, label=label, p_struct=p_struct, cnn_in=second["cnn_in"])
        return res

    def run_dense(self, src: WindowSource, eps1: Optional[torch.Tensor], eps2: Optional[torch.Tensor], n: Optional[int] = None,
                  max_flagged: Optional[int] = None, out: Optional[dict] = None) -> dict:
        """The whole loop as one C call; no host synchronisation.  eps2 must hold [max_flagged, Z] (or [n, Z]) rows.  The second
        pass runs in bounded chunks inside the library, so max_flagged=None (= n) costs no extra memory.  `status` =
        device int32[2] {flagged count, overflow}: check `status[1] == 0` when max_flagged < n."""
        res = ops.hybrid4dof_score(self.vae, self.cnn, src, eps1, eps2, self.thr, n=n, max_flagged=max_flagged, out=out)
        return res

    @staticmethod
    def scatter(res: dict, n: int):
        """y_pred / hyb_score_full of 06_test_full_pipeline.py:336,356,368-372 (0 for unflagged windows)."""
        if "y_pred" in res and res["y_pred"].shape[0] == n:
            return res["y_pred"], res["p_full"]
        cap = int(res["label"].shape[0])
        return ops.scatter_flagged_4dof(res["idx"], res["count"], cap, res["label"], res["p_struct"], n)


class HybridOpenLab:
    def __init__(self, vae: ops.VaeScorer, cnn: ops.CnnOpenLab, vae_thr: float, cnn_thr: float):
        self.vae, self.cnn, self.vae_thr, self.cnn_thr = vae, cnn, float(vae_thr), float(cnn_thr)

    def run(self, src_gate: WindowSource, src_raw: WindowSource, eps: Optional[torch.Tensor], n: Optional[int] = None,
            sync_count: bool = True, max_flagged: Optional[int] = None) -> dict:
        n = src_gate.n_windows if n is None else int(n)
        if not sync_count:
            return self.run_dense(src_gate, src_raw, eps, n=n, max_flagged=max_flagged)
        score = self.vae.score(src_gate, eps, n=n)["score"]
        flag, idx, count = ops.compact(score, self.vae_thr)
        n_f = int(count.item())
        res = dict(score=score, flag=flag, idx=idx, count=count, n_flagged=n_f)
        dev = score.device
        if n_f == 0:
            res.update(logits=torch.empty((0, 2), device=dev), prob=torch.empty((0,), dtype=torch.float64, device=dev),
                       pred=torch.empty((0,), dtype=torch.int64, device=dev))
            return res
        logits, prob = self.cnn.forward(src_raw, n=n_f, idx=idx, n_dev=count, want_prob=True)
        pred, y_pred, prob_full = ops.scatter_flagged_openlab(idx, count, n_f, prob, self.cnn_thr, n)
        res.update(logits=logits, prob=prob, pred=pred, y_pred=y_pred, prob_full=prob_full)
        return res

    def run_dense(self, src_gate: WindowSource, src_raw: WindowSource, eps: Optional[torch.Tensor], n: Optional[int] = None,
                  max_flagged: Optional[int] = None, out: Optional[dict] = None) -> dict:
        """One C call (shm_hybridol_score), no host synchronisation; `status` = device int32[2] {flagged count, overflow}."""
        return ops.hybridol_score(self.vae, self.cnn, src_gate, src_raw, eps, self.vae_thr, self.cnn_thr, n=n,
                                  max_flagged=max_flagged, out=out)
