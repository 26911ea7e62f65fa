"""shm_vae_rescore: the hybrid loop's second pass (06_test_full_pipeline.py:360-365, `recon, _, _ = vae(z_sel)` with fresh noise)
reusing the first pass's encoder outputs must equal the full forward bit for bit -- the encoder is deterministic in eval mode."""
import numpy as np
import pytest
import torch

from shmfast import ops, synth
from shmfast.pipeline import Hybrid4dof, guard_std_4dof

pytestmark = pytest.mark.gpu


def _setup(dev, N=700, seed=7):
    T, D, Z = 100, 12, 16
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=seed, scale=2.0), dev)
    mean, std = synth.stats(D, seed=seed)
    series = torch.from_numpy(synth.series(N + T - 1, D, seed=seed)).to(dev)
    src = ops.WindowSource(series, T, stride=1, mean=mean, std=guard_std_4dof(std), nan_to_zero=True)
    g = torch.Generator().manual_seed(seed)
    eps1 = torch.randn((N, Z), generator=g).to(dev)
    eps2 = torch.randn((N, Z), generator=g).to(dev)
    return vae, src, eps1, eps2


def test_rescore_equals_full_second_pass(cuda_dev):
    vae, src, eps1, eps2 = _setup(cuda_dev)
    first = vae.score(src, eps1, want_latent=True)
    thr = float(torch.quantile(first["score"], 0.55))
    flag, idx, count = ops.compact(first["score"], thr)
    k = int(count.item())
    assert 100 < k < src.n_windows
    full = vae.score(src, eps2, idx=idx, n=k, want_score=True, want_recon=True, want_cnn_in=True)
    re = vae.rescore(src, first["mu"], first["logvar"], eps2, idx=idx, n=k, want_score=True, want_recon=True, want_cnn_in=True)
    assert re is not None
    for key in ("score", "recon", "cnn_in"):
        assert torch.equal(full[key], re[key]), key
    # capacity larger than the device-side count (the asynchronous pipeline's form): the first k entries are the same
    cap = min(src.n_windows, k + 200)
    re2 = vae.rescore(src, first["mu"], first["logvar"], eps2, idx=idx, n=cap, n_dev=count, want_cnn_in=True)
    assert torch.equal(re2["cnn_in"][:k], full["cnn_in"])
    # without idx: every window, same as a plain second forward
    full_all = vae.score(src, eps2, want_score=True)
    re_all = vae.rescore(src, first["mu"], first["logvar"], eps2, want_score=True, want_cnn_in=False)
    assert torch.equal(full_all["score"], re_all["score"])


def test_hybrid_run_uses_rescore_and_matches_full_path(cuda_dev):
    vae, src, eps1, eps2 = _setup(cuda_dev, N=500, seed=9)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=9), cuda_dev)
    s = vae.score(src, eps1)["score"]
    hyb = Hybrid4dof(vae, cnn, float(torch.quantile(s, 0.5)))
    res = hyb.run(src, eps1, eps2)
    k = int(res["count"].item())
    ref = vae.score(src, eps2, idx=res["idx"], n=k, want_score=False, want_cnn_in=True)
    assert torch.equal(res["cnn_in"][:k], ref["cnn_in"])
    logits, label, p = cnn.forward(ref["cnn_in"], n=k, want_labels=True)
    assert torch.equal(res["logits"][:k], logits) and torch.equal(res["label"][:k], label)


def test_rescore_reports_unsupported_for_the_two_tile_model(cuda_dev):
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=0), cuda_dev)
    series = torch.from_numpy(synth.series(199 + 20 * 40, 3, seed=1)).to(cuda_dev)
    src = ops.WindowSource(series, 200, stride=20)
    first = vae.score(src, None, want_latent=True)
    assert vae.rescore(src, first["mu"], first["logvar"], None, want_score=True, want_cnn_in=False) is None
    with pytest.raises(ops.ShmfastError):
        vae.rescore(src, first["mu"][:, :4], first["logvar"][:, :4], None)
