"""openLAB extraction front-end on the device through the C ABI (shm_openlab_extract):
bit-identical to the reference's 01_extract_windows_and_labels.py outputs on the reference's own raw runs
(tests/golden/openlab_frontend.npz), to the oracle on synthetic edge cases, and wired into the hybrid scorer without a
host round trip."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from shmfast import ops, openlab_frontend as FE, synth
from shmfast.pipeline import HybridOpenLab

pytestmark = pytest.mark.gpu

F32 = ("u_min", "u_max", "dms_range", "raw_invalid_ratio", "raw_outlier_ratio", "removed_ratio")
I32 = ("flatline_loadaware", "all_nan_struct")


def _check_against(run: FE.ExtractedRun, ref: dict):
    assert run.n_windows == ref["n_windows"] and run.a_clean.shape[0] == ref["rows_kept"]
    assert np.array_equal(run.a_clean.cpu().numpy(), ref["A_clean"], equal_nan=True)      # every cleaned sample, bit for bit
    assert np.array_equal(run.a_raw.cpu().numpy(), ref["A_raw"], equal_nan=True)
    if run.n_windows == 0:
        return
    assert np.array_equal(run.label.cpu().numpy(), ref["label"])
    for k in F32:
        assert np.array_equal(run.meta[k].cpu().numpy(), ref[k], equal_nan=True), k
    for k in I32:
        assert np.array_equal(run.meta[k].cpu().numpy(), ref[k]), k


def test_frontend_reference_runs_bit_exact(cuda_dev, golden_dir):
    g = np.load(golden_dir / "openlab_frontend.npz")
    offs, wpr = g["run_row_offsets"], g["win_per_run"]
    w0 = 0
    for r in range(len(wpr)):
        raw = g["raw"][offs[r]:offs[r + 1]]
        run = FE.extract_run(torch.from_numpy(raw).to(cuda_dev))
        w1 = w0 + int(wpr[r])
        # against the reference's own outputs ...
        assert run.n_windows == w1 - w0
        assert np.array_equal(run.label.cpu().numpy(), g["label"][w0:w1])
        assert np.array_equal(run.win_start_idx.cpu().numpy(), g["win_start_idx"][w0:w1])
        for k in F32:
            assert np.array_equal(run.meta[k].cpu().numpy(), g[k][w0:w1], equal_nan=True), k
        for k in I32:
            assert np.array_equal(run.meta[k].cpu().numpy(), g[k][w0:w1]), k
        Xc = ops.window_normalize(ops.WindowSource(run.a_clean, 200, stride=20)).cpu().numpy()
        Xr = ops.window_normalize(ops.WindowSource(run.a_raw, 200, stride=20)).cpu().numpy()
        assert np.array_equal(np.nansum(Xc.astype(np.float64), axis=(1, 2)), g["xc_sum"][w0:w1])
        assert np.array_equal(np.nansum(Xr.astype(np.float64), axis=(1, 2)), g["xr_sum"][w0:w1])
        for k, w in enumerate(g["sample_idx"]):
            if w0 <= w < w1:
                assert np.array_equal(Xc[w - w0], g["xc_sample"][k], equal_nan=True)
                assert np.array_equal(Xr[w - w0], g["xr_sample"][k], equal_nan=True)
        # ... and against the oracle, sample by sample
        _check_against(run, O.openlab_extract_run(raw))
        w0 = w1
    assert w0 == 6432


def test_frontend_edge_cases_vs_oracle(cuda_dev):
    rng = np.random.Generator(np.random.PCG64(11))
    base = (30 + np.cumsum(0.05 * rng.standard_normal((5000, 4)), axis=0)).astype(np.float32)
    cases = []
    a = base.copy(); a[700:720, 0] = np.nan; a[2500, 0] = np.inf; cases.append(a)                 # DMS gaps: rows dropped
    a = base.copy(); a[:, 3] = -2e5; cases.append(a)                                               # a channel obstructed throughout
    a = base.copy(); a[1234, 1] = 90.0; a[3000:3003, 2] = np.nan; cases.append(a)                  # AND-rule hit, invalid samples
    a = base.copy(); a[0, 2] = np.nan; cases.append(a)                                             # first sample invalid: all-NaN clean channel
    a = base.copy(); a[:, 2] = 21.0; a[:, 0] += np.linspace(0, 50, 5000, dtype=np.float32); cases.append(a)   # flatline under load
    a = (base + 60).astype(np.float32); cases.append(a)                                            # |u| >= 65 with small steps: no hit
    cases.append(base[:150].copy())                                                                # shorter than one window
    cases.append(base[:200].copy())                                                                # exactly one window
    for raw in cases:
        _check_against(FE.extract_run(torch.from_numpy(raw).to(cuda_dev)), O.openlab_extract_run(raw))
    cfg = FE.default_cfg(); cfg.ma_window = 4
    with pytest.raises(FE.ShmfastError):
        FE.extract_run(torch.from_numpy(base).to(cuda_dev), cfg)
    with pytest.raises(FE.ShmfastError):
        FE.extract_run(torch.from_numpy(base))                                                     # CPU tensor: no CPU fallback


def test_frontend_feeds_the_hybrid_scorer(cuda_dev, golden_dir):
    """Device-resident series from the front-end as strided window views == the same windows materialised on the host."""
    g = np.load(golden_dir / "openlab_frontend.npz")
    offs = g["run_row_offsets"]
    raw = g["raw"][offs[2]:offs[3]]
    run = FE.extract_run(torch.from_numpy(raw).to(cuda_dev))
    ref = O.openlab_extract_run(raw)
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=3, scale=2.0), cuda_dev)
    cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=3), cuda_dev)
    vmu, vsd = synth.stats(3, seed=1)
    cmu, csd = synth.stats(4, seed=2)
    N = run.n_windows
    eps = torch.from_numpy(synth.eps(N, 8, seed=4)).to(cuda_dev)
    src_g, src_r = run.sources([1, 2, 3], vmu, vsd, cmu, csd)
    s1 = vae.score(src_g, eps)["score"]
    thr = float(torch.quantile(s1, 0.6))
    r1 = HybridOpenLab(vae, cnn, thr, 0.5).run(src_g, src_r, eps)
    Xc = torch.from_numpy(np.stack([ref["A_clean"][i:i + 200] for i in ref["win_start_idx"]])).to(cuda_dev)
    Xr = torch.from_numpy(np.stack([ref["A_raw"][i:i + 200] for i in ref["win_start_idx"]])).to(cuda_dev)
    g2 = ops.WindowSource(Xc, 200, chan=[1, 2, 3], mean=vmu, std=vsd, clip=10.0, nan_to_zero=True)
    r2s = ops.WindowSource(Xr, 200, mean=cmu, std=csd, clip=10.0, nan_to_zero=True)
    r2 = HybridOpenLab(vae, cnn, thr, 0.5).run(g2, r2s, eps)
    assert torch.equal(r1["score"], r2["score"]) and torch.equal(r1["flag"], r2["flag"])
    assert torch.equal(r1["logits"], r2["logits"]) and torch.equal(r1["pred"], r2["pred"])
    assert int(r1["count"].item()) > 0
