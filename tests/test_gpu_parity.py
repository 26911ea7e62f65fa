"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors
produced by the reference's own modules.  Tolerances are BASELINE.json's: scores and logits within
1e-4 relative (logits with an absolute floor), flags / labels bit-identical
except windows whose score lies within that tolerance of the threshold (counted and reported)."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from shmfast import ops, synth
from shmfast.pipeline import Hybrid4dof, HybridOpenLab, guard_std_4dof

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4          # north_star tolerance on scores / logits
LOGIT_ABS = 2e-4        # absolute floor for logits (fp32 re-association alone moves them ~3e-5 rel)


def rel_err(a, b, floor=1e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def flags_match_outside_band(score_gpu, score_ref, thr, flag_gpu):
    """Bit-identical flags except scores within REL_TOL of the threshold; returns the band count."""
    ref_flag = np.asarray(score_ref, np.float32) > np.float32(thr)
    band = np.abs(np.asarray(score_ref, np.float64) - thr) <= REL_TOL * abs(thr)
    diff = ref_flag != flag_gpu.astype(bool)
    assert not np.any(diff & ~band), f"{int(np.sum(diff & ~band))} flags differ outside the tolerance band"
    return int(band.sum())


def engines():
    return [ops.ENGINE_FP32, ops.ENGINE_TC_BF16X3]


def to_dev(x, dev):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("stage", ["4dof", "openlab", "1dof"])
@pytest.mark.parametrize("scale", [1, 3])
@pytest.mark.parametrize("engine", engines())
def test_vae_score_vs_golden_and_oracle(cuda_dev, golden_dir, stage, scale, engine):
    g = np.load(golden_dir / f"synth_vae_{stage}_s{scale}.npz")
    s = synth.STAGES[stage]
    seed, N = int(g["seed"]), int(g["N"])
    sd = synth.stage_vae_weights(stage, seed=seed, scale=float(g["scale"]))
    X = synth.windows(N, s["T"], s["D"], seed=seed, amp=1.0 if scale == 1 else 2.5)
    eps = synth.eps(N, s["Z"], seed=seed)
    if engine == ops.ENGINE_TC_BF16X3 and stage == "1dof":
        with pytest.raises(ops.ShmfastError):          # H=32 is served by the fp32 engine only
            ops.VaeScorer(sd, cuda_dev, engine=engine)
        return
    vae = ops.VaeScorer(sd, cuda_dev, engine=engine)
    assert vae.engine == engine
    out = vae.score(ops.WindowSource(to_dev(X, cuda_dev), s["T"]), to_dev(eps, cuda_dev), want_latent=True, want_recon=True)
    score = out["score"].cpu().numpy()
    # against the reference module's own outputs
    print(f"{stage} s{scale} engine={engine}: score rel err {rel_err(score, g['score']):.2e}, "
          f"mu abs err {np.abs(out['mu'].cpu().numpy() - g['mu']).max():.2e}, "
          f"recon abs err {np.abs(out['recon'].cpu().numpy()[:8] - g['recon']).max():.2e}")
    assert rel_err(score, g["score"]) < REL_TOL
    # latent heads sit behind a LayerNorm that amplifies absolute error: 2e-5 for the fp32 engine, 1e-4 for
    # the 3-pass bf16 split (2^-17 operand error); the contract (north_star) is on scores and logits
    lat_atol = 2e-5 if engine == ops.ENGINE_FP32 else 1e-4
    assert np.allclose(out["mu"].cpu().numpy(), g["mu"], rtol=REL_TOL, atol=lat_atol)
    assert np.allclose(out["logvar"].cpu().numpy(), g["logvar"], rtol=REL_TOL, atol=lat_atol)
    assert np.allclose(out["recon"].cpu().numpy()[:8], g["recon"], rtol=REL_TOL, atol=1e-4)
    # against the oracle on the same inputs
    recon_o, mu_o, lv_o = O.vae_forward(sd, X, eps, np.float64)
    assert rel_err(score, O.mse_score(X.astype(np.float64), recon_o)) < REL_TOL


@pytest.mark.parametrize("stage", ["4dof", "openlab", "1dof"])
def test_vae_ragged_tail_idx_and_deterministic_mode(cuda_dev, stage):
    """N not a multiple of the CTA tile, gather list, n_dev clamp, eps=None (z=mu) mode."""
    s = synth.STAGES[stage]
    N = 133
    sd = synth.stage_vae_weights(stage, seed=5, scale=2.0)
    X = synth.windows(N, s["T"], s["D"], seed=5, amp=2.0)
    vae = ops.VaeScorer(sd, cuda_dev, engine=ops.ENGINE_FP32)
    Xd = to_dev(X, cuda_dev)
    score = vae.score(ops.WindowSource(Xd, s["T"]), None)["score"].cpu().numpy()
    recon_o, _, _ = O.vae_forward(sd, X, None, np.float64)
    ref = O.mse_score(X.astype(np.float64), recon_o)
    assert rel_err(score, ref) < REL_TOL
    sel = np.array([130, 2, 77, 5, 5, 131], dtype=np.int32)
    n_dev = torch.tensor([4], dtype=torch.int32, device=cuda_dev)
    out = vae.score(ops.WindowSource(Xd, s["T"]), None, n=6, idx=to_dev(sel, cuda_dev), n_dev=n_dev,
                    out=dict(score=torch.full((6,), -1.0, device=cuda_dev)))
    got = out["score"].cpu().numpy()
    assert rel_err(got[:4], ref[sel[:4]]) < REL_TOL
    assert np.all(got[4:] == -1.0)            # beyond *n_dev nothing is written
    # empty input
    assert vae.score(ops.WindowSource(Xd[:0], s["T"]), None)["score"].numel() == 0


def test_vae_encode_decode_entry_points(cuda_dev):
    from shmfast.models import fourdof
    torch.manual_seed(3)
    m = fourdof.TemporalVAE().to(cuda_dev).eval()
    sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    X = synth.windows(20, 100, 12, seed=9)
    x = to_dev(X, cuda_dev)
    with torch.no_grad():
        mu, lv = m.encode(x)
        z = mu + 0.3
        rec = m.decode(z, 100)
        torch.manual_seed(11)
        r2, mu2, lv2 = m(x)
        torch.manual_seed(11)
        eps = torch.randn((20, 16), device=cuda_dev)
    mu_o, lv_o = O.vae_encode(sd, X, np.float64)
    assert np.allclose(mu.cpu().numpy(), mu_o, rtol=REL_TOL, atol=2e-5)
    rec_o = O.vae_decode(sd, z.cpu().numpy().astype(np.float64), 100, np.float64)
    assert np.allclose(rec.cpu().numpy(), rec_o, rtol=REL_TOL, atol=1e-4)
    # forward() draws eps exactly like the reference: one randn of shape [B,Z] on the input's device
    r_o, _, _ = O.vae_forward(sd, X, eps.cpu().numpy(), np.float64)
    assert np.allclose(r2.cpu().numpy(), r_o, rtol=REL_TOL, atol=1e-4)
    assert torch.equal(mu2, mu)
    # state_dict round trip + weight refresh
    m2 = fourdof.VAE().to(cuda_dev).eval()
    m2.load_state_dict(m.state_dict())
    with torch.no_grad():
        assert torch.equal(m2.encode(x)[0], mu)
        m2.fc_mu.bias.add_(1.0)
        assert torch.allclose(m2.encode(x)[0], mu + 1.0, atol=1e-6)
    with pytest.raises(ops.ShmfastError):
        m(x.cpu())


def test_window_normalize_bit_exact(cuda_dev):
    # 4DOF: stride-1 gather from a series, std==0 guard, NaN / inf -> 0
    X = synth.series(301, 12, seed=3)
    X[17, 4] = np.nan; X[40, 7] = np.inf; X[41, 7] = -np.inf
    mean, std = synth.stats(12, seed=1, zero_std_channel=3)
    stdg = guard_std_4dof(std)
    ref = O.normalize_windows_4dof(O.make_windows(X, 100, 1), mean, O.guard_std_4dof(std))
    src = ops.WindowSource(to_dev(X, cuda_dev), 100, stride=1, mean=mean, std=stdg, nan_to_zero=True)
    got = ops.window_normalize(src).cpu().numpy()
    assert got.shape == ref.shape == (202, 100, 12)
    assert np.array_equal(got, ref)
    # openLAB gate: channel select [1,2,3], stride 20, clip 10 then NaN -> 0
    R = synth.series(1000, 4, seed=4, nan_frac=0.02) * 8
    mu, sd = synth.stats(3, seed=2)
    W = O.make_windows(R, 200, 20)
    ref = O.standardize_openlab(W[:, :, [1, 2, 3]], mu, sd, 10.0)
    src = ops.WindowSource(to_dev(R, cuda_dev), 200, stride=20, chan=[1, 2, 3], mean=mu, std=sd, clip=10.0, nan_to_zero=True)
    got = ops.window_normalize(src).cpu().numpy()
    assert np.array_equal(got, ref) and np.abs(got).max() <= 10
    # gather list + materialised windows source + odd sizes (scalar path)
    sel = np.array([5, 0, 40, 7], dtype=np.int32)
    src = ops.WindowSource(to_dev(W, cuda_dev), 200, chan=[3, 0, 2], mean=mu, std=sd, clip=10.0, nan_to_zero=True)
    got = ops.window_normalize(src, idx=to_dev(sel, cuda_dev)).cpu().numpy()
    assert np.array_equal(got, O.standardize_openlab(W[sel][:, :, [3, 0, 2]], mu, sd, 10.0))
    W7 = synth.windows(9, 7, 3, seed=8)
    got = ops.window_normalize(ops.WindowSource(to_dev(W7, cuda_dev), 7)).cpu().numpy()
    assert np.array_equal(got, W7)
    # short series -> no windows
    assert ops.WindowSource(to_dev(X[:50], cuda_dev), 100, stride=1).n_windows == 0


@pytest.mark.parametrize("N,frac", [(0, 0.5), (1, 1.0), (2047, 0.3), (2048, 0.0), (2049, 1.0), (100_003, 0.47), (1 << 20, 0.01)])
def test_compact_matches_np_where(cuda_dev, N, frac):
    rng = np.random.Generator(np.random.PCG64(N + 1))
    score = rng.random(N, dtype=np.float32)
    thr = float(np.float32(1.0 - frac)) if frac < 1.0 else -1.0
    if N > 10:
        score[3] = np.float32(thr)           # equal to the threshold -> not flagged (strict >)
        score[5] = np.nan                    # NaN > thr is False
    mask, idx = O.flag_compact(score, thr)
    flag, idx_d, count = ops.compact(to_dev(score, cuda_dev), thr)
    k = int(count.item())
    assert k == idx.size
    assert np.array_equal(idx_d[:k].cpu().numpy().astype(np.int64), idx)
    assert np.array_equal(flag.cpu().numpy().astype(bool), mask)


def test_cnn4dof_vs_golden(cuda_dev, golden_dir):
    g = np.load(golden_dir / "synth_cnn4dof.npz")
    sd = synth.cnn4dof_weights(seed=int(g["seed"]))
    z = synth.windows(int(g["N"]), 100, 12, seed=21)
    rec = synth.windows(int(g["N"]), 100, 12, seed=22, amp=0.7)
    xin = O.cnn4dof_inputs(z, rec)
    cnn = ops.Cnn4dof(sd, cuda_dev)
    logits, label, p = cnn.forward(to_dev(xin, cuda_dev), want_labels=True)
    logits = logits.cpu().numpy()
    assert np.allclose(logits, g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)
    lab_o, p_o = O.cnn4dof_labels(g["logits"])
    assert np.array_equal(label.cpu().numpy(), lab_o)
    assert np.allclose(p.cpu().numpy(), p_o, atol=1e-4)
    # nn.Module shim, same state_dict
    from shmfast.models import fourdof
    m = fourdof.CNNClassifier().to(cuda_dev).eval()
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    with torch.no_grad():
        assert np.allclose(m(to_dev(xin, cuda_dev)).cpu().numpy(), g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)


@pytest.mark.parametrize("engine", engines())
def test_cnnol_vs_golden(cuda_dev, golden_dir, engine):
    g = np.load(golden_dir / "synth_cnnol.npz")
    sd = synth.cnnol_weights(seed=int(g["seed"]))
    x = synth.windows(int(g["N"]), 200, 4, seed=31, amp=1.5)
    cnn = ops.CnnOpenLab(sd, cuda_dev, engine=engine)
    assert cnn.engine == engine
    logits, prob = cnn.forward(ops.WindowSource(to_dev(x, cuda_dev), 200), want_prob=True)
    err = np.max(np.abs(logits.cpu().numpy() - g["logits"]) / np.maximum(np.abs(g["logits"]), LOGIT_ABS / REL_TOL))
    print(f"cnnol engine={engine}: logits err (rel, floor {LOGIT_ABS / REL_TOL:g}) {err:.2e}")
    assert np.allclose(logits.cpu().numpy(), g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)
    _, p_o = O.cnnol_decision(g["logits"], 0.5)
    assert np.allclose(prob.cpu().numpy(), p_o, atol=1e-5)
    from shmfast.models import openlab
    m = openlab.CNN(dropout_rate=0.4).to(cuda_dev).eval()
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    with torch.no_grad():
        assert np.allclose(m(to_dev(x[:, None], cuda_dev)).cpu().numpy(), g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)


@pytest.mark.parametrize("engine", engines())
def test_cnnol_ragged_idx_and_device_count(cuda_dev, engine):
    """Flagged-subset call pattern of stage2_predict_cnn (10_test_hybrid_pipeline.py:265-302): idx list + device-side
    count, window counts that are not a multiple of any tile shape, NaN-bearing raw windows; against the oracle."""
    sd = synth.cnnol_weights(seed=77)
    rng = np.random.Generator(np.random.PCG64(78))
    N = 333
    x = synth.windows(N, 200, 4, seed=79, amp=2.0)
    x[5, 10:30, 2] = np.nan
    x[200, :, 0] = np.inf
    mu, sd_ = synth.stats(4, seed=80)
    idx = np.sort(rng.choice(N, size=101, replace=False)).astype(np.int32)
    cnn = ops.CnnOpenLab(sd, cuda_dev, engine=engine)
    src = ops.WindowSource(to_dev(x, cuda_dev), 200, mean=mu, std=sd_, clip=10.0, nan_to_zero=True)
    count = torch.tensor([77], dtype=torch.int32, device=cuda_dev)
    logits, prob = cnn.forward(src, n=101, idx=to_dev(idx, cuda_dev), n_dev=count, want_prob=True)
    xs = O.standardize_openlab(x[idx[:77]], mu, sd_, 10.0)
    ref = O.cnnol_forward(sd, xs[:, None].astype(np.float32))
    assert np.allclose(logits[:77].cpu().numpy(), ref, rtol=REL_TOL, atol=LOGIT_ABS)
    _, p_o = O.cnnol_decision(ref, 0.5)
    assert np.allclose(prob[:77].cpu().numpy(), p_o, atol=1e-5)
    for n in (1, 3, 17):                                  # tiny batches: partial groups in every block
        l2 = cnn.forward(src, n=n)
        xs = O.standardize_openlab(x[:n], mu, sd_, 10.0)
        assert np.allclose(l2.cpu().numpy(), O.cnnol_forward(sd, xs[:, None].astype(np.float32)), rtol=REL_TOL, atol=LOGIT_ABS)


def test_cnnol_two_chunks_position_independent(cuda_dev):
    """Tensor-core openLAB CNN over more windows than one internal chunk (9472) with a ragged tail: a window's logits do not depend
    on the chunk, tile pair or producer group it lands in (the fused operand producer stages 4 / 8 / 16 windows per tile) -- a
    shuffled subset recomputed through a gather list agrees to fp32 rounding of the GroupNorm statistics, a handful with the oracle."""
    sd = synth.cnnol_weights(seed=5)
    N = 9472 + 37
    x = synth.windows(N, 200, 4, seed=6, amp=1.5)
    cnn = ops.CnnOpenLab(sd, cuda_dev)
    src = ops.WindowSource(to_dev(x, cuda_dev), 200)
    a = cnn.forward(src)
    assert a.shape[0] == N and torch.isfinite(a).all()
    sel = torch.cat([torch.tensor([0, 9471, 9472, N - 1]), torch.randperm(N)[:203]]).to(torch.int32).to(cuda_dev)
    b = cnn.forward(src, n=int(sel.numel()), idx=sel)
    assert torch.allclose(b, a[sel.long()], rtol=1e-5, atol=1e-6)
    k = sel[:6].cpu().numpy()
    ref = O.cnnol_forward(sd, x[k][:, None].astype(np.float32))
    assert np.allclose(a[sel[:6].long()].cpu().numpy(), ref, rtol=REL_TOL, atol=LOGIT_ABS)


@pytest.mark.parametrize("engine", engines())
def test_openlab_hybrid_real_windows(cuda_dev, golden_dir, engine):
    g = np.load(golden_dir / "openlab_real_windows.npz")
    seed = int(g["seed"])
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=seed, scale=2.0), cuda_dev)
    cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=seed), cuda_dev, engine=engine)
    N = g["X_clean"].shape[0]
    eps = synth.eps(N, 8, seed=seed)
    Xc, Xr = to_dev(g["X_clean"], cuda_dev), to_dev(g["X_raw"], cuda_dev)
    src_gate = ops.WindowSource(Xc, 200, chan=list(g["channels_idx"]), mean=g["vae_mu"], std=g["vae_sd"], clip=10.0, nan_to_zero=True)
    src_raw = ops.WindowSource(Xr, 200, mean=g["cnn_mu"], std=g["cnn_sd"], clip=10.0, nan_to_zero=True)
    hyb = HybridOpenLab(vae, cnn, float(g["vae_thr"]), float(g["cnn_thr"]))
    r = hyb.run(src_gate, src_raw, to_dev(eps, cuda_dev))
    score = r["score"].cpu().numpy()
    assert rel_err(score, g["score"]) < REL_TOL
    band = flags_match_outside_band(score, g["score"], float(g["vae_thr"]), r["flag"].cpu().numpy())
    if band == 0:
        k = int(r["count"].item())
        assert np.array_equal(r["idx"][:k].cpu().numpy(), np.where(g["mask"])[0])
        assert np.allclose(r["logits"].cpu().numpy(), g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)
        assert np.allclose(r["prob"].cpu().numpy(), g["prob"], atol=1e-5)
        assert np.array_equal(r["pred"].cpu().numpy(), g["pred"])


def test_trained_4dof_hybrid(cuda_dev, golden_dir):
    """Reference-trained weights (numerically harsh: scores 0.2..260, |z| up to ~40) on real test windows."""
    p = golden_dir / "trained_4dof.npz"
    if not p.exists():
        pytest.skip("trained fixture not generated")
    g = np.load(p)
    vae_sd = {k[4:]: g[k] for k in g.files if k.startswith("vae.")}
    cnn_sd = {k[4:]: g[k] for k in g.files if k.startswith("cnn.")}
    W = g["W"]
    N = W.shape[0]
    thr = float(g["thr"])
    vae = ops.VaeScorer(vae_sd, cuda_dev)
    cnn = ops.Cnn4dof(cnn_sd, cuda_dev)
    src = ops.WindowSource(to_dev(W, cuda_dev), 100, mean=g["mean"], std=guard_std_4dof(g["std"]), nan_to_zero=True)
    eps1 = synth.eps(N, 16, seed=41)
    eps2 = synth.eps(int(g["idx"].size), 16, seed=42)
    r = Hybrid4dof(vae, cnn, thr).run(src, to_dev(eps1, cuda_dev), to_dev(eps2, cuda_dev))
    score = r["score"].cpu().numpy()
    assert rel_err(score, g["score"]) < REL_TOL
    band = flags_match_outside_band(score, g["score"], thr, r["flag"].cpu().numpy())
    assert band == 0, "fixture threshold sits mid-gap; no score may be inside the band"
    k = int(r["count"].item())
    assert np.array_equal(r["idx"][:k].cpu().numpy(), g["idx"])
    assert np.allclose(r["logits"].cpu().numpy(), g["logits"], rtol=REL_TOL, atol=LOGIT_ABS)
    assert np.array_equal(r["label"].cpu().numpy(), g["y_pred"])
    assert np.allclose(r["p_struct"].cpu().numpy(), g["p_struct"], atol=1e-4)
    y_pred, p_full = Hybrid4dof.scatter(r, N)
    assert np.array_equal(y_pred.cpu().numpy()[g["idx"]], g["y_pred"]) and int((y_pred == 0).sum()) == N - k
    # cnn_in is stack([z, (z - zhat)^2]) with channel 0 the normalised window, bit-exact
    assert np.array_equal(r["cnn_in"][:, 0].cpu().numpy(), g["Z"][g["idx"]])


def test_onedof_stitch_rmse(cuda_dev, golden_dir):
    g = np.load(golden_dir / "onedof_seen.npz")
    from shmfast.models import onedof
    sd = synth.stage_vae_weights("1dof", seed=int(g["seed"]), scale=2.0)
    data_t, mean, std = g["data_t"], g["mean"], g["std"]
    # the reference standardises the series first, then windows it (datasets.py:17-35)
    src = ops.WindowSource(to_dev(data_t, cuda_dev), 80, stride=1, mean=mean, std=std)
    assert src.n_windows == int(g["n_windows"])
    W = ops.window_normalize(src)
    assert np.array_equal(W[:4].cpu().numpy(), g["windows_f32_head"])
    # ALL windows: the reference's standardised series (datasets.standardize, cast to fp32 like 04_test_seen_variants.py:292) --
    # every window the reference builds is a slice of it
    ref_all = np.lib.stride_tricks.sliding_window_view(g["norm_series_f32"], 80, axis=0).transpose(0, 2, 1)
    assert ref_all.shape == tuple(W.shape) and np.array_equal(W.cpu().numpy(), ref_all)
    assert np.array_equal(W.cpu().numpy().astype(np.float64).sum(axis=(1, 2)), g["windows_checksum"])
    m = onedof.TemporalVAE().to(cuda_dev).eval()
    m.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
    eps = synth.eps(src.n_windows, 5, seed=int(g["seed"]))
    out = m.scorer().score(src, to_dev(eps, cuda_dev), want_recon=True, want_latent=True)
    assert np.allclose(out["mu"].cpu().numpy(), g["mu"], rtol=REL_TOL, atol=2e-5)
    series, rm = ops.stitch_segment_rmse(out["recon"], data_t.shape[0], 1, mean, std, to_dev(data_t, cuda_dev), 100)
    assert np.allclose(series.cpu().numpy(), g["recon_series"], rtol=REL_TOL, atol=1e-5)
    assert np.allclose(rm.cpu().numpy(), g["segment_rmse"], rtol=REL_TOL)
    # stitching itself is bit-exact in fp64 against the oracle on the same reconstructions
    rec = out["recon"].cpu().numpy()
    ref_series = O.destandardize(O.stitch_windows(rec, data_t.shape[0], 1), mean, std)
    assert np.array_equal(series.cpu().numpy(), ref_series)
    assert np.allclose(rm.cpu().numpy(), O.segment_rmse(data_t, ref_series, 100), rtol=1e-12)


@pytest.mark.parametrize("N,q", [(1, 99.0), (2, 50.0), (2010, 99.0), (256, 95.0), (100_000, 99.9), (4097, 0.0), (4097, 100.0), (300_000, 12.5)])
def test_percentile_bit_exact(cuda_dev, N, q):
    rng = np.random.Generator(np.random.PCG64(N))
    s = (rng.standard_normal(N).astype(np.float32)) ** 2
    if N > 3:
        s[1] = -s[1]
    got = float(ops.percentile(to_dev(s, cuda_dev), q).item())
    assert got == float(np.percentile(s, q))


@pytest.mark.parametrize("kind", ["constant", "sorted", "reversed", "two_values", "periodic_adversarial", "with_inf", "large", "unaligned"])
@pytest.mark.parametrize("q", [99.0, 50.0, 3.7])
def test_percentile_single_pass_edge_cases(cuda_dev, kind, q):
    """The one-read front end (sample bracket -> filter -> select among candidates) and its device-side fallback stay bit-exact:
    ties, monotone inputs, a periodic input built so that the strided sample sees only the large values (bracket misses both
    ranks -> full-array select), infinities, 2^23 scores, and a view that is not 16-byte aligned."""
    rng = np.random.Generator(np.random.PCG64(17))
    N = 1 << 19
    if kind == "constant":
        s = np.full(N, 1.25, np.float32)
    elif kind == "sorted":
        s = np.sort(rng.standard_normal(N).astype(np.float32) ** 2)
    elif kind == "reversed":
        s = np.sort(rng.standard_normal(N).astype(np.float32) ** 2)[::-1].copy()
    elif kind == "two_values":
        s = np.where(rng.random(N) < 0.3, 0.5, 2.0).astype(np.float32)
    elif kind == "periodic_adversarial":
        s = rng.random(N).astype(np.float32)
        s[:: N // 16384] += 100.0                                  # exactly the positions the strided sample reads
    elif kind == "with_inf":
        s = rng.standard_normal(N).astype(np.float32) ** 2
        s[rng.integers(0, N, 4000)] = np.inf
        s[rng.integers(0, N, 4000)] = -np.inf
    elif kind == "large":
        N = 1 << 23
        s = np.exp(rng.standard_normal(N)).astype(np.float32)
    else:
        s = (rng.standard_normal(N + 1).astype(np.float32) ** 2)
    if kind == "unaligned":
        d = to_dev(s, cuda_dev)[1:]
        s = s[1:]
    else:
        d = to_dev(s, cuda_dev)
    got = float(ops.percentile(d, q).item())
    ref = float(np.percentile(s, q))
    assert got == ref or (np.isnan(got) and np.isnan(ref)), (kind, q, got, ref)


def test_full_size_properties(cuda_dev):
    """BASELINE-size run (2^18 windows here, same code path as 2^20): size-independent properties --
    determinism, permutation equivariance through the gather list, series-gather == materialised."""
    s = synth.STAGES["4dof"]
    N = 1 << 18
    rows = N + s["T"] - 1
    sd = synth.stage_vae_weights("4dof", seed=0)
    series = to_dev(synth.series(rows, 12, seed=0), cuda_dev)
    eps = to_dev(synth.eps(N, 16, seed=0), cuda_dev)
    vae = ops.VaeScorer(sd, cuda_dev)
    src = ops.WindowSource(series, 100, stride=1)
    a = vae.score(src, eps)["score"]
    b = vae.score(src, eps)["score"]
    assert torch.equal(a, b)                                   # deterministic
    assert torch.isfinite(a).all()
    # a sample of windows recomputed from materialised copies agrees bit for bit
    sel = torch.randint(0, N, (4096,), device=cuda_dev, dtype=torch.int32)
    Wm = ops.window_normalize(src, idx=sel)
    c = vae.score(ops.WindowSource(Wm, 100), eps[sel.long()].contiguous())["score"]
    assert torch.equal(c, a[sel.long()])
    # and against the oracle on a handful
    k = sel[:16].long().cpu().numpy()
    Wk = Wm[:16].cpu().numpy()
    ro, _, _ = O.vae_forward(sd, Wk, eps[sel[:16].long()].cpu().numpy(), np.float64)
    assert rel_err(a[k].cpu().numpy(), O.mse_score(Wk.astype(np.float64), ro)) < REL_TOL
    flag, idx, count = ops.compact(a, float(np.percentile(a.cpu().numpy(), 99.0)))
    kk = int(count.item())
    assert abs(kk - 0.01 * N) <= 2 and bool((idx[1:kk] > idx[:kk - 1]).all())


@pytest.mark.gpu
@pytest.mark.parametrize("N", [(1 << 17) + 128 * 3 + 5, 129, 1])
def test_full_size_properties_openlab(cuda_dev, N):
    """The two-tile H=64 scorer (vae_tc_dual.cuh) at stream size and at the pairing edge cases (odd number of tiles, a lone
    ragged tile, one window): deterministic, and a window's score does not depend on the tile / pair it lands in -- a random
    subset re-scored through a gather list, and as materialised windows, agrees bit for bit; a handful against the oracle."""
    s = synth.STAGES["openlab"]
    sd = synth.stage_vae_weights("openlab", seed=0)
    rows = (N - 1) * s["stride"] + s["T"]
    series = to_dev(synth.series(rows, s["D"], seed=3), cuda_dev)
    eps = to_dev(synth.eps(N, s["Z"], seed=3), cuda_dev)
    vae = ops.VaeScorer(sd, cuda_dev)
    src = ops.WindowSource(series, s["T"], stride=s["stride"])
    a = vae.score(src, eps)["score"]
    assert torch.equal(a, vae.score(src, eps)["score"]) and torch.isfinite(a).all()
    m = min(N, 1000)
    sel = torch.randperm(N, device=cuda_dev)[:m].to(torch.int32)
    e_sel = eps[sel.long()].contiguous()
    b = vae.score(src, e_sel, idx=sel)["score"]                # same windows, different tiles and pair partners
    assert torch.equal(b, a[sel.long()])
    Wm = ops.window_normalize(src, idx=sel)
    c = vae.score(ops.WindowSource(Wm, s["T"]), e_sel)["score"]
    assert torch.equal(c, a[sel.long()])
    k = min(m, 8)
    Wk = Wm[:k].cpu().numpy()
    ro, _, _ = O.vae_forward(sd, Wk, e_sel[:k].cpu().numpy(), np.float64)
    assert rel_err(a[sel[:k].long()].cpu().numpy(), O.mse_score(Wk.astype(np.float64), ro)) < REL_TOL


def test_4dof_hybrid_at_the_repos_flag_rate_vs_port(cuda_dev):
    """BASELINE configs[2] at the repo's real test-set flag rate (47 %): 8,192 windows gathered from a
    series, scored, thresholded at the score's P53, second pass + CNN on ~3,850 flagged windows -- every score / flag / index / logit /
    label against the reference's wiring on torch CPU kernels (oracle.torch_port.hybrid_4dof, same eps).  Band windows are counted."""
    from oracle import torch_port as TP
    N, T, D, Z = 8192, 100, 12, 16
    vae_sd, cnn_sd = synth.stage_vae_weights("4dof", seed=11, scale=2.0), synth.cnn4dof_weights(seed=11)
    series = synth.series(N + T - 1, D, seed=11)
    mean, std = synth.stats(D, seed=11)
    std = guard_std_4dof(std)
    eps1, eps2 = synth.eps(N, Z, seed=12), synth.eps(N, Z, seed=13)
    port_v, port_c = TP.VaePort(vae_sd), TP.Cnn4dofPort(cnn_sd)
    W = np.stack([series[i:i + T] for i in range(N)]).astype(np.float32)
    Zw = np.nan_to_num((W - mean[None, None]) / std[None, None], nan=0.0, posinf=0.0, neginf=0.0).astype(np.float32)
    s_ref = TP.vae_scores_batched(port_v, Zw, eps1, 512)
    ss = np.sort(s_ref.astype(np.float64))
    k53 = int(0.53 * N)
    thr = float(0.5 * (ss[k53 - 1] + ss[k53]))                       # P53, placed mid-gap
    ref = TP.hybrid_4dof(port_v, port_c, series, mean, std, thr, eps1=eps1, eps2=eps2)
    n_ref = int(ref["idx"].size)
    assert 0.45 * N < n_ref < 0.49 * N
    vae, cnn = ops.VaeScorer(vae_sd, cuda_dev), ops.Cnn4dof(cnn_sd, cuda_dev)
    src = ops.WindowSource(to_dev(series, cuda_dev), T, stride=1, mean=mean, std=std, nan_to_zero=True)
    res = Hybrid4dof(vae, cnn, thr).run_dense(src, to_dev(eps1, cuda_dev), to_dev(eps2, cuda_dev))
    torch.cuda.synchronize()
    score = res["score"].cpu().numpy()
    assert rel_err(score, ref["score"]) < REL_TOL
    band = flags_match_outside_band(score, ref["score"], thr, res["flag"].cpu().numpy())
    st = res["status"].cpu().numpy()
    assert st[1] == 0
    if band == 0:
        assert st[0] == n_ref and np.array_equal(res["idx"][:n_ref].cpu().numpy(), ref["idx"])
        lg = res["logits"][:n_ref].cpu().numpy()
        assert np.allclose(lg, ref["logits"], rtol=REL_TOL, atol=LOGIT_ABS)
        tie = np.abs(ref["logits"][:, 0] - ref["logits"][:, 1]) <= 2 * LOGIT_ABS          # argmax may flip only on a logit tie
        y = res["y_pred"].cpu().numpy()
        assert np.array_equal(y[ref["idx"]][~tie], ref["y_pred"][ref["idx"]][~tie])
        assert int((y != 0).sum()) == n_ref
        assert np.allclose(res["p_full"].cpu().numpy(), ref["p_struct"], atol=2e-4)
    print(f"4DOF 47 %: flagged {n_ref}/{N}, band {band}, score rel err {rel_err(score, ref['score']):.2e}")


def test_openlab_hybrid_at_the_repos_flag_rate_vs_port(cuda_dev):
    """BASELINE configs[3] at the repo's real test-set flag rate (38 %): 4,096 windows (T 200, stride 20, NaN runs
    in the raw channels), gate on 3 channels, CNN on the flagged raw windows with its own statistics; vs oracle.torch_port.hybrid_openlab."""
    from oracle import torch_port as TP
    N = 4096
    vae_sd, cnn_sd = synth.stage_vae_weights("openlab", seed=21, scale=2.0), synth.cnnol_weights(seed=21)
    series = synth.series((N - 1) * 20 + 200, 4, seed=21, nan_frac=0.0007)
    vmu, vsd = synth.stats(3, seed=1)
    cmu, csd = synth.stats(4, seed=2)
    eps = synth.eps(N, 8, seed=22)
    port_v, port_c = TP.VaePort(vae_sd), TP.CnnOpenLabPort(cnn_sd)
    first = TP.hybrid_openlab(port_v, port_c, series, [1, 2, 3], vmu, vsd, cmu, csd, float("inf"), 0.5, eps=eps)
    ss = np.sort(first["score"].astype(np.float64))
    k62 = int(0.62 * N)
    thr = float(0.5 * (ss[k62 - 1] + ss[k62]))
    ref = TP.hybrid_openlab(port_v, port_c, series, [1, 2, 3], vmu, vsd, cmu, csd, thr, 0.13, eps=eps)
    n_ref = int(ref["mask"].sum())
    assert 0.36 * N < n_ref < 0.40 * N
    vae, cnn = ops.VaeScorer(vae_sd, cuda_dev), ops.CnnOpenLab(cnn_sd, cuda_dev)
    sd = to_dev(series, cuda_dev)
    g = ops.WindowSource(sd, 200, stride=20, chan=[1, 2, 3], mean=vmu, std=vsd, clip=10.0, nan_to_zero=True)
    r = ops.WindowSource(sd, 200, stride=20, mean=cmu, std=csd, clip=10.0, nan_to_zero=True)
    res = HybridOpenLab(vae, cnn, thr, 0.13).run_dense(g, r, to_dev(eps, cuda_dev))
    torch.cuda.synchronize()
    score = res["score"].cpu().numpy()
    assert rel_err(score, ref["score"]) < REL_TOL
    band = flags_match_outside_band(score, ref["score"], thr, res["flag"].cpu().numpy())
    if band == 0:
        assert int(res["status"][0].item()) == n_ref
        assert np.array_equal(res["idx"][:n_ref].cpu().numpy(), np.where(ref["mask"])[0])
        prob = res["prob"][:n_ref].cpu().numpy()
        assert prob.dtype == np.float64 and np.allclose(prob, ref["prob"], rtol=REL_TOL, atol=2e-4)
        near = np.abs(ref["prob"] - 0.13) < 2e-4
        assert np.array_equal(res["pred"][:n_ref].cpu().numpy()[~near], ref["pred"][~near])
        y3 = res["y_pred"].cpu().numpy()
        assert int((y3 != 0).sum()) == n_ref and np.array_equal(y3[np.where(ref["mask"])[0]][~near], 1 + ref["pred"][~near])
    print(f"openLAB 38 %: flagged {n_ref}/{N}, band {band}, score rel err {rel_err(score, ref['score']):.2e}")
