"""Host-buffer streaming (shmfast.stream): chunked results equal the direct device-resident runs, chunk by chunk, with
the copies on side streams; the scatter kernels equal the reference's y_pred / hyb_score_full scatter
(06_test_full_pipeline.py:336,356,368-372) without a host read of the flagged count; the fused C entries
(shm_hybrid4dof_score / shm_hybridol_score) equal the staged runs bit for bit."""
import numpy as np
import pytest
import torch

from shmfast import ops, synth
from shmfast.pipeline import Hybrid4dof, HybridOpenLab, guard_std_4dof
from shmfast.stream import HostStream

pytestmark = pytest.mark.gpu


def test_scatter_flagged_matches_host_scatter(cuda_dev):
    g = torch.Generator().manual_seed(3)
    n, cap, k = 1000, 300, 137
    idx_h = torch.sort(torch.randperm(n, generator=g)[:k]).values.to(torch.int32)
    idx = torch.full((cap,), 123456, dtype=torch.int32)          # slots past the count hold garbage
    idx[:k] = idx_h
    label = torch.randint(1, 3, (cap,), generator=g)
    p = torch.rand((cap,), generator=g)
    cnt = torch.tensor([k], dtype=torch.int32, device=cuda_dev)
    y, pf = ops.scatter_flagged_4dof(idx.to(cuda_dev), cnt, cap, label.to(cuda_dev), p.to(cuda_dev), n)
    y_ref = torch.zeros(n, dtype=torch.int64); y_ref[idx_h.long()] = label[:k]
    p_ref = torch.zeros(n); p_ref[idx_h.long()] = p[:k]
    assert torch.equal(y.cpu(), y_ref) and torch.equal(pf.cpu(), p_ref)
    y0, _ = ops.scatter_flagged_4dof(idx.to(cuda_dev), torch.zeros(1, dtype=torch.int32, device=cuda_dev), cap, label.to(cuda_dev),
                                     p.to(cuda_dev), n)
    assert int(y0.abs().sum()) == 0
    # openLAB: pred_bin = prob >= thr in fp64 (10_test_hybrid_pipeline.py:300-301), dense labels 0 / 1 (SF) / 2 (ST)
    prob = torch.rand((cap,), generator=g, dtype=torch.float64)
    prob[5] = 0.13                                                # == thr -> structural (>=)
    pred, y3, pfull = ops.scatter_flagged_openlab(idx.to(cuda_dev), cnt, cap, prob.to(cuda_dev), 0.13, n)
    pred_ref = (prob[:k].numpy() >= 0.13).astype(np.int64)
    assert np.array_equal(pred.cpu().numpy()[:k], pred_ref) and pred_ref[5] == 1
    y3_ref = np.zeros(n, np.int64); y3_ref[idx_h.numpy()] = 1 + pred_ref
    pf_ref = np.zeros(n, np.float64); pf_ref[idx_h.numpy()] = prob[:k].numpy()
    assert np.array_equal(y3.cpu().numpy(), y3_ref) and np.array_equal(pfull.cpu().numpy(), pf_ref)


def test_fused_hybrid4dof_entry_equals_staged_run_and_reports_overflow(cuda_dev):
    """shm_hybrid4dof_score (one C call, no host sync) == score -> compact -> rescore -> CNN -> scatter run stage by stage."""
    T, D, Z, N = 100, 12, 16, 900
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=7, scale=2.0), cuda_dev)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=7), cuda_dev)
    mean, std = synth.stats(D, seed=7)
    series = torch.from_numpy(synth.series(N + T - 1, D, seed=7)).to(cuda_dev)
    src = ops.WindowSource(series, T, stride=1, mean=mean, std=guard_std_4dof(std), nan_to_zero=True)
    eps1 = torch.randn((N, Z), device=cuda_dev)
    eps2 = torch.randn((N, Z), device=cuda_dev)
    s0 = vae.score(src, eps1)["score"]
    thr = float(torch.quantile(s0, 0.55))
    hyb = Hybrid4dof(vae, cnn, thr)
    ref = hyb.run(src, eps1, eps2)                                   # staged, reads the count back
    k = int(ref["count"].item())
    assert 0 < k < N
    y_ref, p_ref = Hybrid4dof.scatter(ref, N)
    res = hyb.run_dense(src, eps1, eps2)                             # max_flagged = n: never overflows
    torch.cuda.synchronize()
    st = res["status"].cpu().numpy()
    assert st[0] == k and st[1] == 0
    assert torch.equal(res["score"], ref["score"]) and torch.equal(res["flag"], ref["flag"])
    assert torch.equal(res["idx"][:k], ref["idx"][:k])
    assert torch.equal(res["logits"][:k], ref["logits"]) and torch.equal(res["label"][:k], ref["label"])
    assert torch.equal(res["y_pred"], y_ref) and torch.equal(res["p_full"], p_ref)
    # capacity below the flagged count: the first `cap` flagged windows are attributed, the overflow is REPORTED
    cap = k // 2
    res2 = hyb.run_dense(src, eps1, eps2, max_flagged=cap)
    torch.cuda.synchronize()
    st2 = res2["status"].cpu().numpy()
    assert st2[0] == k and st2[1] == 1
    assert torch.equal(res2["label"][:cap], ref["label"][:cap])
    sel = ref["idx"][:cap].long()
    assert torch.equal(res2["y_pred"][sel], ref["label"][:cap]) and int((res2["y_pred"] != 0).sum()) == cap
    y2, _ = Hybrid4dof.scatter(res2, N)
    assert torch.equal(y2, res2["y_pred"])


def test_fused_hybridol_entry_equals_staged_run(cuda_dev):
    N = 600
    vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=3, scale=2.0), cuda_dev)
    cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=3), cuda_dev)
    series = torch.from_numpy(synth.series((N - 1) * 20 + 200, 4, seed=3, nan_frac=0.0007)).to(cuda_dev)
    vmu, vsd = synth.stats(3, seed=1)
    cmu, csd = synth.stats(4, seed=2)
    g = ops.WindowSource(series, 200, stride=20, chan=[1, 2, 3], mean=vmu, std=vsd, clip=10.0, nan_to_zero=True)
    r = ops.WindowSource(series, 200, stride=20, mean=cmu, std=csd, clip=10.0, nan_to_zero=True)
    eps = torch.randn((N, 8), device=cuda_dev)
    thr = float(torch.quantile(vae.score(g, eps)["score"], 0.7))
    hyb = HybridOpenLab(vae, cnn, thr, 0.5)
    ref = hyb.run(g, r, eps)
    k = int(ref["count"].item())
    assert 0 < k < N
    res = hyb.run_dense(g, r, eps, max_flagged=k + 17)
    torch.cuda.synchronize()
    st = res["status"].cpu().numpy()
    assert st[0] == k and st[1] == 0
    assert torch.equal(res["score"], ref["score"]) and torch.equal(res["idx"][:k], ref["idx"][:k])
    assert torch.equal(res["prob"][:k], ref["prob"]) and torch.equal(res["pred"][:k], ref["pred"][:k])
    assert torch.equal(res["y_pred"], ref["y_pred"]) and torch.equal(res["prob_full"], ref["prob_full"])
    assert set(np.unique(res["y_pred"].cpu().numpy())) <= {0, 1, 2} and int((res["y_pred"] != 0).sum()) == k


def test_stream_4dof_chunks_equal_direct_runs(cuda_dev):
    T, D, Z, N = 100, 12, 16, 700
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=5, scale=2.0), cuda_dev)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=5), cuda_dev)
    mean, std = synth.stats(D, seed=5)
    std = guard_std_4dof(std)
    chunks = [torch.from_numpy(synth.series(N + T - 1 - 13 * c, D, seed=10 + c)).pin_memory() for c in range(5)]   # ragged: fewer rows each chunk
    eps1 = [torch.randn((N, Z), generator=torch.Generator().manual_seed(20 + c)).to(cuda_dev) for c in range(5)]
    eps2 = [torch.randn((N, Z), generator=torch.Generator().manual_seed(40 + c)).to(cuda_dev) for c in range(5)]
    s0 = vae.score(ops.WindowSource(chunks[0].to(cuda_dev), T, stride=1, mean=mean, std=std, nan_to_zero=True), eps1[0])["score"]
    thr = float(torch.quantile(s0, 0.6))
    hyb = Hybrid4dof(vae, cnn, thr)

    def step(sd, i):
        src = ops.WindowSource(sd, T, stride=1, mean=mean, std=std, nan_to_zero=True)
        n = src.n_windows
        res = hyb.run(src, eps1[i], eps2[i], n=n, sync_count=False, max_flagged=n)
        return dict(score=res["score"], y_pred=res["y_pred"], p_struct=res["p_full"], count=res["count"].reshape(1))

    pipe = HostStream(cuda_dev, (N + T - 1, D), {"score": ((N,), torch.float32), "y_pred": ((N,), torch.int64),
                                                "p_struct": ((N,), torch.float32), "count": ((1,), torch.int32)})
    seen = 0
    for i, host in pipe.run(iter(chunks), step):
        assert i == seen
        seen += 1
        src = ops.WindowSource(chunks[i].to(cuda_dev), T, stride=1, mean=mean, std=std, nan_to_zero=True)
        n = src.n_windows
        ref = hyb.run(src, eps1[i], eps2[i], n=n)                       # the synchronous path (reads the count back)
        y_ref, p_ref = Hybrid4dof.scatter(ref, n)
        assert int(host["count"][0]) == int(ref["count"].item()) > 0
        assert torch.equal(host["score"][:n], ref["score"].cpu())
        assert torch.equal(host["y_pred"][:n], y_ref.cpu())
        assert torch.equal(host["p_struct"][:n], p_ref.cpu())
    assert seen == 5
    assert pipe.h2d_bytes == sum(c.numel() * 4 for c in chunks)


def test_stream_rejects_unpinned_and_oversized(cuda_dev):
    pipe = HostStream(cuda_dev, (64, 4), {"x": ((64,), torch.float32)})
    with pytest.raises(ops.ShmfastError):
        list(pipe.run(iter([torch.zeros((64, 4))]), lambda sd, i: dict(x=sd[:, 0])))
    with pytest.raises(ops.ShmfastError):
        list(pipe.run(iter([torch.zeros((65, 4)).pin_memory()]), lambda sd, i: dict(x=sd[:64, 0])))
    with pytest.raises(ops.ShmfastError):
        HostStream(torch.device("cpu"), (4, 4), {})


def test_async_hybrid_with_nothing_flagged(cuda_dev):
    """sync_count=False sizes the second pass by capacity and lets the kernels read the flagged count on the device: a chunk
    with no flagged window (count = 0) must run through rescore / CNN / scatter as a no-op."""
    T, D, Z, N = 100, 12, 16, 300
    vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=2), cuda_dev)
    cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=2), cuda_dev)
    mean, std = synth.stats(D, seed=2)
    series = torch.from_numpy(synth.series(N + T - 1, D, seed=2)).to(cuda_dev)
    src = ops.WindowSource(series, T, stride=1, mean=mean, std=guard_std_4dof(std), nan_to_zero=True)
    eps1 = torch.randn((N, Z), device=cuda_dev)
    eps2 = torch.randn((N, Z), device=cuda_dev)
    res = Hybrid4dof(vae, cnn, float("inf")).run(src, eps1, eps2, sync_count=False, max_flagged=128)
    y, p = res["y_pred"], res["p_full"]
    torch.cuda.synchronize()
    assert int(res["count"].item()) == 0 and int(res["flag"].sum()) == 0
    assert int(y.abs().sum()) == 0 and float(p.abs().sum()) == 0.0
    assert torch.equal(res["score"], vae.score(src, eps1)["score"])
