"""Host-buffer streaming (shmfast.stream): chunked results equal the direct device-resident runs, chunk by chunk, with
the copies on side streams; scatter_flagged equals the reference's y_pred / hyb_score_full scatter
(06_test_full_pipeline.py:336,356,368-372) without a host read of the flagged count."""
import numpy as np
import pytest
import torch

from shmfast import ops, synth
from shmfast.pipeline import Hybrid4dof, HybridOpenLab, guard_std_4dof
from shmfast.stream import HostStream, scatter_flagged

pytestmark = pytest.mark.gpu


def test_scatter_flagged_matches_host_scatter(cuda_dev):
    g = torch.Generator().manual_seed(3)
    n, cap, k = 1000, 300, 137
    idx_h = torch.sort(torch.randperm(n, generator=g)[:k]).values.to(torch.int32)
    idx = torch.full((cap,), 123456, dtype=torch.int32)          # slots past the count hold garbage
    idx[:k] = idx_h
    label = torch.randint(1, 3, (cap,), generator=g)
    p = torch.rand((cap,), generator=g)
    y, pf = scatter_flagged(idx.to(cuda_dev), torch.tensor([k], dtype=torch.int32, device=cuda_dev), n,
                            [label.to(cuda_dev), p.to(cuda_dev)], [torch.int64, torch.float32])
    y_ref = torch.zeros(n, dtype=torch.int64); y_ref[idx_h.long()] = label[:k]
    p_ref = torch.zeros(n); p_ref[idx_h.long()] = p[:k]
    assert torch.equal(y.cpu(), y_ref) and torch.equal(pf.cpu(), p_ref)
    y0, = scatter_flagged(idx.to(cuda_dev), torch.zeros(1, dtype=torch.int32, device=cuda_dev), n, [label.to(cuda_dev)], [torch.int64])
    assert int(y0.abs().sum()) == 0


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
        y, p = scatter_flagged(res["idx"], res["count"], n, [res["label"], res["p_struct"]], [torch.int64, torch.float32])
        return dict(score=res["score"], y_pred=y, p_struct=p, count=res["count"].reshape(1))

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
    y, p = scatter_flagged(res["idx"], res["count"], N, [res["label"], res["p_struct"]], [torch.int64, torch.float32])
    torch.cuda.synchronize()
    assert int(res["count"].item()) == 0 and int(res["flag"].sum()) == 0
    assert int(y.abs().sum()) == 0 and float(p.abs().sum()) == 0.0
    assert torch.equal(res["score"], vae.score(src, eps1)["score"])
