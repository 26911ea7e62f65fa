"""Parity of the training step (BASELINE config 5; 4DOF/Scripts/03_train_vae.py:260-271) on the CUDA path, through
the C ABI (shm_vae_train_forward / _backward / shm_vae_elbo_grad / shm_adam_clip_step), against
  * one optimisation step executed on the reference's own TemporalVAE (tests/golden/train_step_*.npz), and
  * oracle/torch_port.VaeTrainPort (torch autograd on the CPU) on the same seeded inputs, incl. dropout masks.

Tolerances (fp32 arithmetic, different summation order than torch's BPTT): loss 1e-5 relative; every gradient tensor
max|err| <= GRAD_TOL * max|ref| (1e-4, the north-star's relative tolerance, measured ~1e-6); the optimiser kernel on
identical gradients: 1e-6 absolute on the parameters."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O
from oracle import torch_port as TP
from shmfast import synth, train
from shmfast.models import fourdof

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4
LOSS_TOL = 1e-5


def _flat(sd, names):
    return np.concatenate([np.asarray(sd[n], np.float32).reshape(-1) for n in names])


def _grad_errs(got, ref, port):
    """max |err| / max |ref| per parameter tensor."""
    errs, o = {}, 0
    for n in port.names:
        k = port.param(n).numel()
        r = ref[o:o + k]
        errs[n] = float(np.max(np.abs(got[o:o + k] - r)) / (np.max(np.abs(r)) + 1e-20))
        o += k
    return errs


def _run_gpu_step(dev, sd, names, cfg, X, eps, kl_w, masks=(None, None), p=0.0):
    h = train.VaeTrainHandle(cfg, X.shape[1], X.shape[0], dev)
    flat = torch.from_numpy(_flat(sd, names)).to(dev)
    x = torch.from_numpy(X).to(dev)
    me = None if masks[0] is None else torch.from_numpy(masks[0]).to(dev)
    md = None if masks[1] is None else torch.from_numpy(masks[1]).to(dev)
    xhat, mu, lv = h.forward(flat, x, torch.from_numpy(eps).to(dev), me, md, p)
    loss3, d_xhat, d_mu, d_lv = train.elbo_grad(x, xhat, mu, lv, kl_w)
    grads = h.backward(flat, d_xhat, d_mu, d_lv)
    torch.cuda.synchronize()
    h.close()
    return flat, xhat, loss3.cpu().numpy(), grads


@pytest.mark.parametrize("name", ["small", "full"])
def test_train_step_vs_reference_golden(cuda_dev, golden_dir, name):
    g = np.load(golden_dir / f"train_step_{name}.npz")
    D, Z, H, L, B, T = (int(g[k]) for k in "DZHLBT")
    seed, kl_w = int(g["seed"]), float(g["kl_w"])
    sd = synth.vae_weights(D, H, Z, L, True, seed=seed)
    names = TP.vae_param_names(sd)
    assert names == [str(n) for n in g["names"]]          # flat layout == list(model.parameters()) of the reference
    cfg = train._lib.VaeCfg(D, H, Z, L, 1, 1e-5, 0)
    flat, xhat, loss3, grads = _run_gpu_step(cuda_dev, sd, names, cfg, synth.windows(B, T, D, seed=seed), synth.eps(B, Z, seed=seed), kl_w)
    assert np.allclose(loss3, g["loss3"], rtol=LOSS_TOL), (loss3, g["loss3"])
    gn = grads.cpu().numpy()
    if name == "small":
        assert np.max(np.abs(xhat.cpu().numpy() - g["xhat"])) <= 1e-5
    o, worst = 0, 0.0
    for n in names:
        k = int(np.prod(sd[n].shape))
        got, ref = gn[o:o + k], g["g:" + n].reshape(-1)
        if name == "full":
            assert abs(np.linalg.norm(got.astype(np.float64)) - float(g["n:" + n])) <= GRAD_TOL * float(g["n:" + n]), n
            got = got[g["i:" + n]]
        e = float(np.max(np.abs(got - ref)) / (np.max(np.abs(ref)) + 1e-20))
        worst = max(worst, e)
        assert e <= GRAD_TOL, (n, e)
        o += k
    print(f"train golden {name}: worst gradient err (rel. to tensor max) {worst:.2e}")
    # clip + Adam on the device gradient; total norm as clip_grad_norm_ returned it on the reference
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    norm2 = train.adam_clip_step(flat, grads, m, v, 1, 1e-3, weight_decay=1e-5, max_norm=2.0)
    assert abs(float(norm2[1]) - float(g["total_norm"])) <= LOSS_TOL * float(g["total_norm"])


@pytest.mark.parametrize("case", [
    # (D, Z, H, L, has_ln, B, T, dropout p, kl_w)         B picks the windows-per-CTA variant of the recurrence
    (12, 16, 128, 2, True, 64, 100, 0.3, 0.05),            # 4DOF, 1 window / CTA
    (12, 16, 128, 2, True, 256, 100, 0.3, 1.0),            # 4DOF at the reference's batch size (03_train_vae.py:53), 2 / CTA
    (12, 16, 128, 2, True, 301, 100, 0.0, 0.5),            # ragged, 4 / CTA
    (12, 5, 32, 2, False, 40, 80, 0.2, 0.3),               # 1_DOF shape: no LayerNorm
    (3, 8, 64, 1, True, 96, 200, 0.0, 0.7),                # openLAB shape: single layer, T=200
])
def test_train_step_vs_port(cuda_dev, case):
    D, Z, H, L, has_ln, B, T, p, kl_w = case
    sd = synth.vae_weights(D, H, Z, L, has_ln, seed=51)
    port = TP.VaeTrainPort(sd).train()
    names = port.names
    X, eps = synth.windows(B, T, D, seed=52), synth.eps(B, Z, seed=53)
    masks = (None, None)
    if p > 0 and L > 1:
        rng = np.random.Generator(np.random.PCG64(54))
        masks = tuple((rng.random((L - 1, B, T, H)) >= p).astype(np.uint8) for _ in range(2))
    p0 = TP.flat_params(port)
    opt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
    max_norm = 0.05                                        # small enough that clipping engages
    me = None if masks[0] is None else torch.from_numpy(masks[0])
    md = None if masks[1] is None else torch.from_numpy(masks[1])
    loss_ref, g_ref, total_ref = TP.train_step_port(port, opt, torch.from_numpy(X), torch.from_numpy(eps), kl_w, me, md, p, max_norm)
    cfg = train._lib.VaeCfg(D, H, Z, L, 1 if has_ln else 0, 1e-5, 0)
    flat, _, loss3, grads = _run_gpu_step(cuda_dev, sd, names, cfg, X, eps, kl_w, masks, p)
    assert np.allclose(loss3, loss_ref, rtol=LOSS_TOL), (loss3, loss_ref)
    errs = _grad_errs(grads.cpu().numpy(), g_ref, port)
    worst = max(errs, key=errs.get)
    print(f"case {case}: worst gradient err {errs[worst]:.2e} ({worst}); total norm {total_ref:.4f}")
    assert errs[worst] <= GRAD_TOL, (worst, errs[worst])
    # optimiser parity on IDENTICAL gradients (the oracle's), so Adam's sign-like first step cannot amplify noise
    g_dev = torch.from_numpy(g_ref).to(cuda_dev)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    norm2 = train.adam_clip_step(flat, g_dev, m, v, 1, 1e-3, weight_decay=1e-5, max_norm=max_norm)
    assert total_ref > max_norm and abs(float(norm2[1]) - total_ref) <= LOSS_TOL * total_ref
    assert np.max(np.abs(flat.cpu().numpy() - TP.flat_params(port))) <= 1e-6
    p_np, m_np, v_np, _ = O.adam_clip_step(p0, g_ref, np.zeros_like(p0), np.zeros_like(p0), 1, max_norm=max_norm)
    assert np.allclose(m.cpu().numpy(), m_np, rtol=2e-6, atol=1e-12) and np.allclose(v.cpu().numpy(), v_np, rtol=2e-5, atol=1e-16)


def test_reference_loop_through_the_shim(cuda_dev):
    """The reference's own lines (03_train_vae.py:262-270) on the drop-in TemporalVAE: forward in train() mode,
    torch loss, loss.backward(), clip_grad_norm_, torch.optim.Adam.step -- gradients from the CUDA BPTT."""
    import torch.nn.functional as F
    torch.manual_seed(3)
    vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.0).to(cuda_dev)
    sd = {k: v.detach().cpu().numpy() for k, v in vae.state_dict().items()}
    port = TP.VaeTrainPort(sd).train()
    assert port.names == [n for n, _ in vae.named_parameters()]
    xb = torch.from_numpy(synth.windows(32, 100, 12, seed=61)).to(cuda_dev)
    opt = torch.optim.Adam(vae.parameters(), lr=1e-3, weight_decay=1e-5)
    vae.train()
    gen_state = torch.cuda.get_rng_state(cuda_dev)
    xhat, mu, logvar = vae(xb)
    torch.cuda.set_rng_state(gen_state, cuda_dev)
    eps = torch.randn((32, 16), dtype=torch.float32, device=cuda_dev)      # the draw forward() just made
    recon = F.mse_loss(xhat, xb, reduction="mean")
    kl = -0.5 * torch.mean(1.0 + logvar - mu.pow(2) - logvar.exp())
    loss = recon + 0.2 * kl
    opt.zero_grad(set_to_none=True)
    loss.backward()
    total = torch.nn.utils.clip_grad_norm_(vae.parameters(), max_norm=2.0)
    g_gpu = torch.cat([q.grad.reshape(-1) for q in vae.parameters()]).cpu().numpy()
    opt.step()
    popt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
    loss_ref, g_ref, total_ref = TP.train_step_port(port, popt, xb.cpu(), eps.cpu(), 0.2)
    assert abs(loss.item() - loss_ref[0]) <= LOSS_TOL * abs(loss_ref[0])
    scale = min(1.0, 2.0 / (total_ref + 1e-6))
    errs = _grad_errs(g_gpu, g_ref * scale, port)
    assert max(errs.values()) <= GRAD_TOL, errs
    assert abs(float(total) - total_ref) <= LOSS_TOL * total_ref
    # eval() afterwards scores with the UPDATED weights (the scorer handle is re-packed)
    vae.eval()
    with torch.no_grad():
        r2, _, _ = vae(xb)
    assert torch.isfinite(r2).all()


def test_trainer_three_steps_vs_port(cuda_dev):
    """VaeTrainer (flat buffers, fused clip+Adam) over three steps with the sigmoid KL schedule vs the port."""
    sd = synth.stage_vae_weights("4dof", seed=71)
    vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.3)
    vae.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    vae = vae.to(cuda_dev).train()
    tr = train.VaeTrainer(vae, 100, 48)
    port = TP.VaeTrainPort(sd).train()
    opt = torch.optim.Adam(port.ordered_parameters(), lr=1e-3, weight_decay=1e-5)
    rng = np.random.Generator(np.random.PCG64(72))
    for step in range(3):
        X, eps = synth.windows(48, 100, 12, seed=80 + step), synth.eps(48, 16, seed=90 + step)
        masks = tuple((rng.random((1, 48, 100, 128)) >= 0.3).astype(np.uint8) for _ in range(2))
        kl_w = train.kl_anneal_sigmoid(20 + step, 50)
        loss3 = tr.step(torch.from_numpy(X).to(cuda_dev), kl_w, eps=torch.from_numpy(eps).to(cuda_dev),
                        masks=tuple(torch.from_numpy(m).to(cuda_dev) for m in masks))
        ref3, _, total_ref = TP.train_step_port(port, opt, torch.from_numpy(X), torch.from_numpy(eps), kl_w,
                                                torch.from_numpy(masks[0]), torch.from_numpy(masks[1]), 0.3)
        assert np.allclose(loss3.cpu().numpy(), ref3, rtol=5e-5), (step, loss3, ref3)
        assert abs(float(tr.last_norm[1]) - total_ref) <= 1e-4 * total_ref
    # parameters are views of the flat buffer: the module sees the trained weights
    got = torch.cat([q.detach().reshape(-1) for q in vae.parameters()]).cpu().numpy()
    ref = TP.flat_params(port)
    assert np.mean(np.abs(got - ref) <= 2e-5) > 0.99 and np.max(np.abs(got - ref)) <= 3.1e-3
    assert abs(train.kl_anneal_sigmoid(1, 50) - 1.0 / (1.0 + np.exp(5.0))) < 1e-12
    tr.close()


def test_train_api_errors(cuda_dev):
    cfg = train._lib.VaeCfg(12, 128, 16, 2, 1, 1e-5, 0)
    h = train.VaeTrainHandle(cfg, 100, 8, cuda_dev)
    assert h.n_params == 477100
    flat = torch.zeros(h.n_params, device=cuda_dev)
    with pytest.raises(train.ShmfastError):                      # CPU tensors: no CPU fallback
        h.forward(flat, torch.zeros(8, 100, 12), torch.zeros(8, 16, device=cuda_dev))
    with pytest.raises(train.ShmfastError):                      # batch larger than the workspace
        h.forward(flat, torch.zeros(9, 100, 12, device=cuda_dev), torch.zeros(9, 16, device=cuda_dev))
    with pytest.raises(train.ShmfastError):                      # backward without a forward
        h.backward(flat, torch.zeros(8, 100, 12, device=cuda_dev), None, None)
    bad = train._lib.VaeCfg(12, 48, 16, 2, 1, 1e-5, 0)
    with pytest.raises(train.ShmfastError):
        train.VaeTrainHandle(bad, 100, 8, cuda_dev)
    h.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_two_gpus_matches_single(tmp_path):
    """Two ranks x half the batch (one NCCL all-reduce of the flat gradient) == one rank x the full batch."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    script = tmp_path / "ddp_worker.py"
    script.write_text(DDP_WORKER)
    env = dict(os.environ, SHM_PKG=str(root / "hybrid-vae-cnn-for-shm_b200"), SHM_ROOT=str(root), OUT=str(tmp_path))
    import socket
    for world in (1, 2):
        with socket.socket() as sk:                       # a free rendezvous port (concurrent test runs must not collide)
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
                            "127.0.0.1", "--master-port", str(port), str(script)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
    # the collective itself: the all-reduced mean gradient of step 0 equals the full-batch gradient to fp32 summation noise
    g1, g2 = np.load(tmp_path / "grad_w1.npy"), np.load(tmp_path / "grad_w2.npy")
    assert np.max(np.abs(g1 - g2)) <= 2e-5 * np.max(np.abs(g1)), (np.max(np.abs(g1 - g2)), np.max(np.abs(g1)))
    # parameters after two Adam steps: Adam divides by sqrt(v) ~ |g|, so entries whose gradient is summation noise move by
    # +-lr with a noise-determined sign -- the bound is 2 steps x lr (+ clipping slack), the bulk agrees to 2e-5
    a, b = np.load(tmp_path / "flat_w1.npy"), np.load(tmp_path / "flat_w2.npy")
    assert np.mean(np.abs(a - b) <= 2e-5) > 0.99 and np.max(np.abs(a - b)) <= 3.1e-3


DDP_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SHM_PKG"]); sys.path.insert(0, os.environ["SHM_ROOT"])
from shmfast import synth, train
from shmfast.models import fourdof
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", init_method="env://", device_id=dev)
sd = synth.stage_vae_weights("4dof", seed=5 + rank)            # different per rank: the initial broadcast must fix it
vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.0)
vae.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
vae = vae.to(dev).train()
tr = train.VaeTrainer(vae, 100, 64)
B = 64
for step in range(2):
    X, eps = synth.windows(B, 100, 12, seed=10 + step), synth.eps(B, 16, seed=20 + step)
    lo, hi = rank * B // world, (rank + 1) * B // world
    tr.step(torch.from_numpy(X[lo:hi]).to(dev), 0.5, eps=torch.from_numpy(eps[lo:hi]).to(dev))
    if step == 0 and rank == 0:
        np.save(os.path.join(os.environ["OUT"], f"grad_w{world}.npy"), (tr.grads / world).cpu().numpy())     # SUM all-reduce -> mean
torch.cuda.synchronize()
if rank == 0:
    np.save(os.path.join(os.environ["OUT"], f"flat_w{world}.npy"), tr.flat.cpu().numpy())
dist.barrier(); dist.destroy_process_group()
"""
