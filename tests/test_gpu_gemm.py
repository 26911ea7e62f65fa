"""The fp32-grade contraction of the training steps (shm_gemm_f32; csrc/gemm_tc.cu on tcgen05, csrc/train.cu on the FMA pipe):
the three operand layouts of the LSTM-VAE step -- input projection G = X W^T (both K-contiguous), dX = dG W (B N-contiguous),
dW = dG^T H (A M-contiguous, B N-contiguous, split-K with atomics) -- against a float64 product.
Tolerances: FMA pipe 2e-6 and fp16 x 3 split 5e-6 of the largest |C| (fp16 split: O(1) operands), bf16 x 3 split 4e-5."""
import numpy as np
import pytest
import torch

from shmfast import _lib, ops

pytestmark = pytest.mark.gpu

MODES = [(_lib.GEMM_SIMT, 2e-6), (_lib.GEMM_TC_F16X3, 5e-6), (_lib.GEMM_TC_BF16X3, 4e-5)]


def _check(C, ref, tol, what):
    err = float(np.max(np.abs(C.cpu().numpy().astype(np.float64) - ref)))
    assert err <= tol * float(np.max(np.abs(ref))), f"{what}: max err {err:.3e} vs max|C| {float(np.max(np.abs(ref))):.3e}"


@pytest.mark.parametrize("mode,tol", MODES)
@pytest.mark.parametrize("M,N,K", [(25600, 512, 128), (25600, 512, 12), (1000, 96, 100), (257, 33 * 4, 36)])
def test_projection_layout(cuda_dev, mode, tol, M, N, K):
    """C = X[M,K] W[N,K]^T + b: A and B K-contiguous (03_train_vae forward: nn.LSTM input projections over all timesteps)."""
    g = torch.Generator().manual_seed(M + N + K)
    X, W, b = torch.randn((M, K), generator=g), torch.randn((N, K), generator=g) / np.sqrt(K), torch.randn((N,), generator=g)
    ref = X.double().numpy() @ W.double().numpy().T + b.double().numpy()
    C = ops.gemm_f32(X.to(cuda_dev), K, 1, W.to(cuda_dev), 1, K, M, N, K, bias=b.to(cuda_dev), mode=mode)
    _check(C, ref, tol, f"projection {M}x{N}x{K} mode {mode}")


@pytest.mark.parametrize("mode,tol", MODES)
@pytest.mark.parametrize("M,N,K", [(25600, 128, 512), (3000, 128, 12), (644, 64, 48)])
def test_dx_layout(cuda_dev, mode, tol, M, N, K):
    """dX = dG[M,K] W[K,N]: A K-contiguous, B N-contiguous; gradient-sized magnitudes (1e-4) exercise the bf16 range."""
    g = torch.Generator().manual_seed(M + N + K + 1)
    dG, W = 1e-4 * torch.randn((M, K), generator=g), torch.randn((K, N), generator=g) / np.sqrt(K)
    scale = 1.0 if mode != _lib.GEMM_TC_F16X3 else 1e4          # the fp16 split is for O(1) operands (forward activations)
    ref = (scale * dG.double().numpy()) @ W.double().numpy()
    C = ops.gemm_f32((scale * dG).to(cuda_dev), K, 1, W.to(cuda_dev), N, 1, M, N, K, mode=mode)
    _check(C, ref, tol, f"dX {M}x{N}x{K} mode {mode}")


@pytest.mark.parametrize("mode,tol", MODES)
@pytest.mark.parametrize("M,N,K", [(512, 128, 25600), (512, 128, 1000), (128, 32, 7777 * 4)])
def test_dw_layout_splitk(cuda_dev, mode, tol, M, N, K):
    """dW = dG[K,M]^T H[K,N]: A M-contiguous, B N-contiguous, split-K with the atomic epilogue."""
    g = torch.Generator().manual_seed(M + N + K + 2)
    dG, Hm = torch.randn((K, M), generator=g), torch.randn((K, N), generator=g)
    if mode != _lib.GEMM_TC_F16X3:
        dG = 1e-4 * dG
    ref = dG.double().numpy().T @ Hm.double().numpy()
    C = ops.gemm_f32(dG.to(cuda_dev), 1, M, Hm.to(cuda_dev), N, 1, M, N, K, splitk=True, mode=mode)
    # a K-long sum of random-sign terms: the error budget scales with sqrt(K) * |term|, the result with sqrt(K) too
    _check(C, ref, 3 * tol, f"dW {M}x{N}x{K} mode {mode}")


def test_tc_and_simt_paths_agree_on_the_train_step(cuda_dev):
    """shm_train_set_tensor_cores(0/1): the same forward + backward on the FMA pipe and on tcgen05 (every gradient tensor within
    1e-4 of its maximum)."""
    from shmfast import synth, train
    lib = _lib.load()
    cfg = _lib.VaeCfg(12, 128, 16, 2, 1, 1e-5, 0)
    B, T = 64, 100
    h = train.VaeTrainHandle(cfg, T, B, cuda_dev)
    sd = synth.stage_vae_weights("4dof", seed=3)
    from oracle import torch_port as TP
    names = TP.vae_param_names(sd)
    flat = torch.cat([torch.from_numpy(np.asarray(sd[n]).reshape(-1)) for n in names]).to(cuda_dev)
    x = torch.from_numpy(synth.windows(B, T, 12, seed=3)).to(cuda_dev)
    eps = torch.from_numpy(synth.eps(B, 16, seed=3)).to(cuda_dev)
    res = {}
    try:
        for tcn in (0, 1):
            lib.shm_train_set_tensor_cores(tcn)
            xhat, mu, lv = h.forward(flat, x, eps)
            l3, dx, dm, dl = train.elbo_grad(x, xhat, mu, lv, 0.5)
            res[tcn] = (xhat.clone(), float(l3[0].item()), h.backward(flat, dx, dm, dl).clone())
    finally:
        lib.shm_train_set_tensor_cores(1)
    assert abs(res[1][1] - res[0][1]) <= 1e-5 * abs(res[0][1])
    assert float((res[1][0] - res[0][0]).abs().max()) <= 1e-4 * float(res[0][0].abs().max())
    g0, g1 = res[0][2].cpu().numpy(), res[1][2].cpu().numpy()
    o = 0
    for n in names:
        k = int(np.asarray(sd[n]).size)
        a, b = g1[o:o + k], g0[o:o + k]
        o += k
        assert float(np.max(np.abs(a - b))) <= 1e-4 * float(np.max(np.abs(b))) + 1e-9, n
    h.close()
