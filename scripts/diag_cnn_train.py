"""Diagnostic: per-tensor gradient errors of the CNN training kernels vs the fp64 autograd port, several batch sizes."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200"), str(ROOT / "tests")]
import numpy as np, torch
from oracle import torch_port as TP
from shmfast import _lib, cnn_train as CT, synth
from test_gpu_cnn_train import problem
dev = torch.device("cuda", 0)
for arch, Bs in (("4dof", (7, 24, 50, 100)), ("openlab", (3, 128))):
    for B in Bs:
        sd, x = problem(arch, B, seed=60 + B)
        y = np.random.Generator(np.random.PCG64(B)).integers(0, 2, size=B).astype(np.int64)
        port = TP.CnnTrainPort(arch, sd).double()
        if arch == "4dof":
            port.running = [r.double() for r in port.running]
        opt = torch.optim.SGD(port.ordered_parameters(), lr=0.0)
        kw = {} if arch == "4dof" else dict(alpha=torch.tensor([0.8, 1.2], dtype=torch.float64), gamma=2.0)
        lg, loss, g64, _ = TP.cnn_train_step_port(port, opt, torch.from_numpy(x).double(), torch.from_numpy(y), None, 0.0, **kw)
        params = torch.cat([q.detach().reshape(-1) for q in port.ordered_parameters()]).float().to(dev)
        h = CT.CnnTrainHandle(_lib.CNN_4DOF if arch == "4dof" else _lib.CNN_OPENLAB, 128, dev)
        xd = torch.from_numpy(x).to(dev)
        for rep in range(2):
            logits = h.forward(params, xd)
            l, dl = CT.cnn_loss_grad(logits, torch.from_numpy(y).to(dev), None if arch == "4dof" else torch.tensor([0.8, 1.2]), kw.get("gamma", 0.0))
            got = h.backward(params, dl).cpu().numpy()
            o = 0
            line = []
            for n, q in zip(port.names, port.ordered_parameters()):
                k = q.numel()
                e = np.abs(got[o:o + k] - g64[o:o + k]).max() / max(np.abs(g64[o:o + k]).max(), 1e-12)
                line.append(f"{n.split('.')[0][-5:]}.{n.split('.')[-2]}.{n.split('.')[-1][0]}={e:.1e}")
                o += k
            print(arch, B, "rep", rep, "logit err", float(np.abs(logits.cpu().numpy() - lg).max()), " ".join(line))
        h.close()
