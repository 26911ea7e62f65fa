"""One 4DOF CNN forward over N residual-stack windows (for ncu captures of cnn4dof_conv_kernel / cnn4dof_fc_kernel)."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cnn = ops.Cnn4dof(synth.cnn4dof_weights(seed=0), dev)
x = torch.randn((N, 2, 100, 12), device=dev)
for _ in range(2):
    out = cnn.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = cnn.forward(x); e1.record(); torch.cuda.synchronize()
print(f"N={N} {e0.elapsed_time(e1):.3f} ms")
