"""One tensor-core openLAB CNN forward over 8192 windows (for ncu captures of the conv GEMM / staging kernels)."""
import sys, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=0), dev)
x = torch.from_numpy(synth.windows(N, 200, 4, seed=1, amp=1.5)).to(dev)
src = ops.WindowSource(x, 200)
for _ in range(2):
    cnn.forward(src)
torch.cuda.synchronize()
print("done", N, "engine", cnn.engine)
