"""Time the tensor-core openLAB CNN forward over N windows (CUDA events, mean of 10 after 3 warm-ups)."""
import sys, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
cnn = ops.CnnOpenLab(synth.cnnol_weights(seed=0), dev)
x = torch.from_numpy(synth.windows(N, 200, 4, seed=1, amp=1.5)).to(dev)
src = ops.WindowSource(x, 200)
for _ in range(3):
    out = cnn.forward(src)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = cnn.forward(src)
e1.record()
torch.cuda.synchronize()
lg = out[0] if isinstance(out, (tuple, list)) else out
print(f"N={N} engine={cnn.engine} ms={e0.elapsed_time(e1) / 10:.3f} checksum={float(lg.double().abs().sum()):.6f}")
