"""One launch of the threshold + compaction kernel on 2^27 scores (for `ncu --set full -k regex:compact_kernel`)."""
import sys, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops
dev = torch.device("cuda", 0)
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
score = torch.rand(1 << 27, device=dev)
for _ in range(3):
    ops.compact(score, 1.0 - frac)
torch.cuda.synchronize()
print("done")
