import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
import torch
from shmfast import ops
dev = torch.device("cuda", 0)
score = torch.rand(1 << 27, device=dev)
for _ in range(3):
    r = ops.percentile(score, 99.0)
torch.cuda.synchronize()
print(float(r.item()))
