import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
import numpy as np, torch
from shmfast import cnn_train as CT, synth
from shmfast.models import fourdof, openlab
dev = torch.device("cuda", 0)
arch = sys.argv[1] if len(sys.argv) > 1 else "openlab"
if arch == "4dof":
    B = 100
    model = fourdof.CNN(2, 2, 0.5); sd = synth.cnn4dof_weights(seed=0)
    x = np.stack([synth.windows(B, 100, 12, seed=1), synth.windows(B, 100, 12, seed=2) ** 2], axis=1).astype(np.float32)
    alpha = None
else:
    B = 128
    model = openlab.CNN(dropout_rate=0.4); sd = synth.cnnol_weights(seed=0)
    x = np.clip(2.0 * synth.windows(B, 200, 4, seed=1), -10, 10).astype(np.float32)[:, None]
    alpha = torch.tensor([0.8, 1.2])
model.load_state_dict({k: torch.from_numpy(np.array(v)) for k, v in sd.items()})
model = model.to(dev).train()
tr = CT.CnnTrainer(model, B, alpha=alpha)
xd, yd = torch.from_numpy(x).to(dev), torch.randint(0, 2, (B,), device=dev)
for _ in range(3):
    tr.step(xd, yd)
torch.cuda.synchronize()
if len(sys.argv) > 2 and sys.argv[2] == "time":          # python scripts/prof_cnn_train.py openlab time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        tr.step(xd, yd)
    e1.record()
    torch.cuda.synchronize()
    print(f"{arch} cnn train step: {e0.elapsed_time(e1) / 50:.3f} ms (batch {B})")
