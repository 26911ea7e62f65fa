import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
import numpy as np, torch
from shmfast import synth, train
from shmfast.models import fourdof
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vae = fourdof.TemporalVAE(12, 16, 128, 2, dropout=0.3)
vae.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in synth.stage_vae_weights("4dof", seed=0).items()})
vae = vae.to(dev).train()
tr = train.VaeTrainer(vae, 100, B)
x = torch.from_numpy(synth.windows(B, 100, 12, seed=1)).to(dev)
for _ in range(3):
    tr.step(x, 0.5)
torch.cuda.synchronize()
