"""One launch of the tensor-core scorer on 148*128*8 windows (for `ncu --set full -k regex:vae_score_tc`)."""
import sys, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
stage = sys.argv[1] if len(sys.argv) > 1 else "4dof"
s = synth.STAGES[stage]
N = 148 * 128 * 8
vae = ops.VaeScorer(synth.stage_vae_weights(stage, seed=0), dev)
rows = (N - 1) * s["stride"] + s["T"]
series = torch.from_numpy(synth.series(rows, s["D"], seed=1)).to(dev)
src = ops.WindowSource(series, s["T"], stride=s["stride"])
eps = torch.randn((N, s["Z"]), device=dev)
for _ in range(2):
    vae.score(src, eps)
torch.cuda.synchronize()
print("done", N)
