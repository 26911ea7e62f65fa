import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
N = 148 * 128 * 8
vae = ops.VaeScorer(synth.stage_vae_weights("openlab", seed=0), dev)
series = torch.from_numpy(synth.series((N - 1) * 20 + 200, 3, seed=1)).to(dev)
src = ops.WindowSource(series, 200, stride=20)
eps = torch.randn((N, 8), device=dev)
vae.score(src, eps); torch.cuda.synchronize()
vae.debug_counters()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); vae.score(src, eps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c = vae.debug_counters()
print(f"N={N} {ms:.2f} ms -> {N/ms*1e3/1e6:.3f} M windows/s")
if not c.any():
    print("role counters are compiled out (build with SHMFAST_PROF=1 python -m shmfast.build --force to get them)"); sys.exit(0)
m = c.mean(0) / 1e6
print("epilogue w0 (M cycles): wait acc_full %.2f  tmem-ld %.2f  cells %.2f  split+st+arrive %.2f | enc pass %.2f  dec pass %.2f" % tuple(m[2, :6]))
print("issuer tile A waits (M cycles) enc: weights %.2f  h_full %.2f  acc_empty/in_full %.2f  - %.2f | dec: weights %.2f  h_full %.2f  acc_empty %.2f  - %.2f" % tuple(m[1,:8]))
