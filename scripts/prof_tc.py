import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/hybrid-vae-cnn-for-shm_b200")
from shmfast import ops, synth
dev = torch.device("cuda", 0)
N = 148 * 128 * 4
vae = ops.VaeScorer(synth.stage_vae_weights("4dof", seed=0), dev)
series = torch.from_numpy(synth.series(N + 99, 12, seed=1)).to(dev)
src = ops.WindowSource(series, 100, stride=1)
eps = torch.randn((N, 16), device=dev)
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
print("MMA issuer : wait weights %.2f  input %.2f  acc drain %.2f  h %.2f | pass totals %s" % (m[0,0], m[0,1], m[0,2], m[0,3], np.round(m[0,4:], 2)))
print("aux warp 8 : wait (in_empty / xhat_full) %.2f" % (m[1,0],))
print("epilogue w0: wait acc_full c0,c1,c2+3, tmem-ld %s | pass totals %s" % (np.round(m[2,:4], 2), np.round(m[2,4:], 2)))
