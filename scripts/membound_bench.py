"""HBM-bound kernels of the path against the measured copy bandwidth (MEASURED_PEAKS.json: hbm_gbs):
window gather + normalise (materialising), threshold + compaction, percentile, stitch.  CUDA events, L2 flushed between reps.
Usage: python scripts/membound_bench.py  -> one JSON line per kernel."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "hybrid-vae-cnn-for-shm_b200"))
from shmfast import ops, synth
from shmfast.pipeline import guard_std_4dof

dev = torch.device("cuda", 0)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

out = []
# 1. gather + normalise, materialised [N,100,12] from a stride-1 series (4DOF make_windows + normalize_windows)
N, T, D = 1 << 20, 100, 12
series = torch.from_numpy(synth.series(N + T - 1, D, seed=1)).to(dev)
mean, std = synth.stats(D, seed=0)
src = ops.WindowSource(series, T, stride=1, mean=mean, std=guard_std_4dof(std), nan_to_zero=True)
buf = torch.empty((N, T, D), dtype=torch.float32, device=dev)
ms = timed(lambda: ops.window_normalize(src, out=buf))
b = N * T * D * 4 + series.numel() * 4        # bytes written + unique bytes read
out.append(dict(kernel="window_normalize_kernel (4DOF, stride 1, materialising)", ms=ms, algorithmic_bytes=b, gbs=b / ms / 1e6))
# openLAB shape: T=200 stride 20, 4 channels, clip + NaN policy
N2 = 1 << 21
ser2 = torch.from_numpy(synth.series((N2 - 1) * 20 + 200, 4, seed=2, nan_frac=0.0007)).to(dev)
mu2, sd2 = synth.stats(4, seed=2)
src2 = ops.WindowSource(ser2, 200, stride=20, mean=mu2, std=sd2, clip=10.0, nan_to_zero=True)
buf2 = torch.empty((N2, 200, 4), dtype=torch.float32, device=dev)
ms = timed(lambda: ops.window_normalize(src2, out=buf2))
b = N2 * 200 * 4 * 4 + ser2.numel() * 4
out.append(dict(kernel="window_normalize_kernel (openLAB, stride 20)", ms=ms, algorithmic_bytes=b, gbs=b / ms / 1e6))
del buf, buf2
# 2. threshold + compaction on 2^27 scores at 1 % and 47 % flag rates
M = 1 << 27
score = torch.rand(M, device=dev)
for frac in (0.01, 0.47):
    thr = 1.0 - frac
    ms = timed(lambda: ops.compact(score, thr))
    b = M * 4 + M * 1 + int(frac * M) * 4
    out.append(dict(kernel=f"compact_kernel ({frac:.0%} flagged)", ms=ms, algorithmic_bytes=b, gbs=b / ms / 1e6))
# 3. exact percentile (radix select) on 2^27 scores: algorithmic = one read of the scores
ms = timed(lambda: ops.percentile(score, 99.0))
out.append(dict(kernel="percentile (one-read exact select: sample bracket + filter + radix select of the candidates)", ms=ms, algorithmic_bytes=M * 4, gbs=M * 4 / ms / 1e6))
for o in out:
    o["peak_gbs"] = peak; o["frac"] = o["gbs"] / peak
    print(json.dumps(o))
