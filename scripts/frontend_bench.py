"""Throughput of the device extraction front-end (shm_openlab_extract) on a long synthetic run, with the NumPy oracle
(the reference's per-run numerics, loop-free) timed beside it on a bounded sample.  One JSON line."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
from shmfast import openlab_frontend as FE
from oracle import np_oracle as O

dev = torch.device("cuda", 0)
R = 1 << 24
rng = np.random.Generator(np.random.PCG64(3))
raw = (30 + np.cumsum(0.05 * rng.standard_normal((R, 4), dtype=np.float32), axis=0)).astype(np.float32)
raw[rng.integers(0, R, 2000), 0] = np.nan
raw[rng.integers(0, R, 200), 3] = -2e5
d = torch.from_numpy(raw).to(dev)
FE.extract_run(d); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run = FE.extract_run(d); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
# algorithmic bytes: raw rows read once (16 B), clean + raw rows written (32 B), 3 row masks written + read by the windows (24 B)
alg = R * (16 + 32)
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
n_cpu = 1 << 20
t0 = time.perf_counter(); O.openlab_extract_run(raw[:n_cpu]); cpu_s = time.perf_counter() - t0
print(json.dumps({"kernel": "shm_openlab_extract (4 launches + compaction)", "rows": R, "windows": run.n_windows, "ms": ms,
                  "rows_per_s": R / ms * 1e3, "windows_per_s": run.n_windows / ms * 1e3, "algorithmic_bytes": alg,
                  "gbs": alg / ms / 1e6, "peak_gbs": peak, "frac": alg / ms / 1e6 / peak,
                  "cpu_oracle_rows_per_s": n_cpu / cpu_s, "cpu_sample_rows": n_cpu}))
