"""Exact percentile on 2^27 (and 2^16: launch-overhead view) scores, CUDA-event median of 9 with L2 flushed; `ncu` launch list source.
"""
import sys, json, os
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
import numpy as np, torch
from shmfast import ops
dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for logn in (27, 16):
    score = torch.randn(1 << logn, device=dev) ** 2
    for _ in range(3):
        r = ops.percentile(score, 99.0)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = ops.percentile(score, 99.0); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(json.dumps({"log2_n": logn, "ms": ms, "gbs": (4 << logn) / ms / 1e6, "p99": float(r.item())}))
