"""Diagnostic: unmodified 06 main() on cuda with reference Models (cuDNN, TF32 on/off) vs the shmfast stubs."""
import sys, tempfile, io, contextlib
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "hybrid-vae-cnn-for-shm_b200")]
import numpy as np, torch
from oracle import ref_driver as R
g = np.load(ROOT / "tests/golden/trained_4dof.npz")
vae_sd = {k[4:]: g[k] for k in g.files if k.startswith("vae.")}
cnn_sd = {k[4:]: g[k] for k in g.files if k.startswith("cnn.")}
thr = float(g["thr"])
outs = {}
with tempfile.TemporaryDirectory() as td:
    for name, kind, tf32 in (("ref_tf32", "reference", True), ("ref_fp32", "reference", False), ("shm", "shmfast", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        tree = R.Tree4dof(Path(td) / name, kind, vae_sd, cnn_sd, g["mean"], g["std"], thr)
        with contextlib.redirect_stdout(io.StringIO()):
            outs[name] = R.run_06_main(tree)
        print(name, "eval s", round(outs[name]["eval_seconds"], 3), "cm", outs[name]["metrics"]["confusion_matrix_counts"])
def rel(a, b): return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)))
for x, y in (("shm", "ref_fp32"), ("shm", "ref_tf32"), ("ref_tf32", "ref_fp32")):
    a, b = outs[x], outs[y]
    print(x, "vs", y, "score rel", rel(a["gate_scores"], b["gate_scores"]), "hyb max abs", float(np.max(np.abs(a["hyb_scores"] - b["hyb_scores"]))),
          "flag diff", int(((a["gate_scores"] > thr) != (b["gate_scores"] > thr)).sum()))
