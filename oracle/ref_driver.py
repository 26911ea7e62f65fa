"""Run the reference's OWN, unmodified files (TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product).

The files come from /root/reference when it exists (the build container) or from the byte-for-byte copies
`oracle/stage_ref.py` put into the git-ignored `oracle/_ref/` (what the GPU box sees).  Used by

* tests/test_gpu_reference_scripts.py, tests/test_oracle_golden.py: script-level drop-in checks -- the same script
  (`06_test_full_pipeline.main`, `04_vae_thresholding.full_mse_scores_batched`, `10_test_hybrid_pipeline.
  recon_mse_per_window / stage2_predict_cnn`) executed once with the reference `Models/` and once with `Models/`
  resolved to the three-line shmfast stubs of INTEGRATION.md;
* bench.py `--impl reference` / `cpu_baseline` (`kind: "reference"`) and the torch-CUDA incumbent legs.

A script run happens in a scratch tree (the scripts locate data relative to their own path and write next to
themselves); matplotlib is not installed in this image, so a stub stands in for it (plotting is not on the path).
"""
from __future__ import annotations

import importlib.util
import json
import shutil
import sys
import time
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
STAGED = HERE / "_ref"
LIVE = Path("/root/reference")

DIR4 = "4DOF"
DIROL = "20250506_openLAB_tests"
DIR1 = "1_DOF"


def ref_root() -> Path | None:
    """Where the unmodified reference files are, or None (then callers fall back to the port / skip)."""
    if (LIVE / DIR4 / "Scripts" / "Models" / "temporal_vae.py").exists():
        return LIVE
    if (STAGED / "MANIFEST.json").exists() and (STAGED / DIR4 / "Scripts" / "Models" / "temporal_vae.py").exists():
        return STAGED
    return None


def install_matplotlib_stub() -> None:
    """matplotlib is absent from this image and every entry script imports it at module top."""
    try:
        import matplotlib  # noqa: F401
        return
    except Exception:
        pass
    if "matplotlib" in sys.modules:
        return
    m = types.ModuleType("matplotlib")
    m.use = lambda *a, **k: None
    m.rcParams = {}
    plt = mock.MagicMock(name="matplotlib.pyplot")

    def subplots(nrows=1, ncols=1, *a, **k):
        n = int(nrows) * int(ncols)
        fig = mock.MagicMock(name="Figure")
        if n == 1:
            return fig, mock.MagicMock(name="Axes")
        axes = np.empty((n,), dtype=object)
        for i in range(n):
            axes[i] = mock.MagicMock(name=f"Axes{i}")
        return fig, (axes.reshape(nrows, ncols) if nrows > 1 and ncols > 1 else axes)

    plt.subplots = subplots
    m.pyplot = plt
    for name in ("lines", "colors", "cm", "ticker", "patches", "gridspec"):
        sub = mock.MagicMock(name=f"matplotlib.{name}")
        setattr(m, name, sub)
        sys.modules[f"matplotlib.{name}"] = sub
    sys.modules["matplotlib"] = m
    sys.modules["matplotlib.pyplot"] = plt


def load_by_path(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_models(stage: str):
    """The reference's model classes, loaded by file path: ('4dof' -> TemporalVAE, CNN), ('openlab' -> VAE, CNN),
    ('1dof' -> TemporalVAE, None)."""
    root = ref_root()
    if root is None:
        raise FileNotFoundError("no reference files: neither /root/reference nor oracle/_ref (run oracle/stage_ref.py in the build container)")
    if stage == "4dof":
        v = load_by_path(root / DIR4 / "Scripts/Models/temporal_vae.py", "shmref_vae4")
        c = load_by_path(root / DIR4 / "Scripts/Models/cnn_model.py", "shmref_cnn4")
        return v.TemporalVAE, c.CNN
    if stage == "openlab":
        v = load_by_path(root / DIROL / "Codes/Models/temporal_vae_model.py", "shmref_vaeol")
        c = load_by_path(root / DIROL / "Codes/Models/cnn_model.py", "shmref_cnnol")
        return v.VAE, c.CNN
    if stage == "1dof":
        v = load_by_path(root / DIR1 / "Scripts/Models/temporal_vae.py", "shmref_vae1")
        return v.TemporalVAE, None
    raise ValueError(stage)


def _t_sd(sd: dict) -> dict:
    return {k: (v.detach().clone() if isinstance(v, torch.Tensor) else torch.from_numpy(np.array(v))) for k, v in sd.items()}


STUBS_4DOF = {
    "temporal_vae.py": "from shmfast.models.fourdof import TemporalVAE, VAE\n__all__ = ['TemporalVAE', 'VAE']\n",
    "cnn_model.py": "from shmfast.models.fourdof import CNN, CNNClassifier, SEQ_LEN, NUM_FEATURES\n",
    "__init__.py": "",
}
STUBS_OL = {
    "temporal_vae_model.py": "from shmfast.models.openlab import VAE\n",
    "cnn_model.py": "from shmfast.models.openlab import CNN, SEQ_LEN, NUM_FEATURES\n",
    "__init__.py": "",
}


def _purge(prefixes) -> None:
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in prefixes):
            del sys.modules[k]


class Tree4dof:
    """Scratch copy of the 4DOF stage: Scripts/ (unmodified scripts; Models/ = the reference's or the shmfast stubs),
    Data/processed/{run_splits.json, normal_stats.npz, vae_threshold.json}, models/*.pt."""

    def __init__(self, tmp: Path, models: str, vae_sd: dict, cnn_sd: dict, mean, std, thr: float, splits: dict | None = None,
                 link_raw: bool = True):
        root = ref_root()
        if root is None:
            raise FileNotFoundError("no reference files staged")
        self.root = Path(tmp)
        src = root / DIR4
        (self.root / "Scripts" / "Models").mkdir(parents=True, exist_ok=True)
        for f in ("__init__.py", "03_train_vae.py", "04_vae_thresholding.py", "05_train_cnn.py", "06_test_full_pipeline.py"):
            shutil.copyfile(src / "Scripts" / f, self.root / "Scripts" / f)
        if models == "reference":
            for f in ("__init__.py", "temporal_vae.py", "cnn_model.py"):
                shutil.copyfile(src / "Scripts/Models" / f, self.root / "Scripts/Models" / f)
        elif models == "shmfast":
            for f, body in STUBS_4DOF.items():
                (self.root / "Scripts/Models" / f).write_text(body)
        else:
            raise ValueError(models)
        proc = self.root / "Data" / "processed"
        proc.mkdir(parents=True, exist_ok=True)
        if splits is None:
            shutil.copyfile(src / "Data/processed/run_splits.json", proc / "run_splits.json")
            if link_raw and not (self.root / "Data" / "raw").exists():
                (self.root / "Data" / "raw").symlink_to(src / "Data" / "raw")
        else:
            (proc / "run_splits.json").write_text(json.dumps(splits))
        np.savez(proc / "normal_stats.npz", mean=np.asarray(mean, np.float32), std=np.asarray(std, np.float32))
        (proc / "vae_threshold.json").write_text(json.dumps({"threshold": float(thr), "score_def": "full_window_mse"}))
        (self.root / "models").mkdir(exist_ok=True)
        torch.save(_t_sd(vae_sd), str(self.root / "models" / "temporal_vae_state_dict.pt"))
        torch.save(_t_sd(cnn_sd), str(self.root / "models" / "cnn_state_dict.pt"))

    def load(self, script: str):
        """Import Scripts/<script>.py of this tree as a fresh module (its `Scripts.Models` imports resolve inside the tree)."""
        install_matplotlib_stub()
        _purge(["Scripts"])
        sys.path.insert(0, str(self.root))
        try:
            return load_by_path(self.root / "Scripts" / f"{script}.py", f"shmref_{script}_{abs(hash(str(self.root))) % 10**8}")
        finally:
            if str(self.root) in sys.path:
                sys.path.remove(str(self.root))


def run_06_main(tree: Tree4dof, csv_override: dict | None = None) -> dict:
    """Execute the unmodified `06_test_full_pipeline.main()` in `tree`; plotting is replaced by spies that record the
    per-window gate scores / hybrid scores main() hands to them.  `csv_override` {file name: [R,12] array} replaces
    `load_csv_numeric` (bench: synthetic series instead of CSV parsing).  Returns scores, metrics and the wall time of
    the three eval_group calls (first window load -> first metric call)."""
    mod = tree.load("06_test_full_pipeline")
    cap: dict = {}
    stamps: dict = {}

    def spy_roc(y_gate, s_gate, y_hyb, s_hyb, stem):
        cap.update(gate_labels=np.asarray(y_gate).copy(), gate_scores=np.asarray(s_gate).copy(),
                   hyb_labels=np.asarray(y_hyb).copy(), hyb_scores=np.asarray(s_hyb).copy())
        return {}

    mod.plot_roc_two = spy_roc
    mod.plot_pr_curve = lambda *a, **k: {}
    mod.plot_cm_row_norm = lambda *a, **k: None
    real_load = mod.load_csv_numeric

    def load_csv(path):
        stamps.setdefault("t0", time.perf_counter())
        if csv_override is not None:
            return np.asarray(csv_override[Path(path).name], dtype=np.float32)
        return real_load(path)

    mod.load_csv_numeric = load_csv
    real_acc = mod.accuracy_score

    def acc(*a, **k):
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        stamps.setdefault("t1", time.perf_counter())
        return real_acc(*a, **k)

    mod.accuracy_score = acc
    mod.main()
    metrics = json.loads((tree.root / "Output" / "figures" / "pipeline_metrics.json").read_text())
    cap.update(metrics=metrics, eval_seconds=stamps["t1"] - stamps["t0"])
    return cap


class TreeOpenLab:
    """Scratch copy of openLAB Codes/ (10_test_hybrid_pipeline.py, 05_validate_vae.py, config.py; Models/ = reference or stubs)."""

    def __init__(self, tmp: Path, models: str):
        root = ref_root()
        if root is None:
            raise FileNotFoundError("no reference files staged")
        self.root = Path(tmp)
        src = root / DIROL / "Codes"
        (self.root / "Codes" / "Models").mkdir(parents=True, exist_ok=True)
        for f in ("__init__.py", "config.py", "05_validate_vae.py", "10_test_hybrid_pipeline.py", "04_train_vae.py", "06_train_cnn.py"):
            shutil.copyfile(src / f, self.root / "Codes" / f)
        if models == "reference":
            for f in ("temporal_vae_model.py", "cnn_model.py"):
                shutil.copyfile(src / "Models" / f, self.root / "Codes/Models" / f)
        elif models == "shmfast":
            for f, body in STUBS_OL.items():
                (self.root / "Codes/Models" / f).write_text(body)
        else:
            raise ValueError(models)

    def load(self, script: str):
        install_matplotlib_stub()
        _purge(["Models", "config"])
        codes = str(self.root / "Codes")
        sys.path.insert(0, codes)
        try:
            return load_by_path(self.root / "Codes" / f"{script}.py", f"shmref_ol_{script}_{abs(hash(codes)) % 10**8}")
        finally:
            # stage2_predict_cnn imports Models.cnn_model lazily: keep the tree importable until the caller is done
            self._codes = codes

    def release(self):
        codes = getattr(self, "_codes", None)
        if codes and codes in sys.path:
            sys.path.remove(codes)
        _purge(["Models", "config"])


def cnn_artifacts_openlab(tmp: Path, cnn_sd: dict, mu, sd, thr: float) -> dict:
    """The artefact files stage2_predict_cnn reads (10_test_hybrid_pipeline.py:269-271,285)."""
    tmp = Path(tmp)
    tmp.mkdir(parents=True, exist_ok=True)
    np.save(tmp / "cnn_raw_mu_sd.npy", np.stack([np.asarray(mu, np.float32), np.asarray(sd, np.float32)]))
    np.save(tmp / "cnn_best_threshold.npy", np.asarray([thr], np.float64))
    torch.save(_t_sd(cnn_sd), str(tmp / "cnn_state_dict.pt"))
    return {"norm_stats": str(tmp / "cnn_raw_mu_sd.npy"), "thr": str(tmp / "cnn_best_threshold.npy"), "model": str(tmp / "cnn_state_dict.pt")}


# ---------------------------------------------------------------------------------------------------------
# bench.py legs
# ---------------------------------------------------------------------------------------------------------
def rows_for_windows(n: int, T: int = 100, frac=(0.7, 1.0)) -> int:
    """Smallest series length whose slice_frac(frac) part yields >= n windows of length T at stride 1
    (06_test_full_pipeline.py:98-110)."""
    r = int((n + T - 1) / (frac[1] - frac[0]))
    while int(r * frac[1]) - int(r * frac[0]) < n + T - 1:
        r += 1
    return r


def bench_4dof_reference(tmp: Path, vae_sd: dict, cnn_sd: dict, mean, std, thr: float, series3, passes: int, warmup: int) -> dict:
    """`06_test_full_pipeline.main()` -- the reference's own eval_group code, unmodified -- on three synthetic groups
    (normal / sensor / structural), CSV parsing replaced by in-memory series, plots by no-ops; CPU, all host threads.
    Times the three eval_group calls (windowing + normalisation + score loop + threshold + second pass + CNN)."""
    splits = {"normal": {"files": ["g0.csv"]}, "sensor_fault": {"files": ["g1.csv"]}, "structural_fault": {"files": ["g2.csv"]}}
    tree = Tree4dof(tmp, "reference", vae_sd, cnn_sd, mean, std, thr, splits=splits)
    override = {f"g{i}.csv": s for i, s in enumerate(series3)}
    times, last = [], None
    import contextlib
    import io
    # the CPU arm: main() picks its device itself; its progress prints must not reach bench.py's one-JSON-line stdout
    with mock.patch.object(torch.cuda, "is_available", lambda: False), contextlib.redirect_stdout(io.StringIO()):
        for i in range(warmup + passes):
            last = run_06_main(tree, override)
            if i >= warmup:
                times.append(last["eval_seconds"])
    n = int(last["gate_scores"].shape[0])
    flagged = int(sum(v["anom"] for v in last["metrics"]["gate"]["gate_stats"].values()))
    return dict(seconds=sum(times) / len(times), windows=n, flagged=flagged)
