#!/usr/bin/env python3
"""Stage the UNMODIFIED reference files of the hot path into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/stage_ref.py            # also called by __graft_entry__.build() when /root/reference is present

The reference is a Python program: there is nothing to compile, so the "build" of the real reference is a
byte-for-byte copy of the few files the path lives in (Models/, the numbered scripts whose loops the path
replaces, the small checked-in data they read).  oracle/_ref/ is git-ignored -- no reference source enters the
repository's history -- but NOT gpurun-ignored, so the copies travel to the GPU box like a built .so would.
Users: tests/test_gpu_reference_scripts.py (script-level drop-in checks), bench.py --impl reference and the
torch-CUDA incumbent legs (oracle/ref_driver.py).  Nothing in the product imports this.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")
DST = HERE / "_ref"

FILES = [
    # 4DOF: models, the four scripts whose inner loops are the hot path, the checked-in split / threshold / CSVs
    "4DOF/Scripts/__init__.py",
    "4DOF/Scripts/Models/__init__.py",
    "4DOF/Scripts/Models/temporal_vae.py",
    "4DOF/Scripts/Models/cnn_model.py",
    "4DOF/Scripts/03_train_vae.py",
    "4DOF/Scripts/04_vae_thresholding.py",
    "4DOF/Scripts/05_train_cnn.py",
    "4DOF/Scripts/06_test_full_pipeline.py",
    "4DOF/Data/processed/run_splits.json",
    "4DOF/Data/processed/vae_threshold.json",
    # openLAB
    "20250506_openLAB_tests/Codes/__init__.py",
    "20250506_openLAB_tests/Codes/config.py",
    "20250506_openLAB_tests/Codes/Models/temporal_vae_model.py",
    "20250506_openLAB_tests/Codes/Models/cnn_model.py",
    "20250506_openLAB_tests/Codes/04_train_vae.py",
    "20250506_openLAB_tests/Codes/05_validate_vae.py",
    "20250506_openLAB_tests/Codes/06_train_cnn.py",
    "20250506_openLAB_tests/Codes/10_test_hybrid_pipeline.py",
    # 1_DOF
    "1_DOF/Scripts/__init__.py",
    "1_DOF/Scripts/Models/__init__.py",
    "1_DOF/Scripts/Models/temporal_vae.py",
    "1_DOF/Scripts/datasets.py",
]
GLOBS = ["4DOF/Data/raw/**/*.csv"]


def stage(verbose: bool = False) -> Path | None:
    """Copy the listed files (idempotent).  Returns oracle/_ref, or None when /root/reference is absent
    (the GPU box: the tree staged in the build container travelled with the snapshot)."""
    if not REF.exists():
        return DST if (DST / "MANIFEST.json").exists() else None
    rels = list(FILES)
    for g in GLOBS:
        rels += sorted(str(p.relative_to(REF)) for p in REF.glob(g))
    manifest = {}
    for rel in rels:
        src, dst = REF / rel, DST / rel
        if not src.exists():
            raise FileNotFoundError(f"reference file missing: {src}")
        data = src.read_bytes()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if dst.exists() and dst.read_bytes() == data:
            continue
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        if verbose:
            print("staged", rel)
    (DST / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    return DST


if __name__ == "__main__":
    out = stage(verbose="-v" in sys.argv)
    print(out if out else "no /root/reference and nothing staged")
