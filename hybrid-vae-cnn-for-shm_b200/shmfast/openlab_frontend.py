"""openLAB extraction front-end: catman MD_*.txt -> device-resident cleaned/raw series,
window metadata and rule labels, i.e. 20250506_openLAB_tests/Codes/01_extract_windows_and_labels.py:86-270 with the
per-run numerics on the GPU (csrc/extract.cu).  The text parser stays on the host and mirrors
openlab_import.import_catman_file (openlab_import.py:33-85); everything after it is libshmfast."""
from __future__ import annotations

import ctypes as C
import os
import re
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import ShmfastError, check
from .ops import WindowSource, _need_cuda, _ptr, _stream

LABELS = ("Normal", "Sensor Fault", "Structural Fault")           # 01_extract_windows_and_labels.py:41-43
CATMAN_COLUMNS = ["Time_1", "DMS_1", "Time_2", "Force_N", "Force_A", "IWA", "Temp_Bridge", "Temp_Ambient", "Time_3", "LWA_1",
                  "LWA_2", "LWA_3", "Time_4", "LWA_4", "LWA_5", "NMA_5", "F_total", "Comment"]   # openlab_import.py:26-29
USED_COLUMNS = ("DMS_1", "LWA_2", "LWA_3", "LWA_4")
_T0 = re.compile(r"T0\s*=\s*(\d{2})\.(\d{2})\.(\d{4})\s+(\d{2}):(\d{2}):(\d{2})")


def read_catman_columns(path: os.PathLike) -> np.ndarray:
    """The four columns the path uses, as float32 [R, 4] (import_catman_file + _to_float: tab separated, decimal comma,
    cp1252, 36 header lines, bad lines skipped, non-numeric -> NaN)."""
    import pandas as pd
    with open(os.fspath(path), encoding="cp1252") as f:
        head = [next(f, "") for _ in range(13)]
    if _T0.search(head[12]) is None:
        raise ValueError(f"T0 not found in header of {os.fspath(path)!r} (expected 'T0 = DD.MM.YYYY HH:MM:SS' on line 13)")
    df = pd.read_csv(os.fspath(path), sep="\t", decimal=",", encoding="cp1252", skiprows=36, on_bad_lines="skip")
    df.columns = CATMAN_COLUMNS
    return np.stack([pd.to_numeric(df[c], errors="coerce").to_numpy(dtype=np.float32) for c in USED_COLUMNS], axis=1)


def default_cfg() -> _lib.ExtractCfg:
    """config.py:27-56 and STRUCT_CLEAN_CHANNELS = ["LWA_3"] (01_extract_windows_and_labels.py:50)."""
    return _lib.ExtractCfg(200, 20, 5, 0b010, -1e5, 1.0, 65.0, 1.0, 65.0, 0.05, 1e-6, 5.0, 20.0)


@dataclass
class ExtractedRun:
    a_clean: torch.Tensor          # [rows_kept, 4] float32: DMS_1, LWA_2/3/4 cleaned
    a_raw: torch.Tensor            # [rows_kept, 4] float32: DMS_1, LWA_2/3/4 raw with NaN for the obstruction sentinel
    n_windows: int
    win_start_idx: torch.Tensor    # [n_windows] int64
    label: torch.Tensor            # [n_windows] int32 (index into LABELS)
    meta: dict                     # u_min, u_max, dms_range, raw_invalid_ratio, raw_outlier_ratio, removed_ratio,
                                   # flatline_loadaware, all_nan_struct  (window_labels.csv columns)
    T: int
    stride: int

    def sources(self, gate_chan, vae_mu, vae_sd, cnn_mu, cnn_sd, clip: float = 10.0):
        """Window sources over the device-resident series for HybridOpenLab.run (10_test_hybrid_pipeline.py:351,272-278)."""
        g = WindowSource(self.a_clean, self.T, stride=self.stride, chan=list(gate_chan), mean=vae_mu, std=vae_sd, clip=clip,
                         nan_to_zero=True)
        r = WindowSource(self.a_raw, self.T, stride=self.stride, mean=cnn_mu, std=cnn_sd, clip=clip, nan_to_zero=True)
        return g, r


def extract_run(raw: torch.Tensor, cfg: Optional[_lib.ExtractCfg] = None) -> ExtractedRun:
    """raw [R,4] float32 CUDA tensor (read_catman_columns(...) moved to the device) -> ExtractedRun.  One host sync (the two
    counts) sizes the returned views."""
    lib = _lib.load()
    _need_cuda(raw, "raw")
    if raw.dtype != torch.float32 or raw.dim() != 2 or raw.shape[1] != 4:
        raise ShmfastError("raw must be float32 [R, 4] = DMS_1, LWA_2, LWA_3, LWA_4")
    raw = raw.contiguous()
    cfg = default_cfg() if cfg is None else cfg
    R = raw.shape[0]
    dev = raw.device
    maxW = 1 if R < cfg.T else (R - cfg.T) // cfg.stride + 1
    a_clean = torch.empty((R, 4), dtype=torch.float32, device=dev)
    a_raw = torch.empty((R, 4), dtype=torch.float32, device=dev)
    counts = torch.zeros((2,), dtype=torch.int32, device=dev)
    f = lambda: torch.empty((maxW,), dtype=torch.float32, device=dev)
    i = lambda: torch.empty((maxW,), dtype=torch.int32, device=dev)
    label, u_min, u_max, dms_range, inv_r, out_r, rem_r, flat, alln = i(), f(), f(), f(), f(), f(), f(), i(), i()
    ws = torch.empty((int(lib.shm_openlab_extract_workspace_bytes(R)),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.shm_openlab_extract(_ptr(raw), R, C.byref(cfg), _ptr(a_clean), _ptr(a_raw), _ptr(counts[0:1]), _ptr(counts[1:2]),
                                      _ptr(label), _ptr(u_min), _ptr(u_max), _ptr(dms_range), _ptr(inv_r), _ptr(out_r), _ptr(rem_r),
                                      _ptr(flat), _ptr(alln), _ptr(ws), _stream()), "shm_openlab_extract")
    rows_kept, nW = (int(v) for v in counts.cpu())
    meta = dict(u_min=u_min[:nW], u_max=u_max[:nW], dms_range=dms_range[:nW], raw_invalid_ratio=inv_r[:nW],
                raw_outlier_ratio=out_r[:nW], removed_ratio=rem_r[:nW], flatline_loadaware=flat[:nW], all_nan_struct=alln[:nW])
    starts = torch.arange(nW, dtype=torch.int64, device=dev) * cfg.stride
    return ExtractedRun(a_clean[:rows_kept], a_raw[:rows_kept], nW, starts, label[:nW], meta, int(cfg.T), int(cfg.stride))
