"""Validates the tcgen05 conventions of csrc/tcgen05.cuh on the real part: shared-memory descriptor
(K-major, no swizzle, LBO/SBO), instruction descriptor, TMEM accumulator layout, A-from-TMEM packing."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def kmajor_image(mat_bf16: torch.Tensor) -> torch.Tensor:
    """[rows, K] bf16 -> UMMA K-major no-swizzle image (8x8 core matrices, 128 B each)."""
    rows, K = mat_bf16.shape
    img = mat_bf16.view(rows // 8, 8, K // 8, 8).permute(2, 0, 1, 3).contiguous()     # [k/8][r/8][r%8][k%8]
    return img.view(-1)


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp("tcprobe") / "libtcprobe.so"
    cmd = ["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-shared",
           "-o", str(out), str(ROOT / "tests/cuda/tc_probe.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(str(out))
    lib.tc_probe.restype = C.c_int
    lib.tc_probe.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    return lib


def run(lib, dev, mode, n_cols, lbo=0, sbo=128):
    g = torch.Generator().manual_seed(n_cols + mode)
    A = torch.randn((128, 64), generator=g).to(torch.bfloat16)
    B = torch.randn((n_cols, 64), generator=g).to(torch.bfloat16)
    ref = (A.float() @ B.float().T)
    a_img, b_img = kmajor_image(A).to(dev), kmajor_image(B).to(dev)
    a_rows = A.contiguous().view(torch.int32).to(dev)                 # [128][32] packed pairs, k even in the low half
    out = torch.zeros((128, n_cols), dtype=torch.float32, device=dev)
    rc = lib.tc_probe(mode, lbo, sbo, a_img.data_ptr(), b_img.data_ptr(), a_rows.data_ptr(), out.data_ptr(), n_cols)
    assert rc == 0
    got = out.cpu()
    return float((got - ref).abs().max()), float(ref.abs().max())


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n_cols", [128, 256, 16])
def test_tcgen05_tile(cuda_dev, probe, mode, n_cols):
    err, scale = run(probe, cuda_dev, mode, n_cols)
    print(f"mode={'TS' if mode else 'SS'} N={n_cols} max|err|={err:.3e} (|ref|max {scale:.1f})")
    assert err < 1e-3 * scale
