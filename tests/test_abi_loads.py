"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/geomae_b200.h declares (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "geomae_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(geomae_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from geomae_b200 import lib as L
    L.build_if_missing()
    return L.lib()


def test_header_declares_something():
    syms = declared_symbols()
    assert "geomae_voxel_scatter" in syms and "geomae_dynamic_voxelize" in syms


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_error_reporting_without_gpu(lib):
    assert lib.geomae_abi_version() >= 1
    g = (ctypes.c_int32 * 3)()
    f3 = lambda *v: (ctypes.c_float * 3)(*v)  # noqa: E731
    assert lib.geomae_grid_size(f3(-51.2, -51.2, -5), f3(51.2, 51.2, 3), f3(0.256, 0.256, 8), g) == 0
    assert list(g) == [400, 400, 1]
    assert lib.geomae_grid_size(f3(-51.2, -51.2, -5), f3(51.2, 51.2, 3), f3(0.064, 0.064, 1), g) == 0
    assert list(g) == [1600, 1600, 8]
    assert lib.geomae_grid_size(f3(0, 0, 0), f3(1, 1, 1), f3(0, 1, 1), g) == -1
    assert b"voxel" in lib.geomae_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from geomae_b200 import lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libgeomae_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        L.lib()


def test_cpu_tensor_is_rejected():
    import torch
    from geomae_b200.voxel import Voxelization
    vox = Voxelization((0.256, 0.256, 8), [-51.2, -51.2, -5, 51.2, 51.2, 3], -1, (-1, -1))
    with pytest.raises(RuntimeError, match="CUDA"):
        vox(torch.zeros(4, 5))
