"""The C-ABI library loads and exports every symbol include/pyrayt_b200.h declares (no GPU calls)."""
import ctypes
import os
import re

import pytest

from pyrayt_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "pyrayt_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(prt_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.prt_abi_version() == _lib.ABI_VERSION
    assert lib.prt_tile_rays() == 256


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.PrtParams) == 32
    assert ctypes.sizeof(_lib.PrtRecords) == 48
    assert ctypes.sizeof(_lib.PrtSourceDesc) == 16 + 24 + 128
    assert _lib.COUNTER_WORDS * 8 == 128  # sizeof(prt_counters)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without the CUDA library the product raises instead of computing."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.PrtError, match="no CPU"):
        _lib.load()


def test_engine_without_cuda_raises():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import pyrayt_b200
    from tests import scene_util as su

    with pytest.raises(pyrayt_b200.PrtError, match="no CPU fallback"):
        pyrayt_b200.Engine(su.build([su.Leaf(su.SPHERE, [1])]))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure; nothing under pyrayt_b200/ may reference it."""
    pkg = os.path.join(ROOT, "pyrayt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libprt_oracle" not in text and "trace_oracle" not in text, f
