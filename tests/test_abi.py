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


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/pyrayt_b200.h is the boundary a C / cgo / JNI caller binds: it must compile as C11 and the
    library must link and answer from a C program (entry points that need no GPU)."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "use_abi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "pyrayt_b200.h"\n'
        "int main(void) {\n"
        "  prt_params p; memset(&p, 0, sizeof p);\n"
        "  if (prt_abi_version() != PRT_ABI_VERSION) return 1;\n"
        "  if (prt_tile_rays() != 256) return 2;\n"
        "  if (prt_axis_table_blocks(2049) != 3) return 3;\n"
        "  /* argument validation happens before any CUDA call */\n"
        "  if (prt_trace(NULL, &p, NULL, 0, 0, NULL, NULL, NULL) == PRT_OK) return 4;\n"
        "  if (strlen(prt_last_error()) == 0) return 5;\n"
        '  printf("abi %d\\n", prt_abi_version());\n'
        "  return 0;\n}\n")
    exe = tmp_path / "use_abi"
    libdir = os.path.join(ROOT, "pyrayt_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
           "-o", str(exe), "-L", libdir, "-lpyrayt_b200", f"-Wl,-rpath,{libdir}"]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip() == f"abi {_lib.ABI_VERSION}", (run.returncode, run.stderr)
