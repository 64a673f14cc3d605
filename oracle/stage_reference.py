"""Stage the UNMODIFIED PyRayT reference under ``baseline/_ref`` so that it can travel to the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (nothing under pyrayt_b200/ imports it).

``baseline/_ref`` is git-ignored (never committed: reference sources do not enter the history) but
not gpurun-ignored, so the install ships with the repo snapshot to the B200 box.  There it is what
``bench.py``'s NumPy arm times (``cpu_baseline.kind = "reference"``) and what the ``gpu`` +
``reference`` drop-in tests compare the CUDA engine with in one process.

The bench contract's own recipe

    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse \
        --target baseline/_ref /root/reference

fails in this image: the reference's build backend is ``poetry.core.masonry.api`` and poetry-core
is in no offline wheelhouse, and its ``python = ">=3.7.0 <3.9"`` pin excludes 3.12.  So the two
packages (``pyrayt``, ``tinygfx``) are installed from a copy under /tmp whose only change is the
packaging metadata (a three-line setuptools ``setup.py`` instead of ``pyproject.toml``); the
``examples/`` scripts are copied beside them because the drop-in tests run them unchanged.

    python oracle/stage_reference.py            # idempotent
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PYRAYT_REF", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

SETUP = """from setuptools import find_packages, setup
setup(name="pyrayt", version="0.3.1", packages=find_packages(include=["pyrayt*", "tinygfx*"]))
"""


def staged() -> bool:
    return os.path.isdir(os.path.join(DST, "pyrayt")) and os.path.isdir(os.path.join(DST, "tinygfx"))


def stage(force: bool = False) -> str:
    if staged() and not force:
        return DST
    if not os.path.isdir(os.path.join(SRC, "pyrayt")):
        raise RuntimeError(f"PyRayT reference tree not found at {SRC}")
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    with tempfile.TemporaryDirectory(prefix="pyrayt_src_") as tmp:
        for pkg in ("pyrayt", "tinygfx"):
            shutil.copytree(os.path.join(SRC, pkg), os.path.join(tmp, pkg),
                            ignore=shutil.ignore_patterns("__pycache__"))
        with open(os.path.join(tmp, "setup.py"), "w") as fh:
            fh.write(SETUP)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-deps",
                               "--no-build-isolation", "--no-compile", "--target", DST, tmp])
    shutil.copytree(os.path.join(SRC, "examples"), os.path.join(DST, "examples"),
                    ignore=shutil.ignore_patterns("__pycache__", ".ipynb_checkpoints"))
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
