"""Import the UNMODIFIED PyRayT reference (pure Python).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  Used (1) to validate the C restatement in
``oracle/trace_oracle.c`` against the real reference, (2) to generate the golden fixtures
committed under ``tests/golden/``, (3) by ``bench.py``'s NumPy arm (``cpu_baseline.kind =
"reference"``) and (4) by the ``gpu`` + ``reference`` drop-in tests.  Nothing under
``pyrayt_b200/`` imports this file.

Search order: ``$PYRAYT_REF``, ``/root/reference`` (build container), ``baseline/_ref`` (the
git-ignored install made by ``oracle/stage_reference.py``; it travels with the repo snapshot, so
it is what the GPU box sees -- ``/root/reference`` does not exist there and is never read at run
time by the ``-m gpu`` tests, ``smoke()`` or ``bench.py``).

Two shims are applied *outside* the read-only reference tree:

1. ``matplotlib`` is not installed but is imported at module import time by
   ``pyrayt/_pyrayt.py:4`` and ``tinygfx/g3d/renderers.py:6``; an empty stub is
   registered (only ``show()``/``draw()`` would use it).
2. pandas 3 removed ``DataFrame.append`` (used at ``pyrayt/_pyrayt.py:186``);
   it is re-added with the pandas-1.2 semantics the reference relied on.

``stable_argsort()`` additionally makes ``np.argsort`` default to
``kind="stable"``: the reference pins numpy 1.20.2, whose argsort of <=16
element lanes is an insertion sort (stable); numpy 2.x on AVX-512 hosts is not
stable and not reproducible across machines.  The stable
order is the parity contract.
"""
import contextlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("PYRAYT_REF"), "/root/reference", STAGED):
        if cand and os.path.isdir(os.path.join(cand, "pyrayt")) and os.path.isdir(os.path.join(cand, "tinygfx")):
            return cand
    return os.environ.get("PYRAYT_REF", "/root/reference")


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pyrayt"))


def examples_dir() -> str:
    """The reference's examples/ directory (scripts the drop-in tests run unchanged)."""
    return os.path.join(REF_ROOT, "examples")


def load():
    """Returns the imported reference ``pyrayt`` module (shims applied)."""
    if not available():
        raise RuntimeError(f"PyRayT reference not found at {REF_ROOT}")
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    import pandas as pd

    if not hasattr(pd.DataFrame, "append"):

        def _append(self, other, ignore_index=False):
            if len(self) == 0:
                return other.reset_index(drop=True)
            return pd.concat([self, other], ignore_index=ignore_index)

        pd.DataFrame.append = _append
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import pyrayt  # noqa: E402

    return pyrayt


@contextlib.contextmanager
def stable_argsort():
    """np.argsort defaults to kind='stable' inside the block (numpy 1.20 behaviour)."""
    import numpy as np

    orig = np.argsort

    def _stable(a, axis=-1, kind=None, order=None, **kw):
        return orig(a, axis=axis, kind="stable" if kind is None else kind, order=order, **kw)

    np.argsort = _stable
    try:
        yield
    finally:
        np.argsort = orig
