"""Import the UNMODIFIED PyRayT reference (pure Python) from /root/reference.

TEST INFRASTRUCTURE ONLY.  This module is used in the build container to
(1) validate the C restatement in ``oracle/trace_oracle.c`` against the real
reference and (2) generate the golden fixtures committed under
``tests/golden/``.  ``/root/reference`` does not exist on the GPU box, so
nothing on the product path, in ``bench.py`` or in the ``-m gpu`` tests may
import this file.

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

REF_ROOT = os.environ.get("PYRAYT_REF", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pyrayt"))


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
