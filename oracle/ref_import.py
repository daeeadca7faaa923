"""TEST INFRASTRUCTURE ONLY -- import the *unmodified* reference ``Python/apvast.py`` in place.

The reference imports two modules that are not installed in this image and that the
``perceptual=False`` path never touches (``matplotlib.pyplot``: imported, unused,
reference ``Python/apvast.py:3``; ``libdetectability``: only used when ``perceptual=True``,
``Python/apvast.py:4,77-83,318-319``).  Both are stubbed in ``sys.modules`` before the import.

``/root/reference`` exists only in the build container, never on the GPU box, so this module is
used by ``oracle/make_golden.py`` (fixture generation) and by the CPU tests that pin the NumPy
restatement against the live reference; those tests skip when the reference is absent.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_CANDIDATES = (
    os.environ.get("APVAST_REF", ""),
    "/root/reference/Python",
)


def reference_dir():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "apvast.py")):
            return c
    return None


def have_reference() -> bool:
    return reference_dir() is not None


def load_reference():
    """Return the reference module (``apvast`` class, ``jdiag`` function)."""
    d = reference_dir()
    if d is None:
        raise FileNotFoundError("reference Python/apvast.py not found (set $APVAST_REF)")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "libdetectability" not in sys.modules:
        ld = types.ModuleType("libdetectability")

        class Detectability:  # pragma: no cover - placeholder, replaced by tests when needed
            def __init__(self, *a, **k):
                raise RuntimeError("libdetectability is not installed; inject a model instead")

        ld.Detectability = Detectability
        sys.modules["libdetectability"] = ld
    spec = importlib.util.spec_from_file_location("apvast_reference", os.path.join(d, "apvast.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_rirs():
    """``Python/rirs.mat`` -> (rirA, rirB), float64 (800, 8, 9)."""
    import numpy as np
    import scipy.io as sio

    d = reference_dir()
    m = sio.loadmat(os.path.join(d, "rirs.mat"))
    return np.ascontiguousarray(m["rirA"], dtype=np.float64), np.ascontiguousarray(m["rirB"], dtype=np.float64)
