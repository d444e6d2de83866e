"""Import the UNMODIFIED reference package: from /root/reference in the build container, else
from the vendored copy `oracle/_ref/` (made by oracle/vendor_ref.py, shipped with the snapshot).

TEST INFRASTRUCTURE -- not part of the product path.  `oracle/gen_golden.py` uses it to write
the golden vectors; `bench.py` uses it for the CPU baseline arm (`cpu_baseline`,
`--impl reference`), where it resolves to `oracle/_ref/` on the GPU box.

The reference imports matplotlib unconditionally (stif/predictor.py:4,26) and
matplotlib is not installed in this image, so a no-op stub is injected first.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("STIF_REFERENCE_ROOT", "/root/reference")


def root(vendored_only=False):
    """Directory that holds the `stif` package, or None."""
    cands = [VENDORED_ROOT] if vendored_only else [REFERENCE_ROOT, VENDORED_ROOT]
    for r in cands:
        if os.path.isdir(os.path.join(r, "stif")):
            return r
    return None


def available(vendored_only=False) -> bool:
    return root(vendored_only) is not None


def _stub_matplotlib():
    if "matplotlib" in sys.modules:
        return
    try:
        import matplotlib  # noqa: F401
        return
    except ImportError:
        pass
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")

    class _Style:
        def use(self, *a, **k):
            return None

    plt.style = _Style()
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


def load(vendored_only=False):
    """Return the reference `stif` module (numba-jitted, unmodified).  vendored_only: never
    touch /root/reference (bench.py and anything else that runs on the GPU box)."""
    r = root(vendored_only)
    if r is None:
        raise RuntimeError(f"reference package not found ({REFERENCE_ROOT} or {VENDORED_ROOT}; "
                           "run `python -m oracle.vendor_ref` in the build container)")
    _stub_matplotlib()
    if r not in sys.path:
        sys.path.insert(0, r)
    import stif  # noqa: E402
    return stif


def load_pm10():
    """PM10 Germany fixture of the reference (tests/data/pm10.csv) as arrays."""
    import pandas as pd
    df = pd.read_csv(os.path.join(REFERENCE_ROOT, "tests", "data", "pm10.csv"), index_col=0)
    return df[["x", "y", "time", "PM10"]]
