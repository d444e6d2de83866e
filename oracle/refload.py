"""Import the UNMODIFIED reference package from /root/reference (this container only).

TEST INFRASTRUCTURE -- not part of the product path.  Only `oracle/gen_golden.py`
and ad-hoc validation scripts use this; nothing that runs on the GPU box may
import it (`/root/reference` does not exist there).

The reference imports matplotlib unconditionally (stif/predictor.py:4,26) and
matplotlib is not installed in this image, so a no-op stub is injected first.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STIF_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "stif"))


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


def load():
    """Return the reference `stif` module (numba-jitted, unmodified)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub_matplotlib()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import stif  # noqa: E402
    return stif


def load_pm10():
    """PM10 Germany fixture of the reference (tests/data/pm10.csv) as arrays."""
    import pandas as pd
    df = pd.read_csv(os.path.join(REFERENCE_ROOT, "tests", "data", "pm10.csv"), index_col=0)
    return df[["x", "y", "time", "PM10"]]
