import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """libstif_b200 context on cuda:0 (GPU tests only; fails loudly without a GPU)."""
    from stif_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def record_quantiles(name, **series):
    """Print (and, on the GPU box, append to gpurun_out/parity_quantiles.jsonl) the p50 / p99 /
    max of error series, so that the bounds the tests assert are the measured ones."""
    import json

    import numpy as np
    row = {"name": name}
    for key, vals in series.items():
        v = np.asarray(vals, dtype=np.float64)
        v = v[np.isfinite(v)]
        if v.size == 0:
            row[key] = None
            continue
        row[key] = {"n": int(v.size), "p50": float(np.quantile(v, 0.5)), "p99": float(np.quantile(v, 0.99)),
                    "max": float(v.max())}
    print("PARITY", json.dumps(row))
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "parity_quantiles.jsonl"), "a") as fh:
            fh.write(json.dumps(row) + "\n")
    except OSError:
        pass
    return row
