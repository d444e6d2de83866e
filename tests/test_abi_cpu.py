"""CPU: the C-ABI shared library builds, loads and exports every symbol that
include/stif_b200.h declares; without a GPU the product path fails loudly."""
import ctypes
import os
import re

import pytest

from stif_b200 import _native, build


HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include",
                      "stif_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stif_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in stif_b200.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_native.SIGNATURES) == names


def test_version_and_error_string():
    lib = _native.load_library()
    assert lib.stif_version() >= 100
    assert isinstance(lib.stif_last_error(), bytes)


def test_no_cpu_fallback():
    """No GPU here: creating a context must raise, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.StifNativeError):
        _native.Context(0)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "stif_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "liboracle" not in src, f


def test_flag_constants_match_header():
    """The bit values of the ctypes layer are the header's #defines (STIF_KRIG_*, STIF_VARIO_*)."""
    text = open(HEADER).read()
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(STIF_(?:KRIG|VARIO)_[A-Z0-9_]+)\s+(\d+)", text)}
    assert "STIF_KRIG_TRY_LDL" in defines and "STIF_VARIO_DETERMINISTIC" in defines
    checked = 0
    for name, value in defines.items():
        py = name[len("STIF_"):]
        if hasattr(_native, py):
            assert getattr(_native, py) == value, name
            checked += 1
    assert checked >= 6
    # solver flags are distinct bits
    bits = [defines[n] for n in defines if n.startswith("STIF_KRIG_")]
    assert len(set(bits)) == len(bits) and all(b & (b - 1) == 0 for b in bits)
