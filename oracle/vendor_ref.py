"""Vendor the UNMODIFIED reference package for the CPU baseline arm -- TEST / BENCH INFRASTRUCTURE.

`/root/reference` exists only in the build container.  The reference is pure Python (numba does
the compiling at run time), so "building" it for the GPU box is a file copy: this recipe copies
`/root/reference/stif/*.py` (5 files) and the licence into `oracle/_ref/`, which is git-ignored
(no reference source enters the history) but NOT gpurun-ignored, so it travels with the repo
snapshot like the built `.so` files.  `__graft_entry__.build()` runs it whenever the reference
tree is present.

Only `bench.py` (`cpu_baseline`, `--impl reference`) and `oracle/gen_golden.py` import the
vendored package, through `oracle/refload.py`; the product (`stif_b200/`) never does.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("STIF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def vendor(force=False):
    """Copy the reference package; returns the vendored root or None if there is no reference."""
    src = os.path.join(REF_ROOT, "stif")
    if not os.path.isdir(src):
        return OUT if os.path.isdir(os.path.join(OUT, "stif")) else None
    dst = os.path.join(OUT, "stif")
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        a, b = os.path.join(src, name), os.path.join(dst, name)
        if force or not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
            shutil.copyfile(a, b)
    lic = os.path.join(REF_ROOT, "LICENSE")
    if os.path.exists(lic):
        shutil.copyfile(lic, os.path.join(OUT, "LICENSE"))
    with open(os.path.join(OUT, "README"), "w") as fh:
        fh.write("Unmodified copy of johanneskopton/stif (stif/*.py), made by oracle/vendor_ref.py for the\n"
                 "CPU baseline arm of bench.py.  Git-ignored; not part of the product.\n")
    return OUT


if __name__ == "__main__":
    print(vendor(force=True))
