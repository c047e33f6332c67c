"""Stage the UNMODIFIED reference package for the reference arm / CPU baseline.

    python oracle/make_ref.py        (also run by __graft_entry__.build() when /root/reference exists)

The reference is pure Python (`/root/reference/src/myrtle_vision`, no build step): the equivalent of
`pip install --target baseline/_ref /root/reference` is a byte-for-byte copy of that package into
`oracle/_ref/` — git-ignored, so no reference source enters the history, but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot (where /root/reference does not exist).  `bench.py --impl
reference` and the `cpu_baseline` leg import it from there through `oracle/shim/qtorch` (qtorch==0.3.0
cannot be installed offline; the shim is the C restatement of its CPU kernels, oracle/quant_oracle.c).
TEST / MEASUREMENT INFRASTRUCTURE ONLY: nothing under myrtle-vision_b200/ imports it.
"""
import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/src/myrtle_vision"
DST = os.path.join(ROOT, "oracle", "_ref", "myrtle_vision")


def staged():
    return os.path.exists(os.path.join(DST, "models", "vit.py"))


def make(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/make_ref.py: %s not present (GPU box?) — using the staged copy: %s"
                  % (SRC, "found" if staged() else "MISSING"))
        return staged()
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    same = filecmp.cmp(os.path.join(SRC, "models", "vit.py"), os.path.join(DST, "models", "vit.py"), shallow=False)
    assert same
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print("oracle/make_ref.py: staged %d files of the unmodified reference package in %s" % (n, DST))
    return True


def import_reference():
    """-> the reference's `myrtle_vision.models.vit` module, imported from oracle/_ref through the qtorch shim.
    Must run in a process that has not imported this repo's own `myrtle_vision` package."""
    if not staged():
        raise ImportError("oracle/_ref is not staged: run `python oracle/make_ref.py` where /root/reference exists")
    if "myrtle_vision" in sys.modules:
        mod = sys.modules["myrtle_vision"]
        if not os.path.abspath(mod.__file__).startswith(os.path.join(ROOT, "oracle", "_ref")):
            raise ImportError("another `myrtle_vision` is already imported in this process: " + mod.__file__)
    for p in (os.path.join(ROOT, "oracle", "shim"), os.path.join(ROOT, "oracle", "_ref"), ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import importlib
    return importlib.import_module("myrtle_vision.models.vit")


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
