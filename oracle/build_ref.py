"""TEST INFRASTRUCTURE ONLY — recipe for `oracle/_ref`: an UNMODIFIED copy of the reference's Python package (and of its
own hot-path tests) taken from where the sources lie under /root/reference.

The reference is pure Python on PyTorch: there is nothing to compile.  `oracle/_ref/` is git-ignored (no reference source
enters the history) but not gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist; it
is what `bench.py --impl reference`, the `gpu_eager_baseline` leg and `tests/test_reference_suite.py` run there.

    python oracle/build_ref.py          # idempotent; a no-op when /root/reference is absent
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("FLOWCON_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
# the package, and the reference's own tests of the hot path
TREES = ["flowcon", "tests"]


def build(verbose=False):
    if not os.path.isfile(os.path.join(SRC, "flowcon", "__init__.py")):
        if verbose:
            print("oracle/_ref: no reference tree at {}; keeping whatever is there".format(SRC))
        return os.path.isdir(os.path.join(DST, "flowcon"))
    os.makedirs(DST, exist_ok=True)
    for tree in TREES:
        src, dst = os.path.join(SRC, tree), os.path.join(DST, tree)
        if not os.path.isdir(src):
            continue
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache"))
    if verbose:
        print("oracle/_ref: copied {} from {}".format(TREES, SRC))
    return True


if __name__ == "__main__":
    sys.exit(0 if build(verbose=True) else 1)
