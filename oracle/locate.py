"""TEST INFRASTRUCTURE ONLY — locate and import the real FlowConductor reference (`flowcon`).

The reference is pure Python on PyTorch; it imports here once a few stub packages
(`oracle/stubs`: matplotlib, UMNN, torchdiffeq, torchtestcase, parameterized) are on sys.path
(see SURVEY.md Appendix A).  The reference tree lives at /root/reference in the build
container and does NOT exist on the GPU box; `oracle/build_ref.py` leaves an unmodified copy of
its package under `oracle/_ref/` (git-ignored, shipped with gpurun), which is what is found there.
Used by `oracle/make_golden.py`, by the not-gpu tests that pin the restatement (`oracle/restated.py`)
against the live reference, by bench.py's reference legs and by tests/test_reference_suite.py.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = os.path.join(_HERE, "stubs")
CANDIDATES = [os.environ.get("FLOWCON_REFERENCE", ""), "/root/reference", os.path.join(_HERE, "_ref")]


def reference_root():
    for root in CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "flowcon", "__init__.py")):
            return root
    return None


def have_reference():
    return reference_root() is not None


def import_reference():
    """Return the imported `flowcon` package of the unmodified reference, or raise ImportError."""
    root = reference_root()
    if root is None:
        raise ImportError("FlowConductor reference tree not found (looked in {})".format(CANDIDATES))
    for path in (root, STUBS):
        if path not in sys.path:
            sys.path.insert(0, path)
    import flowcon  # noqa: F401

    return flowcon
