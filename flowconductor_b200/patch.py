"""`patch_reference()` — install the CUDA kernels behind the REFERENCE package's own function-level seam (SURVEY.md 7
step 1, 8b): the reference has no plugin / FFI interface, its hot functions are plain module attributes.

    import flowcon, flowconductor_b200
    undo = flowconductor_b200.patch_reference(flowcon)     # flowcon's coupling / autoregressive / CDF layers now run their
    ...                                                     # spline through libflowcon_b200.so (CUDA tensors only)
    undo()

What is replaced (same signatures, same return values, same exceptions):
  rational_quadratic_spline / unconstrained_rational_quadratic_spline   flowcon/transforms/splines/rational_quadratic.py:13-181
  linear_spline / unconstrained_linear_spline                           flowcon/transforms/splines/linear.py:9-105
  quadratic_spline / unconstrained_quadratic_spline                     flowcon/transforms/splines/quadratic.py:11-159
  cubic_spline / unconstrained_cubic_spline                             flowcon/transforms/splines/cubic.py:15-267
in every module that holds a binding: `flowcon.transforms.splines` and its sub-modules (coupling.py and nonlinearities.py
look the functions up there by attribute at call time: coupling.py:342-346,412-415,483-486,566-569), and the module
globals that were bound at import (autoregressive/autoregressive.py:9-19, conditional.py:8-13,
autoregressive/deep_sigmoid.py:7-17).  There is no CPU fallback: the patched functions raise on CPU tensors like every
other entry point of this package.
"""
import functools
import importlib

from .transforms import splines as _ours
from .transforms.base import InputOutsideDomain as _OurInputOutsideDomain

FUNCTIONS = ("rational_quadratic_spline", "unconstrained_rational_quadratic_spline", "linear_spline",
             "unconstrained_linear_spline", "quadratic_spline", "unconstrained_quadratic_spline", "cubic_spline",
             "unconstrained_cubic_spline")
MODULES = ("flowcon.transforms.splines", "flowcon.transforms.splines.rational_quadratic",
           "flowcon.transforms.splines.linear", "flowcon.transforms.splines.quadratic",
           "flowcon.transforms.splines.cubic", "flowcon.transforms.coupling", "flowcon.transforms.nonlinearities",
           "flowcon.transforms.conditional", "flowcon.transforms.autoregressive.autoregressive",
           "flowcon.transforms.autoregressive.deep_sigmoid")


def patch_reference(flowcon=None, wrap=None):
    """Replace the reference's spline functions by the kernels.  `flowcon`: the imported reference package (default:
    `import flowcon`).  `wrap`: optional decorator applied to each replacement (tests use it to move CPU tensors to the
    GPU and back).  Returns a callable that restores the original functions."""
    if flowcon is None:
        flowcon = importlib.import_module("flowcon")
    originals = {}
    for name in FUNCTIONS:
        src = getattr(importlib.import_module("flowcon.transforms.splines"), name, None)
        if src is not None:
            originals[name] = src
    ref_exc = importlib.import_module("flowcon.transforms.base").InputOutsideDomain

    def reference_exceptions(fn):
        # the reference's callers (and its tests) catch flowcon.transforms.base.InputOutsideDomain
        @functools.wraps(fn)
        def call(*args, **kwargs):
            try:
                return fn(*args, **kwargs)
            except _OurInputOutsideDomain:
                raise ref_exc() from None
        return call

    saved = []
    for modname in MODULES:
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        for name, orig in originals.items():
            if getattr(mod, name, None) is orig:
                repl = reference_exceptions(getattr(_ours, name))
                if wrap is not None:
                    repl = wrap(repl)
                saved.append((mod, name, orig))
                setattr(mod, name, repl)

    def undo():
        for mod, name, orig in saved:
            setattr(mod, name, orig)

    undo.patched = [(m.__name__, n) for m, n, _ in saved]
    return undo
