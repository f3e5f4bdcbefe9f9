"""flowconductor_b200 — B200-native (sm_100a) drop-in for FlowConductor's element-wise bijection hot path.

Mirrors the `flowcon` API for that path only (`transforms`, `flows`, `distributions`, `nn.nets`,
`utils.torchutils`); the arithmetic runs in hand-written CUDA kernels behind a C-ABI shared library
(`include/flowcon_b200.h`).  There is no CPU fallback: calling a kernel-backed op without the built
library, or with CPU tensors, raises.
"""
__version__ = "0.1.0"


def patch_reference(flowcon=None, wrap=None):
    """Install the kernels behind the reference package's own spline functions (see flowconductor_b200/patch.py)."""
    from .patch import patch_reference as _patch

    return _patch(flowcon, wrap)
