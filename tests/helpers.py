"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    """Load tests/golden/<name>.npz as {key: torch tensor}."""
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def golden_state(gold, dtype=None, device=None):
    state = {}
    for k, v in gold.items():
        if k.startswith("state/"):
            if dtype is not None and v.is_floating_point():
                v = v.to(dtype)
            if device is not None:
                v = v.to(device)
            state[k[len("state/"):]] = v
    return state


def rel_err(a, b, floor):
    """|a-b| / max(|b|, floor), elementwise, computed in fp64."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return (a - b).abs() / b.abs().clamp_min(floor)


def assert_parity(ours, ref32, ref64, tol, floor, what="", slack=4.0):
    """Three-way parity (SURVEY.md §7 'the 1e-5 tolerance is at the reference's own fp32 noise floor'):
    every element must satisfy EITHER  |ours - ref32| <= tol * max(|ref32|, floor)
    OR      |ours - ref64| <= slack * |ref32 - ref64| + tol * max(|ref64|, floor)
    i.e. wherever we differ from the fp32 reference by more than `tol` we must be no further from the
    fp64 ground truth than the fp32 reference itself is (ill-conditioned elements)."""
    ours = ours.detach().double().cpu()
    r32 = ref32.detach().double().cpu()
    r64 = ref64.detach().double().cpu()
    assert ours.shape == r32.shape, (what, ours.shape, r32.shape)
    assert torch.isfinite(ours).all(), what + ": non-finite values"
    e32 = (ours - r32).abs()
    ok1 = e32 <= tol * r32.abs().clamp_min(floor)
    e64 = (ours - r64).abs()
    noise = (r32 - r64).abs()
    ok2 = e64 <= slack * noise + tol * r64.abs().clamp_min(floor)
    bad = ~(ok1 | ok2)
    if bad.any():
        i = torch.nonzero(bad)[0]
        idx = tuple(i.tolist())
        raise AssertionError(
            "{}: {} / {} elements fail parity; first at {}: ours={:.9g} ref32={:.9g} ref64={:.9g}".format(
                what, int(bad.sum()), bad.numel(), idx, ours[idx].item(), r32[idx].item(), r64[idx].item()))
    return float((e32 / r32.abs().clamp_min(floor)).max()) if e32.numel() else 0.0
