"""CPU check of the exact per-element arithmetic the CUDA kernels execute.

`flowconductor_b200/csrc/fc_math.cuh` is `__host__ __device__`; tests/hostmath/hostmath.cpp compiles it
with g++ into a TEST-ONLY shared object.  Here its results are compared with the golden vectors of the
unmodified reference (fp32 and fp64) with the same three-way rule the GPU parity tests use.  This is what
lets the hand-derived backward formulas be validated in a container without a GPU; the product package
never loads this shim.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from tests.helpers import ROOT, assert_parity, load_golden

SRC = os.path.join(ROOT, "tests", "hostmath", "hostmath.cpp")
LIB = os.path.join(ROOT, "tests", "hostmath", "libhostmath.so")
HDR = os.path.join(ROOT, "flowconductor_b200", "csrc", "fc_math.cuh")


class RqsConfig(ctypes.Structure):
    _fields_ = [("num_bins", ctypes.c_int32), ("tails", ctypes.c_int32), ("identity_init", ctypes.c_int32),
                ("inverse", ctypes.c_int32), ("left", ctypes.c_float), ("right", ctypes.c_float),
                ("bottom", ctypes.c_float), ("top", ctypes.c_float), ("min_bin_width", ctypes.c_float),
                ("min_bin_height", ctypes.c_float), ("min_derivative", ctypes.c_float), ("wh_scale", ctypes.c_float)]


@pytest.fixture(scope="module")
def hm():
    newest = max(os.path.getmtime(SRC), os.path.getmtime(HDR))
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-x", "c++", SRC,
                               "-o", LIB])
    return ctypes.CDLL(LIB)


@pytest.fixture(scope="module")
def gold():
    return load_golden("functions")


def fptr(t):
    return ctypes.c_void_p(t.data_ptr())


def make_cfg(k, lin, tb, inv, ident, wh_scale=1.0):
    lo, hi = (-tb, tb) if lin else (0.0, 1.0)
    return RqsConfig(int(k), 1 if lin else 0, int(ident), int(inv), lo, hi, lo, hi, 1e-3, 1e-3, 1e-3, wh_scale)


RQ_CASES = ["rq_fwd_lin_k8", "rq_inv_lin_k8", "rq_fwd_lin_k16_id", "rq_inv_lin_k16_id", "rq_fwd_none_k5",
            "rq_inv_none_k5", "rq_fwd_lin_k10_b1", "rq_fwd_lin_k8_zero_id"]


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("name", RQ_CASES)
def test_rqs_forward_inverse(hm, gold, name, generic):
    k, lin, tb, inv, ident = gold[name + "/meta"].tolist()
    x = gold[name + "/x"].contiguous()
    p = gold[name + "/params"].contiguous()
    n = x.numel()
    y = torch.empty_like(x)
    lad = torch.empty_like(x)
    status = ctypes.c_uint(0)
    cfg = make_cfg(k, lin, tb, inv, ident)
    rc = hm.hm_rqs_apply(fptr(x), fptr(p), fptr(y), fptr(lad), ctypes.c_long(n), ctypes.byref(cfg), generic,
                         ctypes.byref(status))
    assert rc == 0 and status.value == 0
    floor = max(tb, 1.0)
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, floor, name + " outputs")
    assert_parity(lad, gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " logabsdet")


@pytest.mark.parametrize("generic", [0, 1])
@pytest.mark.parametrize("name", ["rq_fwd_lin_k8", "rq_inv_lin_k8", "rq_fwd_lin_k16_id", "rq_fwd_none_k5"])
def test_rqs_backward(hm, gold, name, generic):
    k, lin, tb, inv, ident = gold[name + "/meta"].tolist()
    x = gold[name + "/x"].contiguous()
    p = gold[name + "/params"].contiguous()
    gy = gold[name + "/gy"].contiguous()
    gl = gold[name + "/gl"].contiguous()
    gx = torch.empty_like(x)
    gp = torch.empty_like(p)
    cfg = make_cfg(k, lin, tb, inv, ident)
    rc = hm.hm_rqs_backward(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx), fptr(gp), ctypes.c_long(x.numel()),
                            ctypes.byref(cfg), generic)
    assert rc == 0
    s = max(1.0, gold[name + "/gx64"].abs().median().item())
    assert_parity(gx, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " grad x")
    s = max(1e-2, gold[name + "/gp64"].abs().mean().item())
    assert_parity(gp, gold[name + "/gp32"], gold[name + "/gp64"], 1e-4, s, name + " grad params")


def test_rqs_wh_scale_matches_prescaled_params(hm, gold):
    """wh_scale = 1/sqrt(H) must equal dividing the raw widths/heights beforehand (coupling.py:554-556)."""
    name = "rq_fwd_lin_k8"
    x = gold[name + "/x"].contiguous()
    p = gold[name + "/params"].clone()
    scale = 1.0 / 16.0
    p[..., :16] *= 16.0
    y, lad = torch.empty_like(x), torch.empty_like(x)
    status = ctypes.c_uint(0)
    cfg = make_cfg(8, 1, 3.0, 0, 0, wh_scale=scale)
    hm.hm_rqs_apply(fptr(x), fptr(p.contiguous()), fptr(y), fptr(lad), ctypes.c_long(x.numel()), ctypes.byref(cfg), 0,
                    ctypes.byref(status))
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, 3.0, "scaled outputs")


def test_rqs_domain_status(hm):
    x = torch.tensor([-0.1, 0.5, 1.1])
    p = torch.zeros(3, 16)
    y, lad = torch.empty_like(x), torch.empty_like(x)
    status = ctypes.c_uint(0)
    cfg = make_cfg(5, 0, 1.0, 0, 0)
    hm.hm_rqs_apply(fptr(x), fptr(p), fptr(y), fptr(lad), ctypes.c_long(3), ctypes.byref(cfg), 0, ctypes.byref(status))
    assert status.value & 1


def test_bin_indices_match_reference(hm, gold):
    """Bin parity (north_star): identical bins except for inputs within 1e-6 (normalised units) of a knot.
    Checked through the outputs: in a wrong bin the output error would be O(bin width)."""
    name = "rq_fwd_lin_k8"
    x = gold[name + "/x"]
    knots = gold[name + "/knots"]
    d = (x[..., None] - knots).abs().min(-1).values
    assert (d > 6e-6).float().mean() > 0.99  # the vectors do exercise the generic position


@pytest.mark.parametrize("act,layout", [(0, "blocked_sigmoid2"), (1, "blocked_softplus_clamp3"), (2, "interleaved")])
def test_affine(hm, gold, act, layout):
    x = gold["affine/x"].contiguous()
    p = gold["affine/params"]
    n, d = x.shape
    if layout.startswith("blocked"):
        shift, raw = p[:, :d].contiguous(), p[:, d:].contiguous()
    else:
        raw, shift = p.view(n, d, 2)[..., 0].contiguous(), p.view(n, d, 2)[..., 1].contiguous()
    y, lad = torch.empty_like(x), torch.empty_like(x)
    hm.hm_affine_apply(fptr(x), fptr(raw), fptr(shift), fptr(y), fptr(lad), ctypes.c_long(x.numel()), act, 0)
    assert_parity(y, gold["affine/%s_fwd_y32" % layout], gold["affine/%s_fwd_y64" % layout], 1e-5, 1.0, "affine y")
    assert_parity(lad.sum(1), gold["affine/%s_fwd_lad32" % layout], gold["affine/%s_fwd_lad64" % layout], 1e-5, 1.0,
                  "affine lad")
    hm.hm_affine_apply(fptr(x), fptr(raw), fptr(shift), fptr(y), fptr(lad), ctypes.c_long(x.numel()), act, 1)
    assert_parity(y, gold["affine/%s_inv_y32" % layout], gold["affine/%s_inv_y64" % layout], 1e-5, 1.0, "affine inv y")
    # backward against autograd of the same formulas (fp64)
    gy = torch.randn(n, d, generator=torch.Generator().manual_seed(1))
    gl = torch.randn(n, d, generator=torch.Generator().manual_seed(2))
    for inverse in (0, 1):
        gx, graw, gshift = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        hm.hm_affine_backward(fptr(x), fptr(raw), fptr(shift), fptr(gy.contiguous()), fptr(gl.contiguous()), fptr(gx),
                              fptr(graw), fptr(gshift), ctypes.c_long(x.numel()), act, inverse)
        xd, rd, sd = (t.double().clone().requires_grad_(True) for t in (x, raw, shift))
        if act == 0:
            scale = torch.sigmoid(rd + 2) + 1e-3
        elif act == 1:
            scale = (torch.nn.functional.softplus(rd) + 1e-3).clamp(0, 3)
        else:
            scale = torch.nn.functional.softplus(rd) + 1e-3
        yy = (xd - sd) / scale if inverse else xd * scale + sd
        ll = -torch.log(scale) if inverse else torch.log(scale)
        ex, er, es = torch.autograd.grad((yy * gy.double()).sum() + (ll * gl.double()).sum(), [xd, rd, sd])
        for ours, ref, what in ((gx, ex, "gx"), (graw, er, "graw"), (gshift, es, "gshift")):
            err = (ours.double() - ref).abs().max().item()
            assert err <= 1e-4 * max(1.0, ref.abs().max().item()), (what, inverse, err)


@pytest.mark.parametrize("name", ["sos_n10", "sos_n3_wide"])
def test_sum_of_sigmoids(hm, gold, name):
    ns = int(gold[name + "/meta"][0])
    x = gold[name + "/x"].contiguous()
    p = gold[name + "/params"].contiguous()
    n, d = x.shape
    y, lj = torch.empty_like(x), torch.empty_like(x)
    hm.hm_sos_apply(fptr(x), fptr(p), fptr(y), fptr(lj), ctypes.c_long(x.numel()), ns)
    floor = max(1.0, gold[name + "/y64"].abs().median().item())
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, floor, name + " y")
    assert_parity(lj.sum(-1), gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad")
    # backward
    gy = gold[name + "/gy"].contiguous()
    gl = gold[name + "/gl"][:, None].expand(n, d).contiguous()
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    hm.hm_sos_backward(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx), fptr(gp), ctypes.c_long(x.numel()), ns)
    s = max(1e-2, gold[name + "/gx64"].abs().mean().item())
    assert_parity(gx, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " gx")
    s = max(1e-2, gold[name + "/gp64"].abs().mean().item())
    assert_parity(gp, gold[name + "/gp32"], gold[name + "/gp64"], 1e-4, s, name + " gp")
    # numerical inverse: the reference's own tests accept 1e-5 .. 1e-3 (adaptive_sigmoid_test.py:40-41,64-79)
    z = gold[name + "/y32"].contiguous()
    xi, lji = torch.empty_like(x), torch.empty_like(x)
    hm.hm_sos_invert(fptr(z), fptr(p), fptr(xi), fptr(lji), ctypes.c_long(x.numel()), ns, 50, ctypes.c_float(120.0))
    ref = gold[name + "/inv_x64"]
    scale = ref.abs().clamp_min(1.0)
    assert ((xi.double() - x.double()).abs() / scale).max() < 1e-3
    assert ((xi.double() - ref).abs() / scale).max() < 1e-3
    assert (-lji.sum(-1).double() - gold[name + "/inv_lad64"]).abs().max() < 5e-3
    if ns == 10:  # the compile-time-n instantiations (paired reciprocals, two-pass backward) the n = 10 kernels run
        y10, lj10 = torch.empty_like(x), torch.empty_like(x)
        hm.hm_sos_apply_n10(fptr(x), fptr(p), fptr(y10), fptr(lj10), ctypes.c_long(x.numel()))
        assert_parity(y10, gold[name + "/y32"], gold[name + "/y64"], 1e-5, floor, name + " y (n10)")
        assert_parity(lj10.sum(-1), gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad (n10)")
        gx10, gp10 = torch.empty_like(x), torch.empty_like(p)
        hm.hm_sos_backward_n10(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx10), fptr(gp10), ctypes.c_long(x.numel()))
        s = max(1e-2, gold[name + "/gx64"].abs().mean().item())
        assert_parity(gx10, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " gx (n10)")
        s = max(1e-2, gold[name + "/gp64"].abs().mean().item())
        assert_parity(gp10, gold[name + "/gp32"], gold[name + "/gp64"], 1e-4, s, name + " gp (n10)")
        xi10, lji10 = torch.empty_like(x), torch.empty_like(x)
        hm.hm_sos_invert_n10.restype = ctypes.c_long
        worst = hm.hm_sos_invert_n10(fptr(z), fptr(p), fptr(xi10), fptr(lji10), ctypes.c_long(x.numel()), 50,
                                     ctypes.c_float(120.0))
        assert ((xi10.double() - ref).abs() / scale).max() < 1e-3
        assert (-lji10.sum(-1).double() - gold[name + "/inv_lad64"]).abs().max() < 5e-3
        assert worst <= 16, worst  # safeguarded Newton: far fewer evaluations than 50 bisections + brackets
        print("sos inverse: worst-case evaluations per element", worst)


@pytest.mark.parametrize("name", ["lin_fwd_k8", "lin_inv_k8", "lin_fwd_tails_k10", "lin_inv_tails_k10", "lin_fwd_k5"])
@pytest.mark.parametrize("unrolled", [0, 1])
def test_linear_spline(hm, name, unrolled):
    """linspline_eval / linspline_backward_elem (fc_math.cuh) against the reference's linear_spline golden vectors."""
    gold = load_golden("functions_linear")
    k, has_tails, tb, inverse = gold[name + "/meta"].tolist()
    k, has_tails, inverse = int(k), int(has_tails), int(inverse)
    lo, hi = (-tb, tb) if has_tails else (0.0, 1.0)
    x = gold[name + "/x"].contiguous()
    p = gold[name + "/params"].contiguous()
    y, lad = torch.empty_like(x), torch.empty_like(x)
    status = torch.zeros(1, dtype=torch.int32)
    hm.hm_linspline_apply(fptr(x), fptr(p), fptr(y), fptr(lad), fptr(status), ctypes.c_long(x.numel()), k, has_tails,
                          ctypes.c_float(lo), ctypes.c_float(hi), inverse, unrolled)
    assert int(status.item()) == 0
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, 1.0, name + " y")
    assert_parity(lad, gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad")
    gy, gl = gold[name + "/gy"].contiguous(), gold[name + "/gl"].contiguous()
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    hm.hm_linspline_backward(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx), fptr(gp), ctypes.c_long(x.numel()), k,
                             has_tails, ctypes.c_float(lo), ctypes.c_float(hi), inverse, unrolled)
    s = max(1e-2, gold[name + "/gx64"].abs().mean().item())
    assert_parity(gx, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " gx")
    s = max(1e-2, gold[name + "/gp64"].abs().mean().item())
    assert_parity(gp, gold[name + "/gp32"], gold[name + "/gp64"], 1e-4, s, name + " gp")


@pytest.mark.parametrize("name", ["quad_fwd_k8", "quad_inv_k8", "quad_fwd_tails_k10", "quad_inv_tails_k10", "quad_fwd_k5"])
@pytest.mark.parametrize("unrolled", [0, 1])
def test_quadratic_spline(hm, name, unrolled):
    """quadspline_eval / quadspline_backward_elem (fc_math.cuh) against the reference's quadratic_spline golden vectors."""
    gold = load_golden("functions_quadratic")
    k, has_tails, tb, inverse = gold[name + "/meta"].tolist()
    k, has_tails, inverse = int(k), int(has_tails), int(inverse)
    lo, hi = (-tb, tb) if has_tails else (0.0, 1.0)
    x = gold[name + "/x"].contiguous()
    p = torch.cat([gold[name + "/uw"], gold[name + "/uh"]], dim=-1).contiguous()
    y, lad = torch.empty_like(x), torch.empty_like(x)
    status = torch.zeros(1, dtype=torch.int32)
    hm.hm_quadspline_apply(fptr(x), fptr(p), fptr(y), fptr(lad), fptr(status), ctypes.c_long(x.numel()), k, has_tails,
                           ctypes.c_float(lo), ctypes.c_float(hi), inverse, unrolled)
    assert int(status.item()) == 0
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, 1.0, name + " y")
    assert_parity(lad, gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad")
    gy, gl = gold[name + "/gy"].contiguous(), gold[name + "/gl"].contiguous()
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    hm.hm_quadspline_backward(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx), fptr(gp), ctypes.c_long(x.numel()), k,
                              has_tails, ctypes.c_float(lo), ctypes.c_float(hi), inverse, unrolled)
    s = max(1e-2, gold[name + "/gx64"].abs().mean().item())
    assert_parity(gx, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " gx")
    g32 = torch.cat([gold[name + "/gw32"], gold[name + "/gh32"]], dim=-1)
    g64 = torch.cat([gold[name + "/gw64"], gold[name + "/gh64"]], dim=-1)
    s = max(1e-2, g64.abs().mean().item())
    assert_parity(gp, g32, g64, 1e-4, s, name + " gp")


@pytest.mark.parametrize("name", ["cubic_fwd_k8", "cubic_inv_k8", "cubic_fwd_k5", "cubic_inv_k5"])
@pytest.mark.parametrize("unrolled", [0, 1])
def test_cubic_spline_element_math(hm, name, unrolled):
    """cubicspline_eval / cubicspline_backward_elem (fc_math.cuh; not wired into a kernel yet) against the reference's
    cubic_spline golden vectors: the inverse is a safeguarded Newton iteration instead of Blinn's closed form."""
    gold = load_golden("functions_cubic")
    k, inverse = (int(v) for v in gold[name + "/meta"].tolist())
    x = gold[name + "/x"].contiguous()
    p = torch.cat([gold[name + "/uw"], gold[name + "/uh"], gold[name + "/dl"], gold[name + "/dr"]], dim=-1).contiguous()
    y, lad = torch.empty_like(x), torch.empty_like(x)
    hm.hm_cubicspline_apply(fptr(x), fptr(p), fptr(y), fptr(lad), ctypes.c_long(x.numel()), k, inverse, unrolled)
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, 1.0, name + " y")
    assert_parity(lad, gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad")
    gy, gl = gold[name + "/gy"].contiguous(), gold[name + "/gl"].contiguous()
    gx, gp = torch.empty_like(x), torch.empty_like(p)
    hm.hm_cubicspline_backward(fptr(x), fptr(p), fptr(gy), fptr(gl), fptr(gx), fptr(gp), ctypes.c_long(x.numel()), k,
                               inverse, unrolled)
    s = max(1e-2, gold[name + "/gx64"].abs().mean().item())
    assert_parity(gx, gold[name + "/gx32"], gold[name + "/gx64"], 1e-4, s, name + " gx")
    g32 = torch.cat([gold[name + "/" + key + "32"] for key in ("gw", "gh", "gdl", "gdr")], dim=-1)
    g64 = torch.cat([gold[name + "/" + key + "64"] for key in ("gw", "gh", "gdl", "gdr")], dim=-1)
    s = max(1e-2, g64.abs().mean().item())
    assert_parity(gp, g32, g64, 1e-4, s, name + " gp")


@pytest.mark.parametrize("inverse", [0, 1])
def test_cubic_spline_width_height_prescale(hm, inverse):
    """wh_scale (1/sqrt(hidden) of the coupling layer): evaluating raw parameters p with scale s equals evaluating the
    pre-multiplied parameters with scale 1, and the width / height gradients pick up the factor s."""
    gold = load_golden("functions_cubic")
    name = "cubic_inv_k8" if inverse else "cubic_fwd_k8"
    k, scale = 8, 0.25
    x = gold[name + "/x"].contiguous()
    p = torch.cat([gold[name + "/uw"], gold[name + "/uh"], gold[name + "/dl"], gold[name + "/dr"]], dim=-1).contiguous()
    p_raw = p.clone()
    p_raw[:, :2 * k] /= scale
    gy, gl = gold[name + "/gy"].contiguous(), gold[name + "/gl"].contiguous()
    outs = []
    for params, sc in ((p_raw.contiguous(), scale), (p, 1.0)):
        y, lad, gx, gp = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x), torch.empty_like(p)
        hm.hm_cubicspline_scaled(fptr(x), fptr(params), fptr(gy), fptr(gl), fptr(y), fptr(lad), fptr(gx), fptr(gp),
                                 ctypes.c_long(x.numel()), k, inverse, ctypes.c_float(sc))
        outs.append((y, lad, gx, gp))
    (y0, l0, gx0, gp0), (y1, l1, gx1, gp1) = outs
    assert (y0 - y1).abs().max() < 1e-5 and (l0 - l1).abs().max() < 1e-4
    assert (gx0 - gx1).abs().max() <= 1e-4 * max(1.0, gx1.abs().max().item())
    gs = max(1.0, gp1.abs().max().item())
    assert (gp0[:, :2 * k] - scale * gp1[:, :2 * k]).abs().max() <= 1e-4 * gs
    assert (gp0[:, 2 * k:] - gp1[:, 2 * k:]).abs().max() <= 1e-4 * gs
