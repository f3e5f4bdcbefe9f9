"""Pin the oracle (oracle/restated.py) against golden vectors written by the unmodified reference
(oracle/make_golden.py).  CPU only.  Because the restatement keeps the reference's operation order
and uses the same CPU PyTorch primitives, it must agree with the fp32 reference essentially to the
bit; the fp64 runs must agree to ~1e-13."""
import math

import pytest
import torch

from flowconductor_b200 import workloads
from oracle import locate, restated
from tests.helpers import golden_state, load_golden

F32_TOL = 2e-6   # restatement vs reference, fp32, absolute on O(1) values
F64_TOL = 1e-11


def _close(a, b, tol, what):
    assert a.shape == b.shape, what
    err = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
    assert err <= tol, "{}: max abs err {} > {}".format(what, err, tol)


@pytest.fixture(scope="module")
def fn_gold():
    return load_golden("functions")


def test_searchsorted_known_answer(fn_gold):
    # reference KAT: tests/utils/torchutils_test.py:81-91
    idx = restated.bin_index(fn_gold["searchsorted/knots"][None, :], fn_gold["searchsorted/x"])
    assert torch.equal(idx, fn_gold["searchsorted/idx"])
    assert torch.equal(idx, torch.arange(9))


RQ_CASES = ["rq_fwd_lin_k8", "rq_inv_lin_k8", "rq_fwd_lin_k16_id", "rq_inv_lin_k16_id", "rq_fwd_none_k5",
            "rq_inv_none_k5", "rq_fwd_lin_k10_b1", "rq_fwd_lin_k8_zero_id"]


def _run_rq(gold, name, dtype, requires_grad=False):
    k, lin, tb, inv, ident = gold[name + "/meta"].tolist()
    k = int(k)
    x = gold[name + "/x"].to(dtype).clone().requires_grad_(requires_grad)
    p = gold[name + "/params"].to(dtype).clone().requires_grad_(requires_grad)
    uw, uh, ud = p[..., :k], p[..., k:2 * k], p[..., 2 * k:]
    kw = dict(inverse=bool(inv), enable_identity_init=bool(ident))
    if lin:
        y, lad = restated.unconstrained_rational_quadratic_spline(x, uw, uh, ud, tails="linear", tail_bound=tb, **kw)
    else:
        y, lad = restated.rational_quadratic_spline(x, uw, uh, ud, **kw)
    return x, p, y, lad


@pytest.mark.parametrize("name", RQ_CASES)
def test_rq_spline_matches_reference(fn_gold, name):
    for dtype, tag, tol in ((torch.float32, "32", F32_TOL), (torch.float64, "64", F64_TOL)):
        _, _, y, lad = _run_rq(fn_gold, name, dtype)
        _close(y, fn_gold[name + "/y" + tag], tol, name + " outputs " + tag)
        _close(lad, fn_gold[name + "/lad" + tag], tol, name + " logabsdet " + tag)


def test_rq_identity_init_known_answer(fn_gold):
    # tests/transforms/splines/rational_quadratic_test.py:33-62: constrained spline, zero params (K+1
    # derivatives) + identity init => identity map with zero logabsdet, both directions, eps 1e-6
    k, shape = 10, (2, 3, 4)
    z = torch.zeros(*shape, k)
    zd = torch.zeros(*shape, k + 1)
    for inverse in (False, True):
        x = torch.rand(*shape, generator=torch.Generator().manual_seed(3))
        y, lad = restated.rational_quadratic_spline(x, z, z, zd, inverse=inverse, enable_identity_init=True)
        _close(y, x, 1e-6, "identity outputs")
        _close(lad, torch.zeros_like(lad), 1e-6, "identity logabsdet")
    # linear tails: the padded boundary derivative is 1.29499, not 1 (reference quirk), so
    # only points outside the interval and interior bins are exactly identity
    name = "rq_fwd_lin_k8_zero_id"
    x, _, y, lad = _run_rq(fn_gold, name, torch.float32)
    interior = x.abs() <= 3.0 * (1 - 2.0 / 8)
    outside = x.abs() > 3.0
    sel = interior | outside
    _close(y[sel], x[sel], 2e-6, "identity outputs (interior bins / tails)")
    _close(lad[sel], torch.zeros_like(lad[sel]), 2e-6, "identity logabsdet (interior bins / tails)")
    assert (y[~sel] - x[~sel]).abs().max() > 1e-3  # the quirk is reproduced


@pytest.mark.parametrize("name", ["rq_fwd_lin_k8", "rq_inv_lin_k8", "rq_fwd_lin_k16_id", "rq_fwd_none_k5"])
def test_rq_gradients_match_reference(fn_gold, name):
    for dtype, tag, tol in ((torch.float32, "32", 5e-5), (torch.float64, "64", 1e-9)):
        x, p, y, lad = _run_rq(fn_gold, name, dtype, requires_grad=True)
        gy, gl = fn_gold[name + "/gy"].to(dtype), fn_gold[name + "/gl"].to(dtype)
        gx, gp = torch.autograd.grad((y * gy).sum() + (lad * gl).sum(), [x, p])
        scale = max(1.0, fn_gold[name + "/gx" + tag].abs().max().item())
        _close(gx / scale, fn_gold[name + "/gx" + tag] / scale, tol, name + " grad x " + tag)
        scale = max(1.0, fn_gold[name + "/gp" + tag].abs().max().item())
        _close(gp / scale, fn_gold[name + "/gp" + tag] / scale, tol, name + " grad params " + tag)


def test_rq_domain_error():
    # tests/transforms/nonlinearities_test.py:60-76: constrained spline raises outside [0,1]
    p = torch.zeros(1, 1, 16)
    for bad in (-1.0, -0.1, 1.1, 2.0):
        with pytest.raises(restated.InputOutsideDomain):
            restated.rational_quadratic_spline(torch.tensor([[bad]]), p[..., :5], p[..., 5:10], p[..., 10:])


def test_affine_matches_reference(fn_gold):
    for dtype, tag, tol in ((torch.float32, "32", F32_TOL), (torch.float64, "64", F64_TOL)):
        x, p = fn_gold["affine/x"].to(dtype), fn_gold["affine/params"].to(dtype)
        for act in ("sigmoid2", "softplus_clamp3"):
            y, lad = restated.affine_elementwise(x, p, "blocked", act, inverse=False)
            _close(y, fn_gold["affine/blocked_%s_fwd_y%s" % (act, tag)], tol, act + " y")
            _close(lad, fn_gold["affine/blocked_%s_fwd_lad%s" % (act, tag)], tol * 10, act + " lad")
            yi, ladi = restated.affine_elementwise(x, p, "blocked", act, inverse=True)
            _close(yi, fn_gold["affine/blocked_%s_inv_y%s" % (act, tag)], tol * 100, act + " inv y")
            _close(ladi, -lad, 0.0, act + " inv lad")
        y, lad = restated.affine_elementwise(x, p, "interleaved", "softplus_eps", inverse=False)
        _close(y, fn_gold["affine/interleaved_fwd_y" + tag], tol, "maf y")
        _close(lad, fn_gold["affine/interleaved_fwd_lad" + tag], tol * 10, "maf lad")
        yi, _ = restated.affine_elementwise(x, p, "interleaved", "softplus_eps", inverse=True)
        _close(yi, fn_gold["affine/interleaved_inv_y" + tag], tol * 1000, "maf inv y")


@pytest.mark.parametrize("name", ["sos_n10", "sos_n3_wide"])
def test_sum_of_sigmoids_matches_reference(fn_gold, name):
    n = int(fn_gold[name + "/meta"][0])
    for dtype, tag, tol in ((torch.float32, "32", 1e-5), (torch.float64, "64", 1e-10)):
        x = fn_gold[name + "/x"].to(dtype).clone().requires_grad_(True)
        raw = fn_gold[name + "/params"].to(dtype).clone().requires_grad_(True)
        y, lad = restated.sos_forward(x, raw, n)
        ref_y = fn_gold[name + "/y" + tag]
        scale = max(1.0, ref_y.abs().max().item())
        _close(y / scale, ref_y / scale, tol, name + " y " + tag)
        _close(lad, fn_gold[name + "/lad" + tag], tol * 10, name + " lad " + tag)
        gy, gl = fn_gold[name + "/gy"].to(dtype), fn_gold[name + "/gl"].to(dtype)
        gx, gp = torch.autograd.grad((y * gy).sum() + (lad * gl).sum(), [x, raw])
        gscale = max(1.0, fn_gold[name + "/gx" + tag].abs().max().item())
        _close(gx / gscale, fn_gold[name + "/gx" + tag] / gscale, tol * 10, name + " gx " + tag)
        gscale = max(1.0, fn_gold[name + "/gp" + tag].abs().max().item())
        _close(gp / gscale, fn_gold[name + "/gp" + tag] / gscale, tol * 10, name + " gp " + tag)
        # numerical inverse (bisection + 2 Newton steps): reference tests use eps 1e-5 .. 1e-3
        xi, ladi = restated.sos_inverse(ref_y.to(dtype), raw.detach(), n)
        ref_xi = fn_gold[name + "/inv_x" + tag]
        iscale = max(1.0, ref_xi.abs().max().item())
        _close(xi / iscale, ref_xi / iscale, 1e-4 if tag == "32" else 1e-8, name + " inverse x " + tag)
        _close(ladi, fn_gold[name + "/inv_lad" + tag], 2e-3 if tag == "32" else 1e-7, name + " inverse lad " + tag)


MODELS = ["cfg1", "cfg2_small", "cfg3_small", "cfg4_small", "affine_coupling_small", "cond_prq_small",
          "maf_sos_small", "prq_coupling_notails_small", "prq_coupling_uncond_small", "plin_coupling_small",
          "maf_plin_small", "pquad_coupling_small", "maf_pquad_small", "pcubic_coupling_small", "maf_pcubic_small", "actnorm_maf_small"]


@pytest.mark.parametrize("name", MODELS)
def test_flow_matches_reference(name):
    gold = load_golden(name)
    wl = workloads.get_workload(name)
    specs = workloads.oracle_specs(wl)
    for dtype, tag, tol in ((torch.float32, "32", 2e-4), (torch.float64, "64", 1e-9)):
        state = golden_state(gold, dtype)
        x = gold["x"].to(dtype)
        ctx = gold["context"].to(dtype) if "context" in gold else None
        with torch.no_grad():
            y, lad = restated.composite(state, specs, x, ctx, inverse=False)
            lp = restated.flow_log_prob(state, specs, x, ctx)
        ys = max(1.0, gold["fwd_y" + tag].abs().max().item())
        ls = max(1.0, gold["fwd_lad" + tag].abs().max().item())
        _close(y / ys, gold["fwd_y" + tag] / ys, tol, name + " forward outputs " + tag)
        _close(lad / ls, gold["fwd_lad" + tag] / ls, tol, name + " forward logabsdet " + tag)
        _close(lp / ls, gold["log_prob" + tag] / ls, tol, name + " log_prob " + tag)
        yi, ladi = restated.composite(state, specs, gold["noise"].to(dtype), ctx, inverse=True)
        ys = max(1.0, gold["inv_y" + tag].abs().max().item())
        ls = max(1.0, gold["inv_lad" + tag].abs().max().item())
        inv_tol = tol if "sos" not in name and name != "cfg4_small" else max(tol, 1e-3 if tag == "32" else 1e-6)
        _close(yi.detach() / ys, gold["inv_y" + tag] / ys, inv_tol, name + " inverse outputs " + tag)
        _close(ladi.detach() / ls, gold["inv_lad" + tag] / ls, inv_tol * 10, name + " inverse logabsdet " + tag)


@pytest.mark.parametrize("name", ["cubic_fwd_k8", "cubic_inv_k8", "cubic_fwd_k5", "cubic_inv_k5"])
def test_cubic_spline_matches_reference(name):
    """restated.cubic_spline against the unmodified reference (functions_cubic.npz): groundwork for the last member of
    SURVEY 8f n3 — the oracle is pinned, the kernel comes next round."""
    gold = load_golden("functions_cubic")
    k, inverse = gold[name + "/meta"].tolist()
    for dtype, tag, tol in ((torch.float32, "32", 5e-5), (torch.float64, "64", 1e-10)):
        args = [gold[name + "/" + key].to(dtype).requires_grad_(True) for key in ("x", "uw", "uh", "dl", "dr")]
        y, lad = restated.cubic_spline(*args, inverse=bool(inverse))
        grads = torch.autograd.grad((y * gold[name + "/gy"].to(dtype)).sum() + (lad * gold[name + "/gl"].to(dtype)).sum(),
                                    args)
        _close(y, gold[name + "/y" + tag], tol, name + " y " + tag)
        ls = max(1.0, gold[name + "/lad" + tag].abs().max().item())
        _close(lad / ls, gold[name + "/lad" + tag] / ls, tol * 10, name + " lad " + tag)
        for got, key in zip(grads, ("gx", "gw", "gh", "gdl", "gdr")):
            ref = gold[name + "/" + key + tag]
            gs = max(1.0, ref.abs().max().item())
            _close(got / gs, ref / gs, tol * 50, name + " " + key + " " + tag)


QUADRATIC_CASES = ["quad_fwd_k8", "quad_inv_k8", "quad_fwd_tails_k10", "quad_inv_tails_k10", "quad_fwd_k5"]


@pytest.mark.parametrize("name", QUADRATIC_CASES)
def test_quadratic_spline_matches_reference(name):
    """restated.quadratic_spline / unconstrained_quadratic_spline against the unmodified reference
    (functions_quadratic.npz): outputs, log-dets and autograd gradients, fp32 and fp64."""
    gold = load_golden("functions_quadratic")
    k, has_tails, tb, inverse = gold[name + "/meta"].tolist()
    for dtype, tag, tol in ((torch.float32, "32", 2e-5), (torch.float64, "64", 1e-11)):
        x = gold[name + "/x"].to(dtype).requires_grad_(True)
        uw = gold[name + "/uw"].to(dtype).requires_grad_(True)
        uh = gold[name + "/uh"].to(dtype).requires_grad_(True)
        if has_tails:
            y, lad = restated.unconstrained_quadratic_spline(x, uw, uh, inverse=bool(inverse), tail_bound=tb)
        else:
            y, lad = restated.quadratic_spline(x, uw, uh, inverse=bool(inverse))
        grads = torch.autograd.grad((y * gold[name + "/gy"].to(dtype)).sum() + (lad * gold[name + "/gl"].to(dtype)).sum(),
                                    [x, uw, uh])
        _close(y, gold[name + "/y" + tag], tol, name + " y " + tag)
        _close(lad, gold[name + "/lad" + tag], tol * 10, name + " lad " + tag)
        for got, key in zip(grads, ("gx", "gw", "gh")):
            ref = gold[name + "/" + key + tag]
            gs = max(1.0, ref.abs().max().item())
            _close(got / gs, ref / gs, tol * 20, name + " " + key + " " + tag)


LINEAR_CASES = ["lin_fwd_k8", "lin_inv_k8", "lin_fwd_tails_k10", "lin_inv_tails_k10", "lin_fwd_k5"]


@pytest.mark.parametrize("name", LINEAR_CASES)
def test_linear_spline_matches_reference(name):
    """restated.linear_spline / unconstrained_linear_spline against the unmodified reference (functions_linear.npz):
    outputs, log-dets and autograd gradients, fp32 and fp64."""
    gold = load_golden("functions_linear")
    k, has_tails, tb, inverse = gold[name + "/meta"].tolist()
    for dtype, tag, tol in ((torch.float32, "32", 1e-5), (torch.float64, "64", 1e-12)):
        x = gold[name + "/x"].to(dtype).requires_grad_(True)
        u = gold[name + "/params"].to(dtype).requires_grad_(True)
        if has_tails:
            y, lad = restated.unconstrained_linear_spline(x, u, inverse=bool(inverse), tail_bound=tb, tails="linear")
        else:
            y, lad = restated.linear_spline(x, u, inverse=bool(inverse))
        gx, gu = torch.autograd.grad((y * gold[name + "/gy"].to(dtype)).sum() + (lad * gold[name + "/gl"].to(dtype)).sum(),
                                     [x, u])
        _close(y, gold[name + "/y" + tag], tol, name + " y " + tag)
        _close(lad, gold[name + "/lad" + tag], tol * 10, name + " lad " + tag)
        gs = max(1.0, gold[name + "/gx" + tag].abs().max().item())
        _close(gx / gs, gold[name + "/gx" + tag] / gs, tol * 10, name + " gx " + tag)
        gs = max(1.0, gold[name + "/gp" + tag].abs().max().item())
        _close(gu / gs, gold[name + "/gp" + tag] / gs, tol * 10, name + " gp " + tag)


@pytest.mark.parametrize("name", ["cfg1", "cfg3_small", "cfg2_small", "plin_coupling_small", "maf_plin_small",
                                  "prq_coupling_uncond_small", "pquad_coupling_small", "maf_pquad_small"])
def test_flow_parameter_gradients_match_reference(name):
    gold = load_golden(name)
    wl = workloads.get_workload(name)
    specs = workloads.oracle_specs(wl)
    dtype, tag = torch.float64, "64"
    state = golden_state(gold, dtype)
    names = [k[len("grad64/"):] for k in gold if k.startswith("grad64/")]
    for n in names:
        state[n] = state[n].clone().requires_grad_(True)
    loss = -restated.flow_log_prob(state, specs, gold["x"].to(dtype), None).mean()
    _close(loss, gold["loss64"], 1e-9, "loss")
    grads = torch.autograd.grad(loss, [state[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        ref = gold["grad64/" + n]
        g = torch.zeros_like(ref) if g is None else g
        _close(g, ref, 1e-8 * max(1.0, ref.abs().max().item()), "grad " + n)


@pytest.mark.skipif(not locate.have_reference(), reason="reference tree not present (GPU box)")
def test_restatement_matches_live_reference():
    """With /root/reference importable, compare on fresh random inputs (not only the stored vectors)."""
    locate.import_reference()
    from flowcon.transforms.splines import rational_quadratic as ref_rq

    g = torch.Generator().manual_seed(7)
    for k, tb, inv, ident in ((8, 3.0, False, False), (8, 3.0, True, False), (16, 3.0, False, True), (4, 1.0, True, True)):
        x = torch.randn(257, 9, generator=g) * tb * 0.6
        p = torch.randn(257, 9, 3 * k - 1, generator=g) * 2
        args = (p[..., :k].clone(), p[..., k:2 * k].clone(), p[..., 2 * k:].clone())
        ry, rl = ref_rq.unconstrained_rational_quadratic_spline(x, *args, inverse=inv, tails="linear", tail_bound=tb,
                                                                enable_identity_init=ident)
        oy, ol = restated.unconstrained_rational_quadratic_spline(x, *args, inverse=inv, tails="linear",
                                                                  tail_bound=tb, enable_identity_init=ident)
        _close(oy, ry, F32_TOL, "live rq y")
        _close(ol, rl, F32_TOL, "live rq lad")


@pytest.mark.skipif(not locate.have_reference(), reason="reference tree not present")
def test_standalone_sum_of_sigmoids_state_dict_matches_reference():
    """the stand-alone SumOfSigmoids keeps the reference's state_dict keys (`extended_softplus.shift` as a
    sub-module parameter, the frozen `log_scale_postact`), so a reference checkpoint loads with strict=True."""
    locate.import_reference()
    from flowcon.transforms.adaptive_sigmoids import SumOfSigmoids as RefSoS

    from flowconductor_b200.transforms import SumOfSigmoids

    torch.manual_seed(0)
    ref = RefSoS(features=5, n_sigmoids=7)
    ours = SumOfSigmoids(features=5, n_sigmoids=7)
    assert set(ours.state_dict().keys()) == set(ref.state_dict().keys())
    ours.load_state_dict(ref.state_dict(), strict=True)
    assert torch.equal(ours.get_raw_params(), ref.get_raw_params())
