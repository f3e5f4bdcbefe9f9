"""TEST INFRASTRUCTURE ONLY — generate the committed golden vectors under tests/golden/ by running
the UNMODIFIED FlowConductor reference (imported from /root/reference through oracle/stubs).

Run in the build container:   python oracle/make_golden.py
The reference tree does not exist on the GPU box; tests there use the files this script wrote.

Files (all numpy .npz, float32 inputs; reference outputs in fp32 and fp64):
  tests/golden/functions.npz   function-level cases for a1-a4, a7, a9, a11-a13 + RQ gradients
  tests/golden/<workload>.npz  model-level cases: state_dict, inputs, log_prob, forward/inverse
                               outputs for the reduced-size twins in flowconductor_b200/workloads.py
"""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import locate  # noqa: E402

flowcon = locate.import_reference()

import torch.nn.functional as F  # noqa: E402
from flowcon import distributions, flows, transforms  # noqa: E402
from flowcon.nn import nets  # noqa: E402
from flowcon.transforms.adaptive_sigmoids import SumOfSigmoids  # noqa: E402
from flowcon.transforms.splines import rational_quadratic as ref_rq  # noqa: E402
from flowcon.utils import torchutils as ref_tu  # noqa: E402

from flowconductor_b200 import workloads  # noqa: E402  (plain data only)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def np32(t):
    return t.detach().to(torch.float32).numpy()


def np64(t):
    return t.detach().to(torch.float64).numpy()


# ------------------------------------------------------------------------------------------------
# function-level cases
# ------------------------------------------------------------------------------------------------
def rq_case(out, name, g, n, d, k, tails, tail_bound, inverse, identity_init, scale, with_grad=False):
    """One RQ-spline case on a [n, d] batch with contiguous raw params [n, d, P]."""
    p = 3 * k - 1 if tails == "linear" else 3 * k + 1
    params = torch.randn(n, d, p, generator=g) * scale
    if tails == "linear":
        x = torch.randn(n, d, generator=g) * (tail_bound / 2.0)
        # exercise the boundary: exact +-tail_bound (inside) and points just outside
        x[0, 0] = tail_bound
        x[0, -1] = -tail_bound
        x[1, 0] = tail_bound * 1.0001
        x[1, -1] = -tail_bound * 1.5
    else:
        lo, hi = (0.0, 1.0)
        x = torch.rand(n, d, generator=g) * (hi - lo) + lo
        x[0, 0] = lo
        x[0, -1] = hi
    res = {}
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        xx = x.to(dt).clone().requires_grad_(with_grad)
        pp = params.to(dt).clone().requires_grad_(with_grad)
        uw, uh, ud = pp[..., :k], pp[..., k:2 * k], pp[..., 2 * k:]
        kw = dict(inverse=inverse, enable_identity_init=identity_init)
        if tails == "linear":
            y, lad = ref_rq.unconstrained_rational_quadratic_spline(xx, uw, uh, ud, tails="linear",
                                                                    tail_bound=tail_bound, **kw)
        else:
            y, lad = ref_rq.rational_quadratic_spline(xx, uw, uh, ud, **kw)
        res["y" + tag], res["lad" + tag] = y, lad
        if with_grad:
            gy = torch.randn(n, d, generator=torch.Generator().manual_seed(77)).to(dt)
            gl = torch.randn(n, d, generator=torch.Generator().manual_seed(78)).to(dt)
            gx, gp = torch.autograd.grad((y * gy).sum() + (lad * gl).sum(), [xx, pp])
            res["gx" + tag], res["gp" + tag] = gx, gp
            if tag == "32":
                out[name + "/gy"] = np32(gy)
                out[name + "/gl"] = np32(gl)
    out[name + "/x"] = np32(x)
    out[name + "/params"] = np32(params)
    out[name + "/meta"] = np.array([k, 1 if tails == "linear" else 0, tail_bound, int(inverse), int(identity_init)],
                                   dtype=np.float64)
    for key, val in res.items():
        out[name + "/" + key] = np32(val) if key.endswith("32") else np64(val)
    # bin indices of the fp32 reference run, for the bin-parity check
    with torch.no_grad():
        uw, uh = params[..., :k], params[..., k:2 * k]
        lo, hi = (-tail_bound, tail_bound) if tails == "linear" else (0.0, 1.0)
        sizes = F.softmax(uh if inverse else uw, dim=-1)
        sizes = 1e-3 + (1 - 1e-3 * k) * sizes
        cum = F.pad(torch.cumsum(sizes, -1), (1, 0)) * (hi - lo) + lo
        cum[..., 0], cum[..., -1] = lo, hi
        idx = ref_tu.searchsorted(cum, x.clamp(lo, hi))
    out[name + "/bin"] = idx.numpy().astype(np.int64)
    out[name + "/knots"] = np32(cum)


def make_functions():
    out = {}
    g = torch.Generator().manual_seed(20261018)
    # a3 known answer (tests/utils/torchutils_test.py:81-91)
    bins = torch.linspace(0, 1, 10)
    inputs = torch.linspace(0, 1, 10)[:-1] + 0.05
    out["searchsorted/knots"] = np32(torch.linspace(0, 1, 10))
    out["searchsorted/x"] = np32(inputs)
    out["searchsorted/idx"] = ref_tu.searchsorted(bins[None, :].clone(), inputs).numpy()
    # a1/a2
    rq_case(out, "rq_fwd_lin_k8", g, 96, 32, 8, "linear", 3.0, False, False, 2.0, with_grad=True)
    rq_case(out, "rq_inv_lin_k8", g, 96, 32, 8, "linear", 3.0, True, False, 2.0, with_grad=True)
    rq_case(out, "rq_fwd_lin_k16_id", g, 64, 16, 16, "linear", 3.0, False, True, 1.5, with_grad=True)
    rq_case(out, "rq_inv_lin_k16_id", g, 64, 16, 16, "linear", 3.0, True, True, 1.5)
    rq_case(out, "rq_fwd_none_k5", g, 50, 7, 5, None, 1.0, False, False, 1.0, with_grad=True)
    rq_case(out, "rq_inv_none_k5", g, 50, 7, 5, None, 1.0, True, False, 1.0)
    rq_case(out, "rq_fwd_lin_k10_b1", g, 33, 3, 10, "linear", 1.0, False, False, 3.0)
    rq_case(out, "rq_fwd_lin_k8_zero_id", g, 16, 4, 8, "linear", 3.0, False, True, 0.0)  # identity-init KAT
    # a7 / a9 affine
    n, d = 64, 12
    x = torch.randn(n, d, generator=g)
    pa = torch.randn(n, 2 * d, generator=g) * 2
    out["affine/x"], out["affine/params"] = np32(x), np32(pa)
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        xx, pp = x.to(dt), pa.to(dt)
        for act_name, act in (("sigmoid2", transforms.AffineCouplingTransform.DEFAULT_SCALE_ACTIVATION),
                              ("softplus_clamp3", transforms.AffineCouplingTransform.GENERAL_SCALE_ACTIVATION)):
            shift, raw = pp[:, :d], pp[:, d:]
            scale = act(raw)
            out["affine/blocked_%s_fwd_y%s" % (act_name, tag)] = (np32 if tag == "32" else np64)(xx * scale + shift)
            out["affine/blocked_%s_fwd_lad%s" % (act_name, tag)] = (np32 if tag == "32" else np64)(
                torch.log(scale).sum(1))
            out["affine/blocked_%s_inv_y%s" % (act_name, tag)] = (np32 if tag == "32" else np64)((xx - shift) / scale)
        p3 = pp.view(n, d, 2)
        scale = F.softplus(p3[..., 0]) + 1e-3
        out["affine/interleaved_fwd_y" + tag] = (np32 if tag == "32" else np64)(scale * xx + p3[..., 1])
        out["affine/interleaved_fwd_lad" + tag] = (np32 if tag == "32" else np64)(torch.log(scale).sum(1))
        out["affine/interleaved_inv_y" + tag] = (np32 if tag == "32" else np64)((xx - p3[..., 1]) / scale)
    # a11-a13 sum of sigmoids
    for name, n, d, ns, xs in (("sos_n10", 64, 8, 10, 3.0), ("sos_n3_wide", 32, 5, 3, 60.0)):
        x = torch.randn(n, d, generator=g) * xs
        raw = torch.randn(n, d, 3 * ns + 1, generator=g) * 1.5
        out[name + "/x"], out[name + "/params"] = np32(x), np32(raw)
        out[name + "/meta"] = np.array([ns], dtype=np.float64)
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            conv = np32 if tag == "32" else np64
            xx = x.to(dt).clone().requires_grad_(True)
            rr = raw.to(dt).clone().requires_grad_(True)
            t = SumOfSigmoids(features=d, n_sigmoids=ns, raw_params=rr)
            y, lad = t(xx)
            out[name + "/y" + tag], out[name + "/lad" + tag] = conv(y), conv(lad)
            gy = torch.randn(n, d, generator=torch.Generator().manual_seed(5)).to(dt)
            gl = torch.randn(n, generator=torch.Generator().manual_seed(6)).to(dt)
            gx, gp = torch.autograd.grad((y * gy).sum() + (lad * gl).sum(), [xx, rr])
            out[name + "/gx" + tag], out[name + "/gp" + tag] = conv(gx), conv(gp)
            if tag == "32":
                out[name + "/gy"], out[name + "/gl"] = np32(gy), np32(gl)
            with torch.no_grad():
                t2 = SumOfSigmoids(features=d, n_sigmoids=ns, raw_params=raw.to(dt))
            xi, ladi = t2.inverse(y.detach())
            out[name + "/inv_x" + tag], out[name + "/inv_lad" + tag] = conv(xi), conv(ladi)
    np.savez_compressed(os.path.join(GOLDEN, "functions.npz"), **out)
    print("functions.npz:", len(out), "arrays")


# ------------------------------------------------------------------------------------------------
def make_linear_functions():
    """functions_linear.npz: the reference's linear_spline / unconstrained_linear_spline (splines/linear.py) forward,
    inverse and autograd gradients, fp32 and fp64, per-element outputs."""
    from flowcon.transforms.splines import linear as ref_lin

    out = {}
    g = torch.Generator().manual_seed(77)
    for name, k, tails, tb, inverse in [("lin_fwd_k8", 8, None, 1.0, False), ("lin_inv_k8", 8, None, 1.0, True),
                                        ("lin_fwd_tails_k10", 10, "linear", 3.0, False),
                                        ("lin_inv_tails_k10", 10, "linear", 3.0, True),
                                        ("lin_fwd_k5", 5, None, 1.0, False)]:
        n = 512
        u = torch.randn(n, k, generator=g) * 1.5
        if tails is None:
            x = torch.rand(n, generator=g)
            x[:4] = torch.tensor([0.0, 1.0, 0.5, 1.0 / k])  # domain ends and a bin edge
        else:
            x = torch.randn(n, generator=g) * 2.5
            x[:4] = torch.tensor([-tb, tb, 0.0, 5.0])
        gy, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)
        out[name + "/meta"] = np.array([k, 0 if tails is None else 1, tb, 1 if inverse else 0], dtype=np.float64)
        out[name + "/x"], out[name + "/params"] = np32(x), np32(u)
        out[name + "/gy"], out[name + "/gl"] = np32(gy), np32(gl)
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            conv = np32 if tag == "32" else np64
            xx = x.to(dt).clone().requires_grad_(True)
            uu = u.to(dt).clone().requires_grad_(True)
            if tails is None:
                y, lad = ref_lin.linear_spline(xx, uu, inverse=inverse)
            else:
                y, lad = ref_lin.unconstrained_linear_spline(xx, uu, inverse=inverse, tail_bound=tb, tails=tails)
            gx, gu = torch.autograd.grad((y * gy.to(dt)).sum() + (lad * gl.to(dt)).sum(), [xx, uu])
            out[name + "/y" + tag], out[name + "/lad" + tag] = conv(y), conv(lad)
            out[name + "/gx" + tag], out[name + "/gp" + tag] = conv(gx), conv(gu)
    np.savez_compressed(os.path.join(GOLDEN, "functions_linear.npz"), **out)
    print("functions_linear.npz:", len(out), "arrays")


def make_quadratic_functions():
    """functions_quadratic.npz: the reference's quadratic_spline / unconstrained_quadratic_spline
    (splines/quadratic.py), forward, inverse and autograd gradients, fp32 and fp64, per-element outputs."""
    from flowcon.transforms.splines import quadratic as ref_q

    out = {}
    g = torch.Generator().manual_seed(78)
    for name, k, tails, tb, inverse in [("quad_fwd_k8", 8, None, 1.0, False), ("quad_inv_k8", 8, None, 1.0, True),
                                        ("quad_fwd_tails_k10", 10, "linear", 3.0, False),
                                        ("quad_inv_tails_k10", 10, "linear", 3.0, True),
                                        ("quad_fwd_k5", 5, None, 1.0, False)]:
        n = 512
        uw = torch.randn(n, k, generator=g) * 1.5
        uh = torch.randn(n, k + 1 if tails is None else k - 1, generator=g) * 1.5
        if tails is None:
            x = torch.rand(n, generator=g)
            x[:3] = torch.tensor([0.0, 1.0, 0.5])
        else:
            x = torch.randn(n, generator=g) * 2.5
            x[:4] = torch.tensor([-tb, tb, 0.0, 5.0])
        gy, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)
        out[name + "/meta"] = np.array([k, 0 if tails is None else 1, tb, 1 if inverse else 0], dtype=np.float64)
        out[name + "/x"], out[name + "/uw"], out[name + "/uh"] = np32(x), np32(uw), np32(uh)
        out[name + "/gy"], out[name + "/gl"] = np32(gy), np32(gl)
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            conv = np32 if tag == "32" else np64
            xx = x.to(dt).clone().requires_grad_(True)
            ww = uw.to(dt).clone().requires_grad_(True)
            hh = uh.to(dt).clone().requires_grad_(True)
            if tails is None:
                y, lad = ref_q.quadratic_spline(xx, ww, hh, inverse=inverse)
            else:
                y, lad = ref_q.unconstrained_quadratic_spline(xx, ww, hh, inverse=inverse, tail_bound=tb, tails=tails)
            gx, gw, gh = torch.autograd.grad((y * gy.to(dt)).sum() + (lad * gl.to(dt)).sum(), [xx, ww, hh])
            out[name + "/y" + tag], out[name + "/lad" + tag] = conv(y), conv(lad)
            out[name + "/gx" + tag], out[name + "/gw" + tag], out[name + "/gh" + tag] = conv(gx), conv(gw), conv(gh)
    np.savez_compressed(os.path.join(GOLDEN, "functions_quadratic.npz"), **out)
    print("functions_quadratic.npz:", len(out), "arrays")


def make_cubic_functions():
    """functions_cubic.npz: the reference's cubic_spline (splines/cubic.py) forward and inverse with autograd
    gradients, fp32 and fp64 (groundwork: the oracle restatement is pinned, no kernel consumes it yet)."""
    from flowcon.transforms.splines import cubic as ref_c

    out = {}
    g = torch.Generator().manual_seed(79)
    for name, k, inverse in [("cubic_fwd_k8", 8, False), ("cubic_inv_k8", 8, True), ("cubic_fwd_k5", 5, False),
                             ("cubic_inv_k5", 5, True)]:
        n = 512
        uw, uh = torch.randn(n, k, generator=g), torch.randn(n, k, generator=g)
        dl, dr = torch.randn(n, 1, generator=g), torch.randn(n, 1, generator=g)
        x = torch.rand(n, generator=g) * 0.998 + 0.001
        gy, gl = torch.randn(n, generator=g), torch.randn(n, generator=g)
        out[name + "/meta"] = np.array([k, 1 if inverse else 0], dtype=np.float64)
        for key, t in (("x", x), ("uw", uw), ("uh", uh), ("dl", dl), ("dr", dr), ("gy", gy), ("gl", gl)):
            out[name + "/" + key] = np32(t)
        for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
            conv = np32 if tag == "32" else np64
            args = [t.to(dt).clone().requires_grad_(True) for t in (x, uw, uh, dl, dr)]
            y, lad = ref_c.cubic_spline(*args, inverse=inverse)
            grads = torch.autograd.grad((y * gy.to(dt)).sum() + (lad * gl.to(dt)).sum(), args)
            out[name + "/y" + tag], out[name + "/lad" + tag] = conv(y), conv(lad)
            for key, gr in zip(("gx", "gw", "gh", "gdl", "gdr"), grads):
                out[name + "/" + key + tag] = conv(gr)
    np.savez_compressed(os.path.join(GOLDEN, "functions_cubic.npz"), **out)
    print("functions_cubic.npz:", len(out), "arrays")


# model-level cases
# ------------------------------------------------------------------------------------------------
def build_reference_flow(wl, seed=0):
    """The workload dict instantiated with the REFERENCE's classes."""
    torch.manual_seed(seed)
    features, ctx = wl["features"], wl.get("context_features")
    layers = []
    for layer in wl["layers"]:
        kind = layer["kind"]
        if kind == "permutation":
            layers.append(transforms.RandomPermutation(features) if layer["mode"] == "random"
                          else transforms.ReversePermutation(features))
        elif kind == "actnorm":
            layers.append(transforms.ActNorm(features))
        elif kind in ("prq_coupling", "affine_coupling"):
            h, b = layer["hidden_features"], layer["num_blocks"]
            create = lambda i, o, h=h, b=b: nets.ResidualNet(i, o, hidden_features=h, num_blocks=b)  # noqa: E731
            mask = workloads.make_mask(features, layer["mask"])
            if kind == "prq_coupling":
                layers.append(transforms.PiecewiseRationalQuadraticCouplingTransform(
                    mask, create, num_bins=layer["num_bins"], tails=layer["tails"], tail_bound=layer["tail_bound"],
                    apply_unconditional_transform=layer.get("unconditional", False)))
            else:
                act = {"sigmoid2": transforms.AffineCouplingTransform.DEFAULT_SCALE_ACTIVATION,
                       "softplus_clamp3": transforms.AffineCouplingTransform.GENERAL_SCALE_ACTIVATION}[
                    layer["scale_activation"]]
                layers.append(transforms.AffineCouplingTransform(mask, create, scale_activation=act))
        elif kind == "plin_coupling":
            h, b = layer["hidden_features"], layer["num_blocks"]
            create = lambda i, o, h=h, b=b: nets.ResidualNet(i, o, hidden_features=h, num_blocks=b)  # noqa: E731
            layers.append(transforms.PiecewiseLinearCouplingTransform(
                workloads.make_mask(features, layer["mask"]), create, num_bins=layer["num_bins"], tails=layer["tails"],
                tail_bound=layer["tail_bound"], apply_unconditional_transform=layer.get("unconditional", False)))
        elif kind == "pquad_coupling":
            h, b = layer["hidden_features"], layer["num_blocks"]
            create = lambda i, o, h=h, b=b: nets.ResidualNet(i, o, hidden_features=h, num_blocks=b)  # noqa: E731
            layers.append(transforms.PiecewiseQuadraticCouplingTransform(
                workloads.make_mask(features, layer["mask"]), create, num_bins=layer["num_bins"], tails=layer["tails"],
                tail_bound=layer["tail_bound"], apply_unconditional_transform=layer.get("unconditional", False)))
        elif kind == "pcubic_coupling":
            h, b = layer["hidden_features"], layer["num_blocks"]
            create = lambda i, o, h=h, b=b: nets.ResidualNet(i, o, hidden_features=h, num_blocks=b)  # noqa: E731
            layers.append(transforms.PiecewiseCubicCouplingTransform(
                workloads.make_mask(features, layer["mask"]), create, num_bins=layer["num_bins"], tails=layer["tails"],
                tail_bound=layer["tail_bound"], apply_unconditional_transform=layer.get("unconditional", False)))
        elif kind == "maf_pquad":
            layers.append(transforms.MaskedPiecewiseQuadraticAutoregressiveTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                num_bins=layer["num_bins"], num_blocks=layer["num_blocks"], tails=layer["tails"],
                tail_bound=layer["tail_bound"]))
        elif kind == "maf_plin":
            layers.append(transforms.MaskedPiecewiseLinearAutoregressiveTransform(
                num_bins=layer["num_bins"], features=features, hidden_features=layer["hidden_features"],
                context_features=ctx, num_blocks=layer["num_blocks"]))
        elif kind == "maf_pcubic":
            layers.append(transforms.MaskedPiecewiseCubicAutoregressiveTransform(
                num_bins=layer["num_bins"], features=features, hidden_features=layer["hidden_features"],
                context_features=ctx, num_blocks=layer["num_blocks"]))
        elif kind == "maf_affine":
            layers.append(transforms.MaskedAffineAutoregressiveTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                num_blocks=layer["num_blocks"]))
        elif kind == "maf_prq":
            layers.append(transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                num_bins=layer["num_bins"], tails=layer["tails"], tail_bound=layer["tail_bound"],
                num_blocks=layer["num_blocks"]))
        elif kind == "maf_sos":
            layers.append(transforms.MaskedSumOfSigmoidsTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                n_sigmoids=layer["n_sigmoids"], num_blocks=layer["num_blocks"]))
        elif kind == "cond_sos":
            layers.append(transforms.ConditionalSumOfSigmoidsTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                n_sigmoids=layer["n_sigmoids"], num_blocks=layer["num_blocks"]))
        elif kind == "cond_prq":
            layers.append(transforms.ConditionalPiecewiseRationalQuadraticTransform(
                features=features, hidden_features=layer["hidden_features"], context_features=ctx,
                num_bins=layer["num_bins"], tails=layer["tails"], tail_bound=layer["tail_bound"],
                num_blocks=layer["num_blocks"]))
        else:
            raise ValueError(kind)
    return flows.Flow(transforms.CompositeTransform(layers), distributions.StandardNormal([features]))


def make_model(name, with_grad=False, x_scale=1.0, uniform01=False, batch=None):
    wl = workloads.get_workload(name)
    if batch is not None:
        wl["batch"] = batch
    flow = build_reference_flow(wl)
    state = {k: v.clone() for k, v in flow.state_dict().items()}
    workloads.trained_like_(state, wl)
    flow.load_state_dict(state)
    n, d, ctx = wl["batch"], wl["features"], wl.get("context_features")
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(n, d, generator=g) if uniform01 else torch.randn(n, d, generator=g) * x_scale
    c = torch.randn(n, ctx, generator=torch.Generator().manual_seed(4321)) if ctx else None
    noise = torch.randn(n, d, generator=torch.Generator().manual_seed(999))
    if uniform01:
        noise = torch.rand(n, d, generator=torch.Generator().manual_seed(999))
    out = {"x": np32(x), "noise": np32(noise)}
    if c is not None:
        out["context"] = np32(c)
    for k, v in state.items():
        out["state/" + k] = v.numpy()
    for dt, tag in ((torch.float32, "32"), (torch.float64, "64")):
        conv = np32 if tag == "32" else np64
        f = build_reference_flow(wl).to(dt)
        f.load_state_dict({k: (v.to(dt) if v.is_floating_point() else v) for k, v in state.items()})
        f.eval()  # ActNorm would otherwise re-initialise itself from the first batch (normalization.py:177-178)
        xx = x.to(dt)
        cc = c.to(dt) if c is not None else None
        with torch.no_grad():
            z, lad = f._transform(xx, context=cc)
            out["fwd_y" + tag], out["fwd_lad" + tag] = conv(z), conv(lad)
            out["log_prob" + tag] = conv(f.log_prob(xx, context=cc))
            xi, ladi = f._transform.inverse(noise.to(dt), context=cc)
            out["inv_y" + tag], out["inv_lad" + tag] = conv(xi), conv(ladi)
        if with_grad:
            f.zero_grad()
            loss = -f.log_prob(xx, context=cc).mean()
            loss.backward()
            out["loss" + tag] = conv(loss)
            for pn, p in f.named_parameters():
                out["grad%s/%s" % (tag, pn)] = conv(p.grad)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(name + ".npz:", len(out), "arrays,", os.path.getsize(os.path.join(GOLDEN, name + ".npz")) // 1024, "KiB")


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(4)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-uncond":  # add one fixture without rewriting the others
        make_model("prq_coupling_uncond_small", with_grad=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-actnorm":
        make_model("actnorm_maf_small", with_grad=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-cubic-functions":
        make_cubic_functions()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-cubic-models":
        make_model("pcubic_coupling_small", with_grad=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-maf-cubic":
        make_model("maf_pcubic_small", with_grad=True, uniform01=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-quadratic-functions":
        make_quadratic_functions()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-quadratic-models":
        make_model("pquad_coupling_small", with_grad=True)
        make_model("maf_pquad_small", with_grad=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "--only-linear":
        make_linear_functions()
        make_model("plin_coupling_small", with_grad=True)
        make_model("maf_plin_small", with_grad=True, uniform01=True)
        sys.exit(0)
    make_functions()
    make_model("cfg1", with_grad=True, batch=2048)
    make_model("cfg2_small", with_grad=True)
    make_model("cfg3_small", with_grad=True)
    make_model("cfg4_small", with_grad=True)
    make_model("affine_coupling_small", with_grad=True)
    make_model("cond_prq_small", with_grad=True)
    make_model("maf_sos_small")
    make_model("prq_coupling_notails_small", uniform01=True)
    make_model("prq_coupling_uncond_small", with_grad=True)
    make_model("actnorm_maf_small", with_grad=True)
