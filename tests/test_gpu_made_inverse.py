"""GPU parity tests of the incremental autoregressive inverse (csrc/fc_made_inverse.cu, C ABI fc_made_inverse_*;
SURVEY 8(f) n2) against the D-pass inverse of the reference (flowcon/transforms/autoregressive/autoregressive.py:44-53;
oracle: restated._autoregressive) and against this package's own D-pass path."""
import pytest
import torch

from flowconductor_b200 import _cabi, graphs, made_inverse, transforms, workloads
from flowconductor_b200.transforms.base import InputOutsideDomain
from oracle import restated
from tests.helpers import assert_parity, golden_state, load_golden, parity_report

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def _d_pass(layer, z):
    made_inverse.ENABLED = False
    try:
        with torch.no_grad():
            return layer.inverse(z)
    finally:
        made_inverse.ENABLED = True


def _count(fn, name):
    _cabi.STATS.reset()
    out = fn()
    return out, _cabi.STATS.counts.get(name, 0)


@pytest.mark.parametrize("name", ["cfg3_small", "cfg1"])
def test_golden_models_take_the_incremental_kernel(dev, name):
    """The reference-generated fixtures: whole-flow inverse through the new kernel, at the reference tests' own 1e-3."""
    gold = load_golden(name)
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl)
    flow.load_state_dict(golden_state(gold), strict=True)
    flow = flow.to(dev).eval()
    entry = "fc_made_inverse_rqs" if name == "cfg3_small" else "fc_made_inverse_affine"
    with torch.no_grad():
        (xi, ladi), n = _count(lambda: flow._transform.inverse(gold["noise"].to(dev)), entry)
    assert n == sum(1 for l in wl["layers"] if l["kind"].startswith("maf_"))
    yfloor = max(1.0, gold["inv_y64"].abs().median().item())
    print(name, "inverse outputs:", parity_report(xi, gold["inv_y32"], gold["inv_y64"], 1e-5, yfloor))
    print(name, "inverse logabsdet:", parity_report(ladi, gold["inv_lad32"], gold["inv_lad64"], 1e-5, 1.0))
    assert_parity(xi, gold["inv_y32"], gold["inv_y64"], 1e-3, yfloor, name + " inverse outputs")
    assert_parity(ladi, gold["inv_lad32"], gold["inv_lad64"], 1e-2, 1.0, name + " inverse logabsdet")


@pytest.mark.parametrize("rows", [1, 33, 4096 + 5])
def test_cfg3_layer_matches_oracle_and_d_pass(dev, rows):
    """One full-size cfg 3 layer (D=16, H=256, K=16, trained-like weights): new kernel vs fp64 oracle D-pass inverse vs our
    D-pass path; then forward(inverse(z)) == z."""
    wl = workloads.get_workload("cfg3")
    flow = workloads.build_flow(wl, seed=0)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl, seed=1)
    flow.load_state_dict(state)
    specs = workloads.oracle_specs(wl)
    idx = [i for i, s in enumerate(specs) if s["kind"] == "maf_prq"][1]
    layer = flow._transform._transforms[idx].to(dev).eval()
    g = torch.Generator().manual_seed(11)
    z = torch.randn(rows, 16, generator=g) * 1.5
    z[0, :4] = torch.tensor([3.0, -3.0, 3.5, -7.0])  # tail boundary (inside) and beyond
    st32 = {k: v.float() for k, v in state.items()}
    st64 = {k: v.double() for k, v in state.items()}
    n_or = min(rows, 512)
    with torch.no_grad():
        r32, l32 = restated.apply_layer(st32, specs[idx], z[:n_or], inverse=True)
        r64, l64 = restated.apply_layer(st64, specs[idx], z[:n_or].double(), inverse=True)
        (x, lad), n = _count(lambda: layer.inverse(z.to(dev)), "fc_made_inverse_rqs")
        assert n == 1
        xd, ladd = _d_pass(layer, z.to(dev))
        zz, ladf = layer(x)
    print("rows %d outputs: %s" % (rows, parity_report(x[:n_or], r32, r64, 1e-5, 1.0)))
    print("rows %d logabsdet: %s" % (rows, parity_report(lad[:n_or], l32, l64, 1e-5, 1.0)))
    assert_parity(x[:n_or], r32, r64, 1e-3, 1.0, "incremental inverse outputs")
    assert_parity(lad[:n_or], l32, l64, 1e-2, 1.0, "incremental inverse logabsdet")
    # against our own D-pass path: both are fp32 evaluations of an ill-conditioned map (trained-like weights), so a few
    # elements in a thousand differ by more than 1e-3; the oracle comparison above is the parity statement
    dx, dl = (x - xd).abs().flatten(), (lad - ladd).abs()
    assert torch.quantile(dx, 0.99) < 1e-3 and dx.max() < 0.3 and torch.quantile(dl, 0.99) < 2e-2
    inside = (z.abs() <= 3.0).all(dim=1).to(dev)
    if bool(inside.any()):
        assert torch.quantile((zz - z.to(dev))[inside].abs().flatten(), 0.99) < 5e-3
        assert torch.quantile((ladf + lad)[inside].abs(), 0.99) < 5e-2
    assert torch.equal(x[0, 2:4].cpu(), z[0, 2:4])  # outside the tails: identity (rational_quadratic.py:38-39)


@pytest.mark.parametrize("features,hidden,blocks,bins,tails", [(5, 48, 1, 8, "linear"), (7, 64, 3, 10, "linear"),
                                                               (3, 20, 2, 5, None), (2, 8, 2, 16, "linear"),
                                                               (33, 96, 2, 8, "linear")])
def test_shapes_and_bin_counts(dev, features, hidden, blocks, bins, tails):
    torch.manual_seed(features * 100 + hidden)
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
        features, hidden, num_bins=bins, tails=tails, tail_bound=2.5, num_blocks=blocks).to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.3)
        z = torch.randn(257, features, device=dev)
        if tails is None:
            z = z.clamp(-1.15, 1.15)
        (x, lad), n = _count(lambda: layer.inverse(z), "fc_made_inverse_rqs")
        assert n == 1
        xd, ladd = _d_pass(layer, z)
        # the oracle's D-pass inverse (fp32 and fp64) with the same weights
        spec = {"kind": "maf_prq", "prefix": "", "num_bins": bins, "tails": tails, "tail_bound": 2.5, "num_blocks": blocks,
                "hidden_features": hidden}
        st32 = {k: v.detach().cpu() for k, v in layer.state_dict().items()}
        st64 = {k: (v.double() if v.is_floating_point() else v) for k, v in st32.items()}
        r32, l32 = restated.apply_layer(st32, spec, z.cpu(), inverse=True)
        r64, l64 = restated.apply_layer(st64, spec, z.cpu().double(), inverse=True)
    print("D=%d H=%d blocks=%d K=%d: outputs %s | D-pass path %s" % (
        features, hidden, blocks, bins, parity_report(x, r32, r64, 1e-5, 1.0), parity_report(xd, r32, r64, 1e-5, 1.0)))
    assert_parity(x, r32, r64, 1e-3, 1.0, "incremental inverse outputs")
    assert_parity(lad, l32, l64, 1e-2, 1.0, "incremental inverse logabsdet")
    assert (x - xd).abs().max() < 1e-3 and (lad - ladd).abs().max() < 1e-2


def test_affine_layer(dev):
    torch.manual_seed(5)
    layer = transforms.MaskedAffineAutoregressiveTransform(11, 40, num_blocks=2).to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.2)
        z = torch.randn(1000, 11, device=dev)
        (x, lad), n = _count(lambda: layer.inverse(z), "fc_made_inverse_affine")
        assert n == 1
        xd, ladd = _d_pass(layer, z)
        zz, ladf = layer(x)
    assert ((x - xd).abs() / xd.abs().clamp_min(1)).max() < 1e-3 and (lad - ladd).abs().max() < 1e-3
    assert ((zz - z).abs() / z.abs().clamp_min(1)).max() < 1e-3 and (ladf + lad).abs().max() < 1e-3


def test_domain_error_without_tails(dev):
    """tails=None: inputs outside [-1.2, 1.2] raise InputOutsideDomain, as the reference's first pass does
    (rational_quadratic.py:81-82 through autoregressive.py:595)."""
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(4, 16, num_bins=4, tails=None).to(dev).eval()
    z = torch.zeros(8, 4, device=dev)
    z[3, 2] = 1.5
    with torch.no_grad(), pytest.raises(InputOutsideDomain):
        layer.inverse(z)


def test_unsupported_nets_keep_the_d_pass_path(dev):
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
        4, 16, num_bins=4, tails="linear", use_residual_blocks=False).to(dev).eval()
    z = torch.randn(16, 4, device=dev)
    with torch.no_grad():
        (x, lad), n = _count(lambda: layer.inverse(z), "fc_made_inverse_rqs")
        zz, _ = layer(x)
    assert n == 0 and (zz - z).abs().max() < 1e-3


def test_weights_change_recompiles_and_graph_replays(dev):
    torch.manual_seed(9)
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
        6, 32, num_bins=8, tails="linear", tail_bound=3.0).to(dev).eval()
    z = torch.randn(300, 6, device=dev)
    with torch.no_grad():
        x0, _ = layer.inverse(z)
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.3)  # bumps the version counters: the cached program must be rebuilt
        x1, lad1 = layer.inverse(z)
        xd, _ = _d_pass(layer, z)
        assert (x1 - x0).abs().max() > 1e-3 and (x1 - xd).abs().max() < 1e-3
        gc = graphs.capture(layer.inverse, z)
        xg, ladg = gc(z)
    assert torch.equal(xg, x1) and torch.equal(ladg, lad1)
