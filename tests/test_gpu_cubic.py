"""GPU parity of the cubic spline layer (fc_cubicspline_apply / fc_cubicspline_backward) against golden vectors
generated from the unmodified reference (flowcon/transforms/splines/cubic.py) and against the fp64 oracle."""
import pytest
import torch

from flowconductor_b200 import transforms
from flowconductor_b200.nn import nets
from oracle import restated
from tests.helpers import assert_parity, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", ["cubic_fwd_k8", "cubic_inv_k8", "cubic_fwd_k5", "cubic_inv_k5"])
def test_cubic_spline_kernels(dev, name):
    gold = load_golden("functions_cubic")
    k, inverse = gold[name + "/meta"].tolist()
    args = [gold[name + "/" + key].to(dev).requires_grad_(True) for key in ("x", "uw", "uh", "dl", "dr")]
    y, lad = transforms.cubic_spline(*args, inverse=bool(inverse))
    assert_parity(y, gold[name + "/y32"], gold[name + "/y64"], 1e-5, 1.0, name + " y")
    assert_parity(lad, gold[name + "/lad32"], gold[name + "/lad64"], 1e-5, 1.0, name + " lad")
    grads = torch.autograd.grad((y * gold[name + "/gy"].to(dev)).sum() + (lad * gold[name + "/gl"].to(dev)).sum(), args)
    for got, key in zip(grads, ("gx", "gw", "gh", "gdl", "gdr")):
        s = max(1e-2, gold[name + "/" + key + "64"].abs().mean().item())
        assert_parity(got, gold[name + "/" + key + "32"], gold[name + "/" + key + "64"], 1e-4, s, name + " " + key)


def test_cubic_coupling_layer_round_trip_and_oracle(dev):
    """PiecewiseCubicCouplingTransform with linear tails and an unconditional PiecewiseCubicCDF: inverse(forward(x)) = x,
    log-dets cancel; inside the box the transformed columns match the fp64 oracle on the conditioner's parameters."""
    torch.manual_seed(3)
    mask = torch.tensor([1, 0, 1, 0, 1, 0, 1, 0])
    layer = transforms.PiecewiseCubicCouplingTransform(
        mask, lambda i, o: nets.ResidualNet(i, o, hidden_features=32, num_blocks=1), num_bins=8, tails="linear",
        tail_bound=3.0, apply_unconditional_transform=True).to(dev)
    for p in layer.parameters():
        p.data.add_(0.2 * torch.randn_like(p))
    x = torch.randn(5000, 8, generator=torch.Generator().manual_seed(2)).to(dev) * 1.5
    with torch.no_grad():
        y, lad = layer(x)
        xi, ladi = layer.inverse(y)
    assert torch.isfinite(y).all() and torch.isfinite(lad).all()
    assert (xi - x).abs().max() < 5e-4 and (lad + ladi).abs().max() < 5e-3
    # a 64-feature CDF layer (TMA-ring kernel) on the unit box against the fp64 oracle
    cdf = transforms.PiecewiseCubicCDF([64], num_bins=8).to(dev)
    u = torch.rand(4096, 64, generator=torch.Generator().manual_seed(4)).to(dev) * 0.998 + 0.001
    with torch.no_grad():
        v, l = cdf(u)
        ui, li = cdf.inverse(v)
    n = u.numel()

    def flat(t, width):
        return t.detach().cpu().double()[None].expand(4096, -1, -1).reshape(n, width)

    ref_v, ref_l = restated.cubic_spline(u.cpu().double().reshape(n), flat(cdf.unnormalized_widths, 8),
                                         flat(cdf.unnormalized_heights, 8), flat(cdf.unnorm_derivatives_left, 1),
                                         flat(cdf.unnorm_derivatives_right, 1))
    assert (v.cpu().double().reshape(n) - ref_v).abs().max() < 2e-5
    assert (l.cpu().double() - ref_l.reshape(4096, 64).sum(-1)).abs().max() < 1e-3
    assert (ui - u).abs().max() < 2e-4 and (l + li).abs().max() < 5e-3
