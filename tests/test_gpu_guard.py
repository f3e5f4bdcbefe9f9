"""Out-of-bounds check of the kernels without compute-sanitizer (closed on this pool): every CUDA buffer the package allocates
while a layer runs is placed inside a larger sentinel-filled buffer, and after the call the guard bands must be untouched.
Ragged batch sizes on purpose (1 row, odd numbers of 128- and 256-row tiles)."""
import math

import pytest
import torch

from flowconductor_b200 import made_inverse, transforms, workloads
from flowconductor_b200.nn import tensorcore
from flowconductor_b200.nn.nets import ResidualNet

pytestmark = pytest.mark.gpu
SENT_F, SENT_I, GUARD = 98765.4321, 0x5A5A5A5A, 1 << 14


class GuardedAllocations:
    """Wraps torch.empty / empty_like / zeros (what the package allocates its outputs and scratch with)."""

    def __enter__(self):
        self.allocs = []
        self._orig = (torch.empty, torch.empty_like, torch.zeros)
        orig_empty = torch.empty

        def guarded(shape, dtype, device, zero):
            n = int(math.prod(shape))
            big = orig_empty((n + 2 * GUARD,), dtype=dtype, device=device)
            big.fill_(SENT_F if dtype.is_floating_point else SENT_I)
            win = big[GUARD:GUARD + n]
            if zero:
                win.zero_()
            self.allocs.append((big, n))
            return win.view(shape)

        def ok(dtype, device):
            return (dtype in (torch.float32, torch.int32) and device is not None and torch.device(device).type == "cuda")

        def empty(*size, dtype=None, device=None, **kw):
            shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
            dtype = dtype or torch.get_default_dtype()
            if ok(dtype, device) and not kw.get("pin_memory"):
                return guarded(shape, dtype, device, False)
            return self._orig[0](*size, dtype=dtype, device=device, **kw)

        def empty_like(t, **kw):
            if t.is_cuda and t.dtype in (torch.float32, torch.int32) and not kw:
                return guarded(tuple(t.shape), t.dtype, t.device, False)
            return self._orig[1](t, **kw)

        def zeros(*size, dtype=None, device=None, **kw):
            shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
            dtype = dtype or torch.get_default_dtype()
            if ok(dtype, device) and not kw:
                return guarded(shape, dtype, device, True)
            return self._orig[2](*size, dtype=dtype, device=device, **kw)

        torch.empty, torch.empty_like, torch.zeros = empty, empty_like, zeros
        return self

    def __exit__(self, *exc):
        torch.empty, torch.empty_like, torch.zeros = self._orig
        return False

    def check(self, what):
        torch.cuda.synchronize()
        assert self.allocs, what + ": nothing was allocated through the guarded functions"
        for big, n in self.allocs:
            sent = SENT_F if big.dtype.is_floating_point else SENT_I
            ref = torch.full((1,), sent, dtype=big.dtype, device=big.device)
            lo, hi = big[:GUARD] != ref, big[GUARD + n:] != ref
            assert not bool(lo.any()) and not bool(hi.any()), "%s: out-of-bounds write around a buffer of %d elements (%d below, %d above)" % (
                what, n, int(lo.sum()), int(hi.sum()))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


ROWS = [1, 129, 300, 641]


@pytest.mark.parametrize("rows", ROWS)
@pytest.mark.parametrize("fused", [True, False])
def test_coupling_flow_paths(dev, rows, fused, monkeypatch):
    monkeypatch.setattr(tensorcore, "FUSED_CONDITIONER", fused)
    flow = workloads.build_flow(workloads.get_workload("cfg2_tc_small"), seed=1).to(dev).eval()
    x = torch.randn(rows, 64, device=dev)
    with torch.no_grad(), GuardedAllocations() as g:
        lp = flow.log_prob(x)
        xi, _ = flow._transform.inverse(x)
    g.check("cfg2_tc_small fused=%s rows=%d" % (fused, rows))
    assert torch.isfinite(lp).all() and torch.isfinite(xi).all()


@pytest.mark.parametrize("rows", ROWS)
def test_other_coupling_families_per_layer_path(dev, rows):
    """Linear / quadratic / cubic / affine couplings with a 128-wide ResidualNet: per-layer tensor-core kernels (T128
    activations) + element-wise kernel."""
    mask = workloads.make_mask(16, "alternating_even")
    create = lambda i, o: ResidualNet(i, o, hidden_features=128, num_blocks=2)  # noqa: E731
    layers = [transforms.PiecewiseLinearCouplingTransform(mask, create, num_bins=8, tails="linear", tail_bound=3.0),
              transforms.PiecewiseQuadraticCouplingTransform(mask, create, num_bins=8, tails="linear", tail_bound=3.0),
              transforms.PiecewiseCubicCouplingTransform(mask, create, num_bins=8, tails="linear", tail_bound=3.0),
              transforms.AffineCouplingTransform(mask, create)]
    x = torch.randn(rows, 16, device=dev)
    for layer in layers:
        layer = layer.to(dev).eval()
        with torch.no_grad(), GuardedAllocations() as g:
            y, lad = layer(x)
            xi, _ = layer.inverse(y)
        g.check("%s rows=%d" % (type(layer).__name__, rows))
        assert (xi - x).abs().max() < 1e-2


@pytest.mark.parametrize("rows", ROWS)
def test_autoregressive_layers(dev, rows, monkeypatch):
    """MAF layers: fused conditioner forward, incremental inverse, D-pass inverse (per-layer kernels with T128 activations:
    the path whose odd-tile-count bug this file exists for)."""
    torch.manual_seed(rows)
    layers = [transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(16, 256, num_bins=16, tails="linear",
                                                                                 tail_bound=3.0),
              transforms.MaskedPiecewiseLinearAutoregressiveTransform(8, 12, 128),
              transforms.MaskedAffineAutoregressiveTransform(10, 64),
              transforms.MaskedSumOfSigmoidsTransform(8, 64, n_sigmoids=10)]
    for layer in layers:
        layer = layer.to(dev).eval()
        d = layer.autoregressive_net.initial_layer.weight.shape[1]
        unit = isinstance(layer, transforms.MaskedPiecewiseLinearAutoregressiveTransform)
        x = torch.rand(rows, d, device=dev) * 0.9 + 0.05 if unit else torch.randn(rows, d, device=dev)
        with torch.no_grad(), GuardedAllocations() as g:
            y, lad = layer(x)
            xi, _ = layer.inverse(y)
            monkeypatch.setattr(made_inverse, "ENABLED", False)
            xd, _ = layer.inverse(y)
            monkeypatch.setattr(made_inverse, "ENABLED", True)
        g.check("%s rows=%d" % (type(layer).__name__, rows))
        assert torch.quantile((xi - xd).abs().flatten(), 0.99) < 1e-3


@pytest.mark.parametrize("rows", [129, 641])
def test_conditional_sum_of_sigmoids_and_actnorm_and_training(dev, rows):
    flow = workloads.build_flow(workloads.get_workload("cfg4"), seed=2).to(dev).eval()
    x, c = torch.randn(rows, 32, device=dev), torch.randn(rows, 8, device=dev)
    with torch.no_grad(), GuardedAllocations() as g:
        lp = flow.log_prob(x, context=c)
    g.check("cfg4 rows=%d" % rows)
    assert torch.isfinite(lp).all()
    # a training step of a MAF with an ActNorm layer in front: forward + backward kernels (tensor-core autograd path)
    model = transforms.CompositeTransform([
        transforms.ActNorm(16),
        transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(16, 128, num_bins=8, tails="linear",
                                                                           tail_bound=3.0)]).to(dev)
    xt = torch.randn(rows, 16, device=dev)
    with GuardedAllocations() as g:
        y, lad = model(xt)
        (-(lad.mean()) + (y ** 2).mean()).backward()
    g.check("training step rows=%d" % rows)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("rows", [1, 37, 1500, 4099])
@pytest.mark.parametrize("D,coupling", [(21, False), (6, True), (64, True)])
def test_element_wise_tile_ring_stays_inside_its_buffers(dev, rows, D, coupling):
    """The tile ring's bulk stores (whole tiles of rows that need not be 16-byte multiples) and the plain stores of a ragged
    last tile: forward, inverse and backward of the spline layer, outputs allocated inside guard bands."""
    from flowconductor_b200 import _cabi, ops

    g0 = torch.Generator(device=dev).manual_seed(rows + D)
    tc = torch.arange(0, D, 2, dtype=torch.int32, device=dev) if coupling else None
    cc = torch.arange(1, D, 2, dtype=torch.int32, device=dev) if coupling else None
    d_t = tc.numel() if coupling else D
    x = torch.randn(rows, D, generator=g0, device=dev)
    p = torch.randn(rows, d_t * 23, generator=g0, device=dev)
    gy, gl = torch.randn(rows, D, generator=g0, device=dev), torch.randn(rows, generator=g0, device=dev)
    rest = (8, _cabi.TAILS_LINEAR, False, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 0.25)
    with torch.no_grad(), GuardedAllocations() as g:
        y, lad, _ = ops.rqs_layer(x, p, tc, cc, *rest)
        assert _cabi.lib().fc_elementwise_last_path() == 2
        xi, _, _ = ops.rqs_layer(y, p, tc, cc, 8, _cabi.TAILS_LINEAR, True, *rest[3:])
        gx, gp = ops.rqs_layer_backward(x, p, gy, gl, tc, cc, *rest)
        assert _cabi.lib().fc_elementwise_last_path() == 2
    g.check("element-wise tile ring rows=%d D=%d" % (rows, D))
    assert (xi - x).abs().max() < 1e-3 and torch.isfinite(gx).all() and torch.isfinite(gp).all()
