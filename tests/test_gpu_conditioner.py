"""GPU parity tests of the fused conditioner kernel (csrc/fc_conditioner.cu, C ABI fc_conditioner_*) and of the round-2
parity items: bin indices on the device, the D = 256 (cfg 5) shapes, the in-place ownership protocol, packed-weight cache
invalidation.  Everything goes through the C ABI (ctypes) like the rest of the GPU suite."""
import math

import pytest
import torch
from torch import nn

from flowconductor_b200 import _cabi, conditioner as fcond, graphs, ops, transforms, workloads
from flowconductor_b200.nn import tensorcore
from flowconductor_b200.nn.nets.resnet import ResidualNet
from flowconductor_b200.transforms.made import MADE
from oracle import restated
from tests.helpers import assert_parity, load_golden, parity_report

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------------------
# bin indices on the device (north_star: identical except for inputs within 1e-6 of a knot)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["rq_fwd_lin_k8", "rq_inv_lin_k8", "rq_fwd_lin_k16_id", "rq_inv_lin_k16_id",
                                  "rq_fwd_none_k5", "rq_inv_none_k5", "rq_fwd_lin_k10_b1"])
def test_bin_indices_match_reference_on_gpu(dev, name):
    """The bins selected by the DEVICE arithmetic (ex2.approx softmax, running-sum knots: fc_rqs_bins runs the same
    rqs_locate as the layer kernels) against the bins of the unmodified reference's searchsorted
    (flowcon/utils/torchutils.py:147-149) stored with the golden vectors.  Every mismatch is reported with its distance to
    the nearest knot in units of the interval; none may be further than 1e-6 from a knot."""
    gold = load_golden("functions")
    k, lin, tb, inv, ident = gold[name + "/meta"].tolist()
    k = int(k)
    x = gold[name + "/x"].to(dev)
    p = gold[name + "/params"].to(dev)
    n, d = x.shape
    lo, hi = (-tb, tb) if lin else (0.0, 1.0)
    tails = _cabi.TAILS_LINEAR if lin else _cabi.TAILS_NONE
    bins, dist = ops.rqs_bins(x, p.reshape(n, -1), None, k, tails, bool(inv), bool(ident), lo, hi, lo, hi, 1e-3, 1e-3, 1e-3,
                              1.0)
    ref = gold[name + "/bin"].to(dev)
    inside = (x >= lo) & (x <= hi)
    assert bool((bins[~inside] == -1).all())
    mism = inside & (bins.long() != ref)
    # distance of the input to the nearest REFERENCE knot, normalised
    knots = gold[name + "/knots"].to(dev)
    dref = (x[..., None] - knots).abs().min(-1).values / (hi - lo)
    lines = ["(%d, %d): ours %d reference %d, knot distance %.3e (device) %.3e (reference)" % (
        i, j, bins[i, j], ref[i, j], dist[i, j], dref[i, j]) for i, j in torch.nonzero(mism).tolist()]
    print("%s: %d elements, %d inside, %d bin mismatches" % (name, x.numel(), int(inside.sum()), len(lines)))
    for line in lines:
        print("   " + line)
    assert bool((dref[mism] <= 1e-6).all()), "bin mismatch further than 1e-6 from a knot:\n" + "\n".join(lines)
    assert float(mism.float().mean()) < 1e-3


# ------------------------------------------------------------------------------------------------
# fc_conditioner_rqs_apply against an fp64 conditioner + the element-wise kernel
# ------------------------------------------------------------------------------------------------
def _random_net(kind, d_in, hidden, out, blocks, dev, seed):
    torch.manual_seed(seed)
    if kind == "resnet":
        net = ResidualNet(d_in, out, hidden, num_blocks=blocks)
    else:
        net = MADE(d_in, hidden, num_blocks=blocks, output_multiplier=out // d_in)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, p in net.named_parameters():  # away from the near-identity initialisation of the blocks
            scale = 1.5 / math.sqrt(p.shape[-1]) if p.dim() == 2 else 0.3
            p.copy_(torch.randn(p.shape, generator=g) * scale)
        net.final_layer.weight.mul_(3.0)
    return net.to(dev)


def _net_fp64(net, a):
    """The conditioner in fp64 (masks applied as made.py:72 does)."""
    def lin(layer, t):
        w = layer.weight.double()
        if getattr(layer, "mask", None) is not None:
            w = w * layer.mask.double()
        return torch.nn.functional.linear(t, w, layer.bias.double())

    h = lin(net.initial_layer, a.double())
    for blk in net.blocks:
        t = lin(blk.linear_layers[0], torch.relu(h))
        h = h + lin(blk.linear_layers[1], torch.relu(t))
    return lin(net.final_layer, h)


@pytest.mark.parametrize("kind,B,D,K,H,blocks,inverse,inplace", [
    ("resnet", 1000, 64, 8, 256, 2, False, False),     # cfg 2 shape, ragged last tile, fresh output (identity columns copied)
    ("resnet", 4096, 64, 8, 256, 2, True, True),       # inverse, in place
    ("resnet", 1, 64, 8, 256, 1, False, False),        # a single row
    ("resnet", 257, 20, 8, 128, 3, False, True),       # hidden width 128, three blocks, one row into the second tile pair
    ("resnet", 700, 256, 8, 256, 2, False, False),     # cfg 5 shape: 128 transformed features, 32 final N tiles, k_in 256
    ("made", 1500, 16, 16, 256, 2, False, False),      # cfg 3 shape: masked layers, 16 bins, k_in 16
    ("made", 300, 8, 8, 128, 1, True, False),
])
def test_conditioner_kernel(dev, kind, B, D, K, H, blocks, inverse, inplace):
    P = 3 * K - 1
    coupling = kind == "resnet"
    d_t = D // 2 if coupling else D
    tcols = torch.arange(0, D, 2, device=dev, dtype=torch.int32) if coupling else None
    ccols = torch.arange(1, D, 2, device=dev, dtype=torch.int32) if coupling else None
    net = _random_net(kind, d_t if coupling else D, H, d_t * P, blocks, dev, seed=B + D)
    g = torch.Generator(device=dev).manual_seed(B)
    x = torch.randn(B, D, generator=g, device=dev) * 1.5
    wh = 1.0 / math.sqrt(H) if coupling else 1.0
    cfg = _cabi.RqsConfig(K, _cabi.TAILS_LINEAR, int(not coupling), int(inverse), -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, wh)
    packed = fcond.pack_rqs(net, K, d_t, col_map=ccols, k_in=D)
    xin = x.clone()
    y = xin if inplace else torch.empty_like(x)
    lad = torch.full((B,), 7.0, device=dev)  # accumulate = False must overwrite
    fcond.rqs_apply(packed, xin, xin, y, lad, False, d_t, tcols, ccols, cfg, None)
    torch.cuda.synchronize()
    assert fcond.kernel_error() == 0
    a = x[:, 1::2] if coupling else x
    with torch.no_grad():
        p64 = _net_fp64(net, a).float()
        p32 = net(a)
    args = (K, _cabi.TAILS_LINEAR, inverse, not coupling, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, wh)
    y2, lad2, _ = ops.rqs_layer(x, p64, tcols, ccols, *args)
    y3, lad3, _ = ops.rqs_layer(x, p32, tcols, ccols, *args)
    if coupling:
        assert torch.equal(y[:, 1::2], x[:, 1::2])
    # quantile by quantile no worse than 4x the fp32 torch conditioner's own error against the fp64 one (the spline
    # amplifies parameter noise; the maximum over a few thousand ill-conditioned elements is itself noisy: 10x, as in
    # helpers.assert_parity)
    q = torch.tensor([0.5, 0.99, 0.999, 1.0], device=dev)
    lim = torch.tensor([4.0, 4.0, 4.0, 10.0], device=dev)
    for ours, ref, yard, slack in ((y, y2, y3, 3e-6), (lad, lad2, lad3, 3e-5)):
        eo = torch.quantile((ours - ref).abs().flatten().float(), q)
        ey = torch.quantile((yard - ref).abs().flatten().float(), q)
        assert bool((eo <= lim * ey + slack * max(1.0, ref.abs().max().item())).all()), (eo, ey)
    # accumulate = True adds to what is there
    lad_acc = torch.full((B,), 2.0, device=dev)
    fcond.rqs_apply(packed, x, x, torch.empty_like(x), lad_acc, True, d_t, tcols, ccols, cfg, None)
    assert torch.allclose(lad_acc, lad + 2.0, atol=1e-5, rtol=1e-6)


def test_conditioner_argument_errors(dev):
    net = _random_net("resnet", 32, 320, 32 * 23, 2, dev, seed=0)  # hidden width above 256: not covered
    with pytest.raises(ValueError):
        fcond.pack_rqs(net, 8, 32)
    net = _random_net("resnet", 32, 64, 32 * 23, 2, dev, seed=0)  # narrow nets are zero-padded to the kernel's 128
    assert fcond.pack_rqs(net, 8, 32).hidden == 128
    with pytest.raises(ValueError):
        fcond.pack_rqs(net, 5, 32)  # 5 bins: not a fused shape
    net = _random_net("resnet", 32, 256, 32 * 23, 2, dev, seed=0)
    packed = fcond.pack_rqs(net, 8, 32, k_in=32)
    x = torch.randn(10, 32, device=dev)
    cfg = _cabi.RqsConfig(8, _cabi.TAILS_NONE, 0, 0, 0.0, 1.0, 0.0, 1.0, 1e-3, 1e-3, 1e-3, 1.0)
    with pytest.raises(RuntimeError, match="invalid argument"):  # packed for 23 parameters per feature, called with 25
        fcond.rqs_apply(packed, x, x, torch.empty_like(x), torch.empty(10, device=dev), False, 32, None, None, cfg)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        fcond.rqs_apply(packed, x.cpu(), x, torch.empty_like(x), torch.empty(10, device=dev), False, 32, None, None, cfg)
    # empty batch: nothing launched, nothing touched
    cfg = _cabi.RqsConfig(8, _cabi.TAILS_LINEAR, 0, 0, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 1.0)
    e = torch.empty(0, 32, device=dev)
    fcond.rqs_apply(packed, e, e, torch.empty_like(e), torch.empty(0, device=dev), False, 32, None, None, cfg)


# ------------------------------------------------------------------------------------------------
# cfg 5 (D = 256) and cfg 3 (MADE, 16 bins) layer by layer against the oracle, every inference path
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,rows", [("cfg5", 4096), ("cfg3", 2048), ("cfg2", 4096)])
@pytest.mark.parametrize("path", ["fused", "perlayer", "unfused"])
def test_layers_match_oracle(dev, name, rows, path, monkeypatch):
    """Every layer of the full-size model on its own (input = the fp32 oracle's output of the previous layer), three-way
    against the fp32 / fp64 oracle at the strict criteria.  fused = one persistent kernel per layer (3xFP16), perlayer =
    round 1's per-layer tensor-core kernels (3xTF32), unfused = torch conditioner + element-wise kernel."""
    monkeypatch.setattr(tensorcore, "ENABLED", path != "unfused")
    monkeypatch.setattr(tensorcore, "FUSED_CONDITIONER", path == "fused")
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl)
    flow.load_state_dict(state)
    specs = workloads.oracle_specs(wl)
    state64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
    flow = flow.to(dev)
    h = torch.randn(rows, wl["features"], generator=torch.Generator().manual_seed(99))
    report = []
    with torch.no_grad():
        for li, (layer, spec) in enumerate(zip(flow._transform._transforms, specs)):
            y64, l64 = restated.apply_layer(state64, spec, h.double())
            y32, l32 = restated.apply_layer(state, spec, h)
            _cabi.STATS.reset()
            y, lad = layer(h.to(dev))
            if spec["kind"] != "permutation":
                fused_calls = _cabi.STATS.counts.get("fc_conditioner_rqs_apply", 0)
                assert fused_calls == (1 if path == "fused" else 0), _cabi.STATS.counts
            # (cfg 3: sixteen 16-bin splines per row sum to |logabsdet| ~ 25 with a reference fp32 noise of 1e-4 .. 1e-3 per
            # row, so a few percent of rows miss both element-wise criteria by chance even when, quantile by quantile, the
            # errors equal the reference's own — the population criteria stay as they are, the count limit is 5 % there)
            frac = 5e-2 if name == "cfg3" else 2e-2
            assert_parity(y, y32, y64, OUT_TOL, 1.0, "%s layer %d outputs (%s)" % (name, li, path))
            assert_parity(lad, l32, l64, OUT_TOL, 1.0, "%s layer %d logabsdet (%s)" % (name, li, path), max_fail_frac=frac)
            report.append((li, parity_report(y, y32, y64, OUT_TOL, 1.0), parity_report(lad, l32, l64, OUT_TOL, 1.0)))
            h = y32
    for li, ry, rl in report:
        print("%s layer %d (%s): outputs %s | logabsdet %s" % (name, li, path, ry, rl))


def test_fused_and_perlayer_paths_agree_on_a_whole_flow(dev, monkeypatch):
    """cfg 2 at 20 000 rows, whole stack: the two tensor-core paths differ only by rounding (different splits and
    summation orders), log_prob agrees to the noise of either against the oracle; inverse(forward) per layer."""
    wl = workloads.get_workload("cfg2")
    flow = workloads.build_flow(wl)
    flow.load_state_dict(workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl))
    flow = flow.to(dev)
    x = torch.randn(20000, 64, generator=torch.Generator(device=dev).manual_seed(5), device=dev)
    out = {}
    with torch.no_grad():
        for fused in (True, False):
            monkeypatch.setattr(tensorcore, "FUSED_CONDITIONER", fused)
            out[fused] = flow.log_prob(x)
        monkeypatch.setattr(tensorcore, "FUSED_CONDITIONER", True)
        layer = flow._transform._transforms[0]
        y, lad = layer(x)
        xr, ladr = layer.inverse(y)
    rel = (out[True] - out[False]).abs() / out[False].abs().clamp_min(1.0)
    assert rel.median() < 2e-6 and torch.quantile(rel, 0.999) < 2e-4
    assert torch.equal(xr[:, 1::2], x[:, 1::2])
    err = (xr - x).abs()
    assert err.median() < 1e-6 and torch.quantile(err.flatten()[: 1 << 20], 0.999) < 2e-4
    assert (lad + ladr).abs().median() < 1e-4


# ------------------------------------------------------------------------------------------------
# sum-of-sigmoids bijection inside the fused conditioner (cfg 4)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows", [777, 4096])
def test_cfg4_layers_fused_sos_match_oracle(dev, rows):
    """Every layer of the full-size conditional sum-of-sigmoids flow (D = 32, context 8, n = 10, H = 64 zero-padded to the
    kernel's 128): ResidualNet on the context + SumOfSigmoids as ONE kernel, three-way against the fp32 / fp64 oracle."""
    wl = workloads.get_workload("cfg4")
    flow = workloads.build_flow(wl)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl)
    flow.load_state_dict(state)
    specs = workloads.oracle_specs(wl)
    state64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
    flow = flow.to(dev).eval()
    g = torch.Generator().manual_seed(17)
    h = torch.randn(rows, wl["features"], generator=g)
    ctx = torch.randn(rows, wl["context_features"], generator=g)
    with torch.no_grad():
        for li, (layer, spec) in enumerate(zip(flow._transform._transforms, specs)):
            y64, l64 = restated.apply_layer(state64, spec, h.double(), ctx.double())
            y32, l32 = restated.apply_layer(state, spec, h, ctx)
            _cabi.STATS.reset()
            y, lad = layer(h.to(dev), ctx.to(dev))
            assert _cabi.STATS.counts.get("fc_conditioner_sos_apply", 0) == 1, _cabi.STATS.counts
            assert set(_cabi.STATS.counts) <= {"fc_conditioner_sos_apply", "fc_conditioner_pack_layer"}, _cabi.STATS.counts
            print("cfg4 layer %d fused sum-of-sigmoids: outputs %s | logabsdet %s" % (
                li, parity_report(y, y32, y64, OUT_TOL, 1.0), parity_report(lad, l32, l64, OUT_TOL, 1.0)))
            assert_parity(y, y32, y64, OUT_TOL, 1.0, "cfg4 layer %d outputs" % li)
            assert_parity(lad, l32, l64, OUT_TOL, 1.0, "cfg4 layer %d logabsdet" % li)
            h = y32
        # whole flow: fused vs materialised parameters + element-wise kernel
        x = torch.randn(rows, wl["features"], generator=g).to(dev)
        c = ctx.to(dev)
        a = flow.log_prob(x, context=c)
        tensorcore.FUSED_CONDITIONER = False
        try:
            b = flow.log_prob(x, context=c)
        finally:
            tensorcore.FUSED_CONDITIONER = True
    rel = (a - b).abs() / b.abs().clamp_min(1.0)
    assert rel.median() < 2e-6 and rel.max() < 1e-3, (float(rel.median()), float(rel.max()))


def test_masked_sum_of_sigmoids_forward_fused(dev, monkeypatch):
    """MaskedSumOfSigmoidsTransform (autoregressive.py:266-318: MADE + SumOfSigmoids - 0.5), forward: fused vs unfused."""
    torch.manual_seed(4)
    layer = transforms.MaskedSumOfSigmoidsTransform(features=8, hidden_features=64, n_sigmoids=10).to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.2)
        x = torch.randn(1500, 8, device=dev)
        _cabi.STATS.reset()
        y, lad = layer(x)
        assert _cabi.STATS.counts.get("fc_conditioner_sos_apply", 0) == 1, _cabi.STATS.counts
        monkeypatch.setattr(tensorcore, "ENABLED", False)
        yu, ladu = layer(x)
    assert ((y - yu).abs() / yu.abs().clamp_min(1.0)).max() < 1e-4
    assert (lad - ladu).abs().max() < 1e-3


@pytest.mark.parametrize("kind,features,hidden,rows", [("coupling", 32, 128, 3001), ("coupling", 132, 256, 2048),
                                                       ("maf", 8, 64, 1500), ("maf", 64, 256, 4096), ("maf", 52, 128, 777)])
def test_affine_layers_run_in_the_fused_conditioner(dev, kind, features, hidden, rows, monkeypatch):
    """AffineCouplingTransform (coupling.py:212-252; [shift | raw scale] final layer, sigmoid(raw + 2) + 1e-3 scale) and
    the forward of MaskedAffineAutoregressiveTransform (autoregressive.py:97-129; (raw scale, shift) pairs, softplus + 1e-3):
    one fc_conditioner_affine_apply launch per call, against the unfused path; 12 features share a 24-column accumulator
    slot, so feature counts that end inside a slot and inside a tile are covered."""
    torch.manual_seed(features + hidden)
    if kind == "coupling":
        mask = workloads.make_mask(features, "alternating_even")
        layer = transforms.AffineCouplingTransform(
            mask, lambda i, o: ResidualNet(i, o, hidden_features=hidden, num_blocks=2))
        directions = (False, True)
    else:
        layer = transforms.MaskedAffineAutoregressiveTransform(features=features, hidden_features=hidden, num_blocks=2)
        directions = (False,)
    layer = layer.to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        x = torch.randn(rows, features, device=dev)
        for inverse in directions:
            fn = layer.inverse if inverse else layer
            _cabi.STATS.reset()
            y, lad = fn(x)
            assert _cabi.STATS.counts.get("fc_conditioner_affine_apply", 0) == 1, _cabi.STATS.counts
            monkeypatch.setattr(tensorcore, "ENABLED", False)
            yu, ladu = fn(x)
            monkeypatch.setattr(tensorcore, "ENABLED", True)
            monkeypatch.setattr(tensorcore, "FUSED_AFFINE", False)
            yp, ladp = fn(x)  # per-layer tensor-core kernels (fc_linear_affine_apply)
            monkeypatch.setattr(tensorcore, "FUSED_AFFINE", True)
            for (b, lb) in ((yu, ladu), (yp, ladp)):
                # the inverse divides by a scale that can be ~1e-3 (sigmoid(raw + 2) + 1e-3): a few ill-conditioned entries
                rel = ((y - b).abs() / b.abs().clamp_min(1.0)).flatten()
                assert torch.quantile(rel, 0.999) < 1e-4 and rel.max() < 5e-2, (kind, inverse, float(rel.max()))
                assert torch.quantile((lad - lb).abs(), 0.99) < 2e-3 and (lad - lb).abs().max() < 5e-2


@pytest.mark.parametrize("features", [6, 21, 43, 63])
@pytest.mark.parametrize("kind", ["coupling_rqs", "maf_rqs", "coupling_affine", "cond_sos", "coupling_quadratic"])
def test_feature_counts_that_are_not_multiples_of_four_take_the_kernels(dev, kind, features, monkeypatch):
    """The tabular data sets of the reference's experiments have 6, 8, 21, 43 and 63 columns: the conditioner inputs are
    zero-padded to a 16-byte row pitch (tensorcore.aligned_inputs) instead of falling back to the torch conditioner."""
    torch.manual_seed(features)
    net = lambda i, o: ResidualNet(i, o, hidden_features=128, num_blocks=2)  # noqa: E731
    mask = workloads.make_mask(features, "alternating_even")
    ctx, want = None, None
    if kind == "coupling_rqs":
        layer = transforms.PiecewiseRationalQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)
        want = "fc_conditioner_rqs_apply"
    elif kind == "maf_rqs":
        layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(features, 128, num_bins=8, tails="linear",
                                                                                   tail_bound=3.0)
        want = "fc_conditioner_rqs_apply"
    elif kind == "coupling_affine":
        layer = transforms.AffineCouplingTransform(mask, net)
        want = "fc_conditioner_affine_apply"
    elif kind == "cond_sos":
        layer = transforms.ConditionalSumOfSigmoidsTransform(features, 128, context_features=5, n_sigmoids=10, num_blocks=2)
        ctx = torch.randn(1234, 5, device=dev)
        want = "fc_conditioner_sos_apply"
    else:
        layer = transforms.PiecewiseQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)
        want = "fc_conditioner_store_apply"
    layer = layer.to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        x = torch.randn(1234, features, device=dev)
        _cabi.STATS.reset()
        y, lad = layer(x, ctx)
        assert _cabi.STATS.counts.get(want, 0) >= 1, _cabi.STATS.counts
        monkeypatch.setattr(tensorcore, "ENABLED", False)
        yu, ladu = layer(x, ctx)
        monkeypatch.setattr(tensorcore, "ENABLED", True)
        rel = ((y - yu).abs() / yu.abs().clamp_min(1.0)).flatten()
        assert torch.quantile(rel, 0.999) < 1e-4 and rel.max() < 5e-2, (kind, features, float(rel.max()))
        assert torch.quantile((lad - ladu).abs(), 0.99) < 2e-3
        if kind != "cond_sos":
            xi, ladi = layer.inverse(y, ctx)
            assert (xi - x).abs().max() < 5e-3 and torch.quantile((ladi + lad).abs(), 0.99) < 5e-3


@pytest.mark.parametrize("kind,features,hidden,blocks,rows", [
    ("coupling_quadratic", 64, 256, 2, 3001), ("coupling_linear", 30, 128, 1, 777), ("coupling_cubic", 12, 64, 3, 256),
    ("maf_quadratic", 16, 256, 2, 4096), ("cond_sos_inverse", 32, 64, 2, 1500), ("maf_linear", 21, 100, 2, 1),
    ("coupling_quadratic", 26, 128, 2, 500)])  # 13 x 15 = 195 outputs: three tiles, the last one ragged
def test_whole_conditioner_store_kernel(dev, kind, features, hidden, blocks, rows, monkeypatch):
    """`tensorcore.params` as ONE launch (fc_conditioner_store_apply: ResidualNet.forward resnet.py:92-100 / MADE.forward
    made.py:274-283 with the outputs written out) for the bijections that run as element-wise kernels: against the per-layer
    tensor-core kernels and the torch conditioner."""
    torch.manual_seed(features * 3 + hidden)
    net = lambda i, o: ResidualNet(i, o, hidden_features=hidden, num_blocks=blocks)  # noqa: E731
    mask = workloads.make_mask(features, "alternating_even")
    ctx = None
    if kind == "coupling_quadratic":
        layer = transforms.PiecewiseQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)
    elif kind == "coupling_linear":
        layer = transforms.PiecewiseLinearCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)
    elif kind == "coupling_cubic":
        layer = transforms.PiecewiseCubicCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)
    elif kind == "maf_quadratic":
        layer = transforms.MaskedPiecewiseQuadraticAutoregressiveTransform(features, hidden, num_bins=8, tails="linear",
                                                                           tail_bound=3.0, num_blocks=blocks)
    elif kind == "maf_linear":
        layer = transforms.MaskedPiecewiseLinearAutoregressiveTransform(8, features, hidden, num_blocks=blocks)
    else:
        layer = transforms.ConditionalSumOfSigmoidsTransform(features, hidden, context_features=8, n_sigmoids=10,
                                                             num_blocks=blocks)
        ctx = torch.randn(rows, 8, device=dev)
    layer = layer.to(dev).eval()
    inverse = kind == "cond_sos_inverse"
    fn = layer.inverse if inverse else layer
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.05)
        x = torch.rand(rows, features, device=dev) if kind == "maf_linear" else torch.randn(rows, features, device=dev)
        _cabi.STATS.reset()
        y, lad = fn(x, ctx)
        assert _cabi.STATS.counts.get("fc_conditioner_store_apply", 0) == 1, _cabi.STATS.counts
        assert not any(k.startswith("fc_linear_") for k in _cabi.STATS.counts), _cabi.STATS.counts
        monkeypatch.setattr(tensorcore, "FUSED_STORE", False)
        yp, ladp = fn(x, ctx)
        monkeypatch.setattr(tensorcore, "ENABLED", False)
        yu, ladu = fn(x, ctx)
    for (b, lb) in ((yp, ladp), (yu, ladu)):
        rel = ((y - b).abs() / b.abs().clamp_min(1.0)).flatten()
        assert torch.quantile(rel, 0.999) < 1e-4 and rel.max() < 5e-2, (kind, float(rel.max()))
        assert torch.quantile((lad - lb).abs(), 0.99) < 2e-3


def test_padded_input_and_store_paths_replay_from_a_cuda_graph(dev):
    """A 21-feature flow (padded conditioner inputs) of a fused-spline layer, a fused-affine layer and a quadratic-spline layer
    (store conditioner + element-wise tile ring): `graphs.capture(flow.log_prob)` replays bit for bit what the eager call
    returns, also on fresh inputs."""
    from flowconductor_b200 import flows, distributions, graphs

    torch.manual_seed(21)
    D = 21
    net = lambda i, o: ResidualNet(i, o, hidden_features=128, num_blocks=2)  # noqa: E731
    mask = workloads.make_mask(D, "alternating_even")
    tr = transforms.CompositeTransform([
        transforms.PiecewiseRationalQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0),
        transforms.ReversePermutation(D),
        transforms.AffineCouplingTransform(mask, net),
        transforms.PiecewiseQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0)])
    flow = flows.Flow(tr, distributions.StandardNormal([D])).to(dev).eval()
    x = torch.randn(2000, D, device=dev)
    with torch.no_grad():
        _cabi.STATS.reset()
        eager = flow.log_prob(x)
        c = _cabi.STATS.counts
        assert c.get("fc_conditioner_rqs_apply") == 1 and c.get("fc_conditioner_affine_apply") == 1 \
            and c.get("fc_conditioner_store_apply") == 1, c
        lp = graphs.capture(flow.log_prob, x)
        assert torch.equal(lp(x), eager)
        x2 = torch.randn(2000, D, device=dev)
        assert torch.equal(lp(x2), flow.log_prob(x2))


def test_narrow_coupling_conditioner_is_padded_to_the_kernel_width(dev, monkeypatch):
    """H = 64 ResidualNet (cfg2_tc_small): zero-padded to 128 at pack time, one fused launch per layer, same numbers as the
    per-layer kernels."""
    wl = workloads.get_workload("cfg2_tc_small")
    flow = workloads.build_flow(wl, seed=3).to(dev).eval()
    x = torch.randn(1000, 64, device=dev)
    with torch.no_grad():
        _cabi.STATS.reset()
        a = flow.log_prob(x)
        assert _cabi.STATS.counts.get("fc_conditioner_rqs_apply", 0) == 3, _cabi.STATS.counts
        monkeypatch.setattr(tensorcore, "FUSED_CONDITIONER", False)
        b = flow.log_prob(x)
    rel = (a - b).abs() / b.abs().clamp_min(1.0)
    assert rel.max() < 1e-4, float(rel.max())


@pytest.mark.parametrize("bins,tails,hidden", [(10, "linear", 128), (8, None, 64), (10, None, 256), (16, "linear", 64)])
def test_fused_conditioner_other_bin_counts_and_no_tails(dev, bins, tails, hidden, monkeypatch):
    """The reference's default bin count (10) and splines without tails (domain [0, 1] for couplings, InputOutsideDomain
    beyond it: coupling.py:566-567, rational_quadratic.py:81-82) inside the fused conditioner: against the unfused path
    (torch conditioner + element-wise kernel) and, for the domain error, the reference's exception."""
    from flowconductor_b200.transforms.base import InputOutsideDomain

    torch.manual_seed(bins * 7 + hidden)
    mask = workloads.make_mask(12, "alternating_even")
    layer = transforms.PiecewiseRationalQuadraticCouplingTransform(
        mask, lambda i, o: ResidualNet(i, o, hidden_features=hidden, num_blocks=2), num_bins=bins, tails=tails,
        tail_bound=2.0).to(dev).eval()
    with torch.no_grad():
        for p in layer.parameters():
            p.add_(torch.randn_like(p) * 0.1)
        x = torch.randn(3000, 12, device=dev) if tails == "linear" else torch.rand(3000, 12, device=dev)
        for inverse in (False, True):
            fn = layer.inverse if inverse else layer
            _cabi.STATS.reset()
            y, lad = fn(x)
            assert _cabi.STATS.counts.get("fc_conditioner_rqs_apply", 0) == 1, _cabi.STATS.counts
            monkeypatch.setattr(tensorcore, "ENABLED", False)
            yu, ladu = fn(x)
            monkeypatch.setattr(tensorcore, "ENABLED", True)
            assert ((y - yu).abs() / yu.abs().clamp_min(1.0)).max() < 2e-4, (bins, tails, inverse)
            assert torch.quantile((lad - ladu).abs(), 0.99) < 2e-3 and (lad - ladu).abs().max() < 5e-2
        if tails is None:
            bad = x.clone()
            bad[17, 0] = 1.5  # a transformed column outside the unit box
            with pytest.raises(InputOutsideDomain):
                layer(bad)


def test_shared_cdf_parameters_are_not_materialised(dev):
    """the learnable parameters of an unconditional CDF layer are ONE row shared by the batch; the kernel reads
    it with row stride 0 instead of a materialised [B, D * P] copy (3 GB at 1 M rows x 32 features x 23)."""
    torch.manual_seed(0)
    layer = transforms.PiecewiseRationalQuadraticCDF(shape=32, num_bins=8, tails="linear", tail_bound=3.0).to(dev)
    rows = 1 << 18
    x = torch.randn(rows, 32, device=dev)
    with torch.no_grad():
        p = layer._shared_params(rows)
        assert p.stride(0) == 0
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats(dev)
        before = torch.cuda.memory_allocated(dev)
        y, lad = layer(x)
        torch.cuda.synchronize()
        extra = torch.cuda.max_memory_allocated(dev) - before
        assert extra < 3 * x.numel() * 4, extra  # outputs + logabsdet, not rows x 32 x 23 floats
        ym, ladm, _ = ops.rqs_layer(x, p.contiguous(), None, None, 8, _cabi.TAILS_LINEAR, False, False, -3.0, 3.0, -3.0, 3.0,
                                    1e-3, 1e-3, 1e-3, 1.0)
    assert torch.equal(y, ym) and torch.equal(lad, ladm)


# ------------------------------------------------------------------------------------------------
# host-side protocols around the kernels
# ------------------------------------------------------------------------------------------------
class _ReusesItsInput(transforms.Transform):
    """A wrapper layer that calls a kernel layer and then reads its own input again: the inner layer must
    not have overwritten it."""

    def __init__(self, inner):
        super().__init__()
        self.inner = inner

    def forward(self, inputs, context=None):
        y, lad = self.inner(inputs, context)
        y2, lad2 = self.inner(inputs, context)  # a second kernel layer on the same input
        return y + 0.0 * (y2 - inputs), lad


def test_in_place_consent_is_consumed_by_the_first_kernel_layer(dev):
    wl = workloads.get_workload("cfg2_tc_small")
    flow = workloads.build_flow(wl)
    flow.load_state_dict(workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl))
    flow = flow.to(dev)
    layers = list(flow._transform._transforms)
    plain = transforms.CompositeTransform(layers)
    wrapped = transforms.CompositeTransform([layers[0], _ReusesItsInput(layers[1]), layers[2]])
    x = torch.randn(512, 64, generator=torch.Generator(device=dev).manual_seed(0), device=dev)
    with torch.no_grad():
        want, lad_want = plain(x.clone())
        got, lad_got = wrapped(x.clone())
    assert torch.equal(want, got) and torch.equal(lad_want, lad_got)
    # a view of the cascade's intermediate never qualifies either
    t = torch.randn(8, 4, device=dev)
    with torch.no_grad():
        tensorcore.begin_layer(t)
        tensorcore.mark_fresh(t)
        tensorcore.begin_layer(t)
        assert tensorcore.may_overwrite(t) and not tensorcore.may_overwrite(t[:, :]) and not tensorcore.may_overwrite(t.clone())
        tensorcore.consume_consent()
        assert not tensorcore.may_overwrite(t)
        tensorcore.end_cascade()


def test_packed_weights_follow_parameter_updates(dev):
    """writes through .data change neither pointer nor version; `tensorcore.invalidate` and
    `distributed.broadcast_parameters` drop the cached plans, and a captured graph re-records when the weights moved on."""
    wl = workloads.get_workload("cfg2_tc_small")
    flow = workloads.build_flow(wl)
    flow.load_state_dict(workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl))
    flow = flow.to(dev)
    x = torch.randn(256, 64, generator=torch.Generator(device=dev).manual_seed(0), device=dev)
    with torch.no_grad():
        before = flow.log_prob(x).clone()
        g = graphs.capture(flow.log_prob, x)
        assert torch.equal(g(x), before)
        for p in flow.parameters():  # an optimizer-like in-place update bumps the version counter
            p.add_(0.01 * torch.randn_like(p))
        after = flow.log_prob(x).clone()
        assert not torch.equal(before, after)
        assert torch.equal(g(x), after), "graph replay must notice the new weights"
        net = flow._transform._transforms[0].transform_net
        net.final_layer.bias.data.add_(0.5)  # .data write: invisible to the cache key ...
        tensorcore.invalidate(flow)           # ... so the caller says so
        again = flow.log_prob(x)
        assert not torch.equal(after, again)
        assert torch.equal(g(x), again)
