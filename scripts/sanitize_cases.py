#!/usr/bin/env python
"""Small shapes of the hand-synchronised kernels for `compute-sanitizer`: the per-warp TMA rings
(pipelined_apply_kernel / pipelined_backward_kernel), the tcgen05 GEMM (linear_tf32x3_kernel), the fused conditioner
(conditioner_f16x3_kernel) and the incremental autoregressive inverse (made_inverse_kernel).

    compute-sanitizer --tool memcheck  python scripts/sanitize_cases.py      # one tool per gpurun call
    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py

Every case also checks its result against the D-pass / unfused path, so a sanitizer-clean run is also a correct one.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import made_inverse, transforms, workloads  # noqa: E402
from flowconductor_b200.nn import nets, tensorcore  # noqa: E402

dev = torch.device("cuda:0")
only = set(sys.argv[1:])


def case(name):
    def deco(fn):
        if not only or name in only:
            fn()
            torch.cuda.synchronize()
            print("ok", name, flush=True)
        return fn
    return deco


@case("rqs_ring")
def _rqs_ring():
    # forward + backward rings (fc_rqs_apply / fc_rqs_backward) through a coupling layer with materialised parameters
    torch.manual_seed(0)
    mask = workloads.make_mask(64, "alternating_even")
    layer = transforms.PiecewiseRationalQuadraticCouplingTransform(
        mask, lambda i, o: nets.ResidualNet(i, o, hidden_features=32, num_blocks=1), num_bins=8, tails="linear",
        tail_bound=3.0).to(dev)
    x = torch.randn(300, 64, device=dev, requires_grad=True)
    y, lad = layer(x)
    (y.sum() + lad.sum()).backward()
    assert torch.isfinite(x.grad).all()


@case("conditioner")
def _conditioner():
    # fused conditioner (tcgen05, CTA pairs) vs the per-layer tensor-core path, forward and inverse
    wl = workloads.get_workload("cfg2_tc_small")
    flow = workloads.build_flow(wl, seed=0).to(dev).eval()
    x = torch.randn(700, 64, device=dev)
    with torch.no_grad():
        a = flow.log_prob(x)
        tensorcore.FUSED_CONDITIONER = False
        b = flow.log_prob(x)  # linear_tf32x3_kernel (store + spline epilogues)
        tensorcore.FUSED_CONDITIONER = True
        z, _ = flow._transform(x)
        xi, _ = flow._transform.inverse(z)
    assert (a - b).abs().max() < 1e-2, float((a - b).abs().max())
    assert (xi - x).abs().max() < 1e-2


@case("made_inverse")
def _made_inverse():
    torch.manual_seed(1)
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
        6, 64, num_bins=8, tails="linear", tail_bound=3.0, num_blocks=2).to(dev).eval()
    z = torch.randn(70, 6, device=dev)
    with torch.no_grad():
        x, lad = layer.inverse(z)
        made_inverse.ENABLED = False
        xd, ladd = layer.inverse(z)
        made_inverse.ENABLED = True
    assert (x - xd).abs().max() < 1e-3 and (lad - ladd).abs().max() < 1e-2


@case("train_step")
def _train_step():
    # tensor-core autograd path (row-major staged GEMM, split-K weight gradient) on a small MAF
    torch.manual_seed(2)
    layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(
        8, 64, num_bins=8, tails="linear", tail_bound=3.0, num_blocks=1).to(dev)
    x = torch.randn(512, 8, device=dev)
    y, lad = layer(x)
    (-(lad.mean()) + (y ** 2).mean()).backward()
    assert all(torch.isfinite(p.grad).all() for p in layer.parameters())
