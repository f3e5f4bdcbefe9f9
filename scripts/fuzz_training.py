#!/usr/bin/env python
"""Random-shape stress of the training path (nn/tc_autograd.py: tensor-core forward / input-gradient / weight-gradient
kernels, fused residual blocks) against the same model on cuBLAS fp32 (tc_autograd.ENABLED = False): loss and every
parameter gradient, ragged batches, coupling and autoregressive layers."""
import copy
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import transforms, workloads  # noqa: E402
from flowconductor_b200.nn import tc_autograd  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    worst = 0.0
    for case in range(n_cases):
        fam = random.choice(["maf_rqs", "coupling_rqs", "maf_affine", "coupling_lin"])
        D = random.choice([4, 8, 16, 24, 64])
        H = random.choice([64, 96, 128, 200, 256])
        blocks = random.choice([1, 2, 3])
        rows = random.choice([1, 100, 128, 129, 1000, 4097, 20000])
        torch.manual_seed(case)
        if fam == "maf_rqs":
            layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(D, H, num_bins=8, tails="linear",
                                                                                       tail_bound=3.0, num_blocks=blocks)
        elif fam == "maf_affine":
            layer = transforms.MaskedAffineAutoregressiveTransform(D, H, num_blocks=blocks)
        else:
            mask = workloads.make_mask(D, "alternating_even")
            create = lambda i, o: ResidualNet(i, o, hidden_features=H, num_blocks=blocks)  # noqa: E731
            cls = (transforms.PiecewiseRationalQuadraticCouplingTransform if fam == "coupling_rqs"
                   else transforms.PiecewiseLinearCouplingTransform)
            layer = cls(mask, create, num_bins=8, tails="linear", tail_bound=3.0)
        layer = layer.to(dev)
        with torch.no_grad():
            for p in layer.parameters():
                p.add_(torch.randn_like(p) * 0.05)
        ref = copy.deepcopy(layer)
        x = torch.randn(rows, D, device=dev)
        out = {}
        for name, model, on in (("tc", layer, True), ("cublas", ref, False)):
            tc_autograd.ENABLED = on
            try:
                y, lad = model(x)
                loss = -(lad.mean()) + 0.5 * (y ** 2).sum(1).mean()
                loss.backward()
                out[name] = (loss.item(), [p.grad.clone() for p in model.parameters()])
            finally:
                tc_autograd.ENABLED = True
        el = abs(out["tc"][0] - out["cublas"][0]) / max(1.0, abs(out["cublas"][0]))
        eg = 0.0
        for g, r in zip(out["tc"][1], out["cublas"][1]):
            eg = max(eg, ((g - r).abs().max() / r.abs().max().clamp_min(1e-6)).item())
        worst = max(worst, eg)
        bad = el > 1e-4 or eg > 2e-3 or not all(torch.isfinite(g).all() for g in out["tc"][1])
        print("case %2d %-12s D=%2d H=%3d blocks=%d rows=%5d: loss %.1e, gradients (max over parameters, relative to each max) %.1e%s" % (
            case, fam, D, H, blocks, rows, el, eg, "   <<<<<< MISMATCH" if bad else ""))
    print("worst relative gradient difference %.2e" % worst)


if __name__ == "__main__":
    main()
