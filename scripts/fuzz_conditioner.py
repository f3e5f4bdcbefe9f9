#!/usr/bin/env python
"""Random-shape stress of the fused conditioner kernel (csrc/fc_conditioner.cu): coupling / autoregressive / conditional
layers with rational-quadratic splines (8 / 10 / 16 bins, with and without tails) and the conditional sum of sigmoids, hidden
widths 64..256 (padded to the kernel's 128 / 256), 1..4 blocks, ragged batches — against the unfused path (torch conditioner +
element-wise kernel) in both directions."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import _cabi, transforms, workloads  # noqa: E402
from flowconductor_b200.nn import tensorcore  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    worst, ran, skipped = 0.0, 0, 0
    for case in range(n_cases):
        fam = random.choice(["coupling", "coupling", "maf", "cond_rqs", "cond_sos", "lin", "quad", "cubic", "affine", "maf_affine"])
        D = random.choice([4, 6, 8, 12, 16, 21, 32, 43, 63, 64, 100])
        H = random.choice([64, 68, 100, 128, 132, 200, 256])
        blocks = random.choice([1, 2, 3, 4])
        K = random.choice([8, 10, 16])
        tails = random.choice(["linear", "linear", None])
        rows = random.choice([1, 127, 128, 129, 300, 1000, 5000])
        torch.manual_seed(case)
        ctx = None
        if fam == "coupling":
            mask = workloads.make_mask(D, random.choice(["alternating_even", "alternating_odd", "mid_split"]))
            layer = transforms.PiecewiseRationalQuadraticCouplingTransform(
                mask, lambda i, o: ResidualNet(i, o, hidden_features=H, num_blocks=blocks), num_bins=K, tails=tails,
                tail_bound=3.0)
            box = (0.0, 1.0)
        elif fam == "maf":
            layer = transforms.MaskedPiecewiseRationalQuadraticAutoregressiveTransform(D, H, num_bins=K, tails=tails,
                                                                                       tail_bound=3.0, num_blocks=blocks)
            box = (-1.2, 1.2)
        elif fam == "cond_rqs":
            layer = transforms.ConditionalPiecewiseRationalQuadraticTransform(D, H, context_features=8, num_bins=K, tails=tails,
                                                                              tail_bound=3.0, num_blocks=blocks)
            ctx = torch.randn(rows, 8, device=dev)
            box = (-1.2, 1.2)
        elif fam in ("lin", "quad", "cubic", "affine"):
            # the other coupling families: per-layer tensor-core kernels (T128 activations), element-wise or fused-affine end
            mask = workloads.make_mask(D, random.choice(["alternating_even", "mid_split"]))
            create = lambda i, o: ResidualNet(i, o, hidden_features=H, num_blocks=blocks)  # noqa: E731
            tails, box = "linear", None
            if fam == "lin":
                layer = transforms.PiecewiseLinearCouplingTransform(mask, create, num_bins=K, tails="linear", tail_bound=3.0)
            elif fam == "quad":
                layer = transforms.PiecewiseQuadraticCouplingTransform(mask, create, num_bins=K, tails="linear", tail_bound=3.0)
            elif fam == "cubic":
                layer = transforms.PiecewiseCubicCouplingTransform(mask, create, num_bins=K, tails="linear", tail_bound=3.0)
            else:
                layer = transforms.AffineCouplingTransform(mask, create)
        elif fam == "maf_affine":
            layer = transforms.MaskedAffineAutoregressiveTransform(D, H, num_blocks=blocks)
            tails, box = "linear", None
        else:
            layer = transforms.ConditionalSumOfSigmoidsTransform(D, H, context_features=8, n_sigmoids=10, num_blocks=blocks)
            ctx = torch.randn(rows, 8, device=dev)
            tails, box = "linear", None
        layer = layer.to(dev).eval()
        with torch.no_grad():
            for p in layer.parameters():
                p.add_(torch.randn_like(p) * 0.05)
            if tails is None:
                x = torch.rand(rows, D, device=dev) * (box[1] - box[0]) * 0.96 + box[0] + 0.02 * (box[1] - box[0])
            else:
                x = torch.randn(rows, D, device=dev)
            res = {}
            for inverse in ((False,) if fam in ("cond_sos", "maf_affine", "maf") else (False, True)):
                fn = layer.inverse if inverse else layer
                _cabi.STATS.reset()
                try:
                    y, lad = fn(x, ctx)
                    fused = any(k.startswith("fc_conditioner_") and k.endswith("_apply") for k in _cabi.STATS.counts)
                    path = "fused" if fused else ("per-layer" if any(k.startswith("fc_linear_") for k in _cabi.STATS.counts)
                                                  else "unfused")
                    tensorcore.ENABLED = False
                    yu, ladu = fn(x, ctx)
                except Exception as e:  # noqa: BLE001
                    print("case %2d %-9s D=%3d H=%3d blocks=%d K=%2d tails=%s rows=%4d inverse=%s: raised %s: %s   <<<<<< MISMATCH" % (
                        case, fam, D, H, blocks, K, tails, rows, inverse, type(e).__name__, str(e)[:100]))
                    tensorcore.ENABLED = True
                    continue
                finally:
                    tensorcore.ENABLED = True
                if path == "unfused":
                    skipped += 1
                    continue
                ran += 1
                ey = ((y - yu).abs() / yu.abs().clamp_min(1.0))
                el = ((lad - ladu).abs() / ladu.abs().clamp_min(1.0))
                q = float(torch.quantile(ey.flatten()[: 1 << 20], 0.999))
                worst = max(worst, q)
                bad = (q > 1e-4 or float(torch.quantile(el, 0.99)) > 1e-3 or not bool(torch.isfinite(y).all())
                       or float(ey.max()) > 5e-2)
                print("case %2d %-9s %-9s D=%3d H=%3d blocks=%d K=%2d tails=%-6s rows=%4d inverse=%d: outputs p99.9 %.1e max %.1e, logabsdet p99 %.1e%s" % (
                    case, fam, path, D, H, blocks, K, tails, rows, inverse, q, float(ey.max()), float(torch.quantile(el, 0.99)),
                    "   <<<<<< MISMATCH" if bad else ""))
    print("ran %d tensor-core calls (torch conditioner only: %d), worst p99.9 relative output difference %.2e" % (ran, skipped, worst))


if __name__ == "__main__":
    main()
