"""A/B of the whole-conditioner store kernel (fc_conditioner_store_apply behind tensorcore.params) against the per-layer
tensor-core kernels, for layers whose bijection runs as an element-wise kernel: forward of one layer at 1 M rows."""
import json
import sys

import torch

sys.path.insert(0, ".")
from flowconductor_b200 import transforms, workloads  # noqa: E402
from flowconductor_b200.nn import tensorcore  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402
from scripts.bench_affine_fused import timed  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    rows = 1 << 20
    cases = []
    for H in (256, 128):
        net = lambda i, o, H=H: ResidualNet(i, o, hidden_features=H, num_blocks=2)  # noqa: E731
        mask = workloads.make_mask(64, "alternating_even")
        cases += [("quadratic coupling D=64 K=8", H, transforms.PiecewiseQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0), None),
                  ("linear coupling D=64 K=8", H, transforms.PiecewiseLinearCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0), None),
                  ("cubic coupling D=64 K=8", H, transforms.PiecewiseCubicCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0), None)]
    cases.append(("quadratic MAF D=16 K=8", 256, transforms.MaskedPiecewiseQuadraticAutoregressiveTransform(
        16, 256, num_bins=8, tails="linear", tail_bound=3.0), None))
    cases.append(("conditional sum of sigmoids D=32 n=10 ctx=8 (262144 rows)", 64,
                  transforms.ConditionalSumOfSigmoidsTransform(32, 64, context_features=8, n_sigmoids=10, num_blocks=2), 262144))
    for name, H, layer, r in cases:
        torch.manual_seed(0)
        layer = layer.to(dev).eval()
        n = r or rows
        D = 32 if "conditional" in name else (16 if "MAF" in name else 64)
        x = torch.randn(n, D, device=dev)
        ctx = torch.randn(n, 8, device=dev) if "conditional" in name else None
        out = {"layer": name, "hidden": H, "rows": n}
        with torch.no_grad():
            variants = [("store_kernel_ms", dict(FUSED_STORE=True, FUSED_SOS=False)), ("perlayer_ms", dict(FUSED_STORE=False, FUSED_SOS=False))]
            if "conditional" in name:
                variants.append(("fused_sos_ms", dict(FUSED_STORE=True, FUSED_SOS=True)))
            for key, flags in variants:
                for k, v in flags.items():
                    setattr(tensorcore, k, v)
                out[key] = round(timed(lambda: layer(x, ctx)), 3)
        tensorcore.FUSED_STORE, tensorcore.FUSED_SOS = True, True
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
