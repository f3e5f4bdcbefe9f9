"""A/B of the fused affine conditioner (fc_conditioner_affine_apply) against the per-layer tensor-core kernels
(fc_linear_* + fc_linear_affine_apply): forward of an affine coupling layer and a masked affine autoregressive layer."""
import json
import sys

import torch

sys.path.insert(0, ".")
from flowconductor_b200 import transforms, workloads  # noqa: E402
from flowconductor_b200.nn import tensorcore  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device("cuda:0")
    rows = 1 << 20
    for kind, D, H in (("coupling", 64, 256), ("coupling", 64, 128), ("maf", 64, 256), ("maf", 16, 128)):
        torch.manual_seed(0)
        if kind == "coupling":
            layer = transforms.AffineCouplingTransform(
                workloads.make_mask(D, "alternating_even"),
                lambda i, o: ResidualNet(i, o, hidden_features=H, num_blocks=2))
        else:
            layer = transforms.MaskedAffineAutoregressiveTransform(features=D, hidden_features=H, num_blocks=2)
        layer = layer.to(dev).eval()
        x = torch.randn(rows, D, device=dev)
        out = {"layer": kind, "features": D, "hidden": H, "rows": rows}
        with torch.no_grad():
            for name, flag in (("fused_ms", True), ("perlayer_ms", False)):
                tensorcore.FUSED_AFFINE = flag
                out[name] = round(timed(lambda: layer(x)), 3)
        tensorcore.FUSED_AFFINE = True
        out["speedup"] = round(out["perlayer_ms"] / out["fused_ms"], 2)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
