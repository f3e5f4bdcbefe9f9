"""One launch each of the fused-conditioner variants added late in round 2 (for `ncu --set full -k regex:conditioner_f16x3`):
store (quadratic coupling, H = 256), affine coupling (H = 256), store on the cfg-4 shapes."""
import sys

import torch

sys.path.insert(0, ".")
from flowconductor_b200 import transforms, workloads  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402

dev = torch.device("cuda:0")
net = lambda i, o: ResidualNet(i, o, hidden_features=256, num_blocks=2)  # noqa: E731
mask = workloads.make_mask(64, "alternating_even")
layers = [(transforms.PiecewiseQuadraticCouplingTransform(mask, net, num_bins=8, tails="linear", tail_bound=3.0), 1 << 20, 64, None),
          (transforms.AffineCouplingTransform(mask, net), 1 << 20, 64, None),
          (transforms.ConditionalSumOfSigmoidsTransform(32, 64, context_features=8, n_sigmoids=10, num_blocks=2), 262144, 32, 8)]
with torch.no_grad():
    for layer, n, D, c in layers:
        layer = layer.to(dev).eval()
        x = torch.randn(n, D, device=dev)
        ctx = torch.randn(n, c, device=dev) if c else None
        if c:
            layer.inverse(x, ctx)  # store + numerical inverse
        else:
            layer(x)
torch.cuda.synchronize()
