"""Cycle-counter profile of the store variant (profile build of the library via FC_LIB)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from flowconductor_b200 import conditioner, transforms, workloads  # noqa: E402
from flowconductor_b200.nn import tensorcore  # noqa: E402
from flowconductor_b200.nn.nets import ResidualNet  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "store"
layer = transforms.PiecewiseQuadraticCouplingTransform(workloads.make_mask(64, "alternating_even"), lambda i, o: ResidualNet(i, o, hidden_features=256, num_blocks=2), num_bins=8, tails="linear", tail_bound=3.0)
if which == "rqs":
    layer = transforms.PiecewiseRationalQuadraticCouplingTransform(workloads.make_mask(64, "alternating_even"), lambda i, o: ResidualNet(i, o, hidden_features=256, num_blocks=2), num_bins=8, tails="linear", tail_bound=3.0)
layer = layer.to(dev).eval()
x = torch.randn(1 << 20, 64, device=dev)
tensorcore.FUSED_SOS = False
with torch.no_grad():
    for _ in range(3):
        layer(x)
torch.cuda.synchronize()
print(which, json.dumps(conditioner.kernel_profile()))
