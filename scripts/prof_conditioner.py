"""Launch the fused conditioner kernel a few times on the cfg-2 shapes (for ncu / compute-sanitizer):
    python scripts/prof_conditioner.py [rows]      # default 1M rows, four launches of flow layer 0"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import conditioner, workloads  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda:0")
wl = workloads.get_workload("cfg2")
flow = workloads.build_flow(wl)
flow.load_state_dict(workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl))
flow = flow.to(dev)
x = torch.randn(rows, 64, generator=torch.Generator(device=dev).manual_seed(0), device=dev)
layer = flow._transform._transforms[0]
with torch.no_grad():
    for _ in range(4):
        y, lad = layer(x)
torch.cuda.synchronize()
assert conditioner.kernel_error() == 0 and bool(torch.isfinite(lad).all())
print("ok", float(lad.double().sum()))
prof = conditioner.kernel_profile()
if prof.get("mma total"):
    import json
    print(json.dumps(prof))
