#!/usr/bin/env python
"""Cycle counters of the incremental-MADE inverse kernel (library built with FC_LINEAR_PROFILE_BUILD=1), one cfg-3 layer."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import made_inverse, workloads  # noqa: E402

dev = torch.device("cuda:0")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32 * 148
wl = workloads.get_workload("cfg3")
flow = workloads.build_flow(wl, seed=0)
state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl, seed=1)
flow.load_state_dict(state)
layer = [t for t in flow._transform._transforms if hasattr(t, "autoregressive_net")][0].to(dev).eval()
z = torch.randn(rows, 16, device=dev)
with torch.no_grad():
    for _ in range(3):
        layer.inverse(z)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    layer.inverse(z)
    e.record()
    torch.cuda.synchronize()
print(json.dumps({"rows": rows, "ms": s.elapsed_time(e), "cycles": made_inverse.kernel_profile()}))
