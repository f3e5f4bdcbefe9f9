"""Per-layer error of the tensor-core path and of the unfused (cuBLAS fp32 + element-wise kernel) path against the
fp64 oracle, next to the fp32 oracle's own error.  cfg 2 trained-like weights, rows of the full-size batch."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import workloads
from flowconductor_b200.nn import tensorcore
from oracle import restated

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
wl = workloads.get_workload(name)
flow = workloads.build_flow(wl)
state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl)
flow.load_state_dict(state)
specs = workloads.oracle_specs(wl)
state64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
flow = flow.to(dev)
D = wl["features"]
x = torch.randn(B, D, generator=torch.Generator().manual_seed(99))
layers = list(flow._transform._transforms)
q = torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999, 1.0], dtype=torch.float64)


def stats(e):
    return " ".join("%.1e" % v for v in torch.quantile(e.double().flatten(), q).tolist())


h = x
with torch.no_grad():
    for li, (layer, spec) in enumerate(zip(layers, specs)):
        y64, l64 = restated.apply_layer(state64, spec, h.double())
        y32, l32 = restated.apply_layer(state, spec, h)
        hg = h.to(dev)
        tensorcore.ENABLED = True
        y_tc, l_tc = layer(hg)
        tensorcore.ENABLED = False
        y_un, l_un = layer(hg)
        print("layer %d %s  (quantiles 50/90/99/99.9/99.99/max of |err vs fp64|)" % (li, spec["kind"]))
        for nm, yy, ll in (("oracle32", y32, l32), ("unfused ", y_un.cpu(), l_un.cpu()), ("tensorcr", y_tc.cpu(), l_tc.cpu())):
            print("   %s  y: %s   lad: %s" % (nm, stats((yy.double() - y64).abs()), stats((ll.double() - l64).abs())))
        h = y32
