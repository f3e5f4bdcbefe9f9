"""Launch the sum-of-sigmoids forward / backward kernels on cfg-4 shapes (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import ops
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
B = 1 << 20
xs = torch.randn(B, 32, generator=g, device=dev); ps = torch.randn(B, 32 * 31, generator=g, device=dev)
gy, gl = torch.randn_like(xs), torch.randn(B, device=dev)
for _ in range(2):
    ops.sos_layer(xs, ps, 10, 0.0, False, 50, 120.0)
    ops.sos_layer_backward(xs, ps, gy, gl, 10)
torch.cuda.synchronize()
print("ok")
