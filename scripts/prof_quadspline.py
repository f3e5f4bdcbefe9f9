"""Launch the piecewise-quadratic spline kernels on coupling shapes (for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import _cabi, ops
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
B = 1 << 20
x = torch.randn(B, 64, generator=g, device=dev); p = torch.randn(B, 32 * 15, generator=g, device=dev)
gy, gl = torch.randn_like(x), torch.randn(B, device=dev)
tc = torch.arange(0, 64, 2, dtype=torch.int32, device=dev); cc = torch.arange(1, 64, 2, dtype=torch.int32, device=dev)
quad = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1.0 / 16.0)
for _ in range(2):
    ops.quadspline_layer(x, p, tc, cc, *quad)
    ops.quadspline_layer_backward(x, p, gy, gl, tc, cc, *quad)
torch.cuda.synchronize()
print("ok")
