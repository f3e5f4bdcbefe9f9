"""Launch the two training-path products on cfg-3 shapes (for ncu): the row-major staged forward GEMM with a skip
connection and the untransposed-grad_y weight-gradient product (262144 rows, 256 x 256)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import linear as fl
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(1)
B, H = 262144, 256
x = torch.randn(B, H, generator=g, device=dev); gy = torch.randn(B, H, generator=g, device=dev)
w = torch.randn(H, H, generator=g, device=dev) / 16; b = torch.randn(H, generator=g, device=dev)
pk = fl.pack(w, b)
for _ in range(3):
    out = fl.linear(x, pk, relu_in=True, residual=x)
    xt = fl.pack_transposed(x, relu=True)
    gw, gb = fl.linear_splitk_t(gy, xt, column_sums=True)
torch.cuda.synchronize()
print("ok")
