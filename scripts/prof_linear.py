"""Launch the tensor-core kernels a few times on cfg-2 shapes (for ncu): hidden linear, final layer + spline."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import _cabi, linear as fl  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
M, H, D, K = 1 << 20, 256, 64, 8
a = torch.randn(M, H, generator=gen, device=dev)
w = torch.randn(H, H, generator=gen, device=dev) / 16
b = torch.randn(H, generator=gen, device=dev)
pk = fl.pack(w, b)
out = torch.empty(M, H, device=dev)
P, ppad, d_t = 23, 24, 32
x = torch.randn(M, D, generator=gen, device=dev)
wf = torch.randn(d_t * P, H, generator=gen, device=dev) / 4
bf = torch.randn(d_t * P, generator=gen, device=dev)
pkf = fl.pack(wf, bf, row_map=fl.grouped_row_map(d_t, P, ppad, dev), n_tile=fl.N_TILE_RQS)
tcols = torch.arange(0, D, 2, device=dev, dtype=torch.int32)
ccols = torch.arange(1, D, 2, device=dev, dtype=torch.int32)
cfg = _cabi.RqsConfig(K, _cabi.TAILS_LINEAR, 0, 0, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 1.0 / math.sqrt(H))
y = x.clone()
lad = torch.zeros(M, device=dev)
at, ot, rt = fl.T128.from_rows(a), fl.T128(M, H, dev), fl.T128.from_rows(out)
for _ in range(3):
    # the launches of one residual block's second layer and of the final layer, as the inference path issues them
    fl.linear(at, pk, relu_in=True, residual=rt, out=ot, out_t128=True)
    fl.linear_rqs(at, pkf, x, x, lad, False, d_t, tcols, ccols, cfg, None)
torch.cuda.synchronize()
print("ok")
