"""One launch each of the tile-ring kernels worth a profile (for `ncu --set full -k regex:tiled_`): RQ-spline forward and
backward at the cfg 2 shapes, linear-spline and quadratic-spline forward."""
import sys

import torch

sys.path.insert(0, ".")
from scripts.bench_tile_ring import backward_cases, cases  # noqa: E402

B = 1 << 20
for name, fn, _ in cases(B):
    if name.startswith(("rqs_fwd cfg2", "linspline_fwd", "quadspline_fwd", "affine_fwd")):
        fn()
for name, fn, _ in backward_cases(B):
    if name.startswith(("rqs_bwd cfg2", "quadspline_bwd")):
        fn()
torch.cuda.synchronize()
