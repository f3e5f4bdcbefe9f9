#!/usr/bin/env python
"""Forward element-wise kernels: the CTA-level tile ring (tiled_apply_kernel, FC_TILE=1) against the per-warp ring
(pipelined_apply_kernel, FC_TILE=0) — algorithmic GB/s, bit-identity of the two, and (--sweep) the tile-ring knobs."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flowconductor_b200 import _cabi, ops  # noqa: E402
from scripts.bench_kernels import PEAK, rqs_case, timeit  # noqa: E402


def set_env(**kw):
    for k in list(os.environ):
        if k.startswith("FC_TILE") or k.startswith("FC_PIPE"):
            del os.environ[k]
    for k, v in kw.items():
        os.environ[k] = str(v)


def cases(B):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    out = []
    for name, (a, nbytes) in (("rqs_fwd cfg2 D=64 K=8 coupling", rqs_case(B, 64, 8)),
                              ("rqs_inv cfg2 D=64 K=8 coupling", rqs_case(B, 64, 8, inverse=True)),
                              ("rqs_fwd cfg5 D=256 K=8 coupling", rqs_case(B // 4, 256, 8)),
                              ("rqs_fwd cfg3 D=16 K=16 autoregressive", rqs_case(B, 16, 16, coupling=False))):
        out.append((name, (lambda a=a: ops.rqs_layer(*a)), nbytes))
    xs = torch.randn(B, 32, generator=g, device=dev)
    ps = torch.randn(B, 32 * 31, generator=g, device=dev)
    out.append(("sos_fwd cfg4 D=32 n=10", lambda: ops.sos_layer(xs, ps, 10, 0.0, False, 50, 120.0), B * (4 * 992 + 8 * 32 + 4)))
    xa = torch.randn(B * 4, 64, generator=g, device=dev)
    pa = torch.randn(B * 4, 64, generator=g, device=dev)
    tca = torch.arange(0, 64, 2, dtype=torch.int32, device=dev)
    cca = torch.arange(1, 64, 2, dtype=torch.int32, device=dev)
    out.append(("affine_fwd coupling D=64 (4M rows)",
                lambda: ops.affine_layer(xa, pa, tca, cca, _cabi.AFFINE_BLOCKED, _cabi.SCALE_SIGMOID2, False),
                xa.shape[0] * (4 * 64 * 3 + 4)))
    xl = torch.randn(B, 64, generator=g, device=dev)
    pl = torch.randn(B, 32 * 8, generator=g, device=dev)
    lin = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0)
    out.append(("linspline_fwd D=64 K=8 coupling", lambda: ops.linspline_layer(xl, pl, tca, cca, *lin),
                B * (4 * 64 + 4 * 256 + 4 * 64 + 4)))
    pq = torch.randn(B, 32 * 15, generator=g, device=dev)
    quad = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1.0 / 16.0)
    out.append(("quadspline_fwd D=64 K=8 coupling", lambda: ops.quadspline_layer(xl, pq, tca, cca, *quad),
                B * (4 * 64 + 4 * 480 + 4 * 64 + 4)))
    pc = torch.randn(B, 32 * 18, generator=g, device=dev)
    out.append(("cubicspline_fwd D=64 K=8 coupling", lambda: ops.cubicspline_layer(xl, pc, tca, cca, *quad),
                B * (4 * 64 + 4 * 576 + 4 * 64 + 4)))
    return out


def backward_cases(B):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(2)
    out = []
    for name, (a, _) in (("rqs_bwd cfg2 D=64 K=8 coupling", rqs_case(B, 64, 8)),
                         ("rqs_bwd cfg3 D=16 K=16 autoregressive", rqs_case(B, 16, 16, coupling=False))):
        x, p, tc, cc = a[:4]
        rest = a[4:]
        gy = torch.randn_like(x)
        gl = torch.randn(x.shape[0], device=dev)
        out.append((name, (lambda x=x, p=p, gy=gy, gl=gl, tc=tc, cc=cc, rest=rest: ops.rqs_layer_backward(x, p, gy, gl, tc, cc, *rest)),
                    x.shape[0] * (2 * 4 * p.shape[1] + 3 * 4 * x.shape[1] + 4)))
    xs = torch.randn(B, 32, generator=g, device=dev)
    ps = torch.randn(B, 32 * 31, generator=g, device=dev)
    gys, gls = torch.randn_like(xs), torch.randn(B, device=dev)
    out.append(("sos_bwd cfg4 D=32 n=10", lambda: ops.sos_layer_backward(xs, ps, gys, gls, 10), B * (8 * 992 + 12 * 32 + 4)))
    xa = torch.randn(B * 4, 64, generator=g, device=dev)
    pa = torch.randn(B * 4, 64, generator=g, device=dev)
    tca = torch.arange(0, 64, 2, dtype=torch.int32, device=dev)
    cca = torch.arange(1, 64, 2, dtype=torch.int32, device=dev)
    gya, gla = torch.randn_like(xa), torch.randn(xa.shape[0], device=dev)
    out.append(("affine_bwd coupling D=64 (4M rows)",
                lambda: ops.affine_layer_backward(xa, pa, gya, gla, tca, cca, _cabi.AFFINE_BLOCKED, _cabi.SCALE_SIGMOID2, False),
                xa.shape[0] * (4 * 64 * 5 + 4)))
    xl = torch.randn(B, 64, generator=g, device=dev)
    gyl, gll = torch.randn_like(xl), torch.randn(B, device=dev)
    pl = torch.randn(B, 32 * 8, generator=g, device=dev)
    lin = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0)
    out.append(("linspline_bwd D=64 K=8 coupling", lambda: ops.linspline_layer_backward(xl, pl, gyl, gll, tca, cca, *lin),
                B * (4 * 64 * 3 + 4 * 256 * 2 + 4)))
    quad = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1.0 / 16.0)
    pq = torch.randn(B, 32 * 15, generator=g, device=dev)
    out.append(("quadspline_bwd D=64 K=8 coupling", lambda: ops.quadspline_layer_backward(xl, pq, gyl, gll, tca, cca, *quad),
                B * (4 * 64 * 3 + 4 * 480 * 2 + 4)))
    pc = torch.randn(B, 32 * 18, generator=g, device=dev)
    out.append(("cubicspline_bwd D=64 K=8 coupling", lambda: ops.cubicspline_layer_backward(xl, pc, gyl, gll, tca, cca, *quad),
                B * (4 * 64 * 3 + 4 * 576 * 2 + 4)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1 << 20)
    ap.add_argument("--sweep", type=str, default="")
    ap.add_argument("--backward", action="store_true")
    args = ap.parse_args()
    cs = cases(args.B) if not args.backward else backward_cases(args.B)
    pre = "FC_TILE_BWD_" if args.backward else "FC_TILE_"
    for name, fn, nbytes in cs:
        res = {}
        for label, env in (("per-warp ring", {"FC_TILE": 0}), ("tile ring", {})):
            set_env(**env)
            med, best = timeit(fn)
            res[label] = fn()
            gbs = nbytes / med / 1e6
            print(json.dumps({"kernel": name, "variant": label, "ms_median": round(med, 4), "GB/s": round(gbs, 1),
                              "frac_of_measured_peak": round(gbs / PEAK, 3), "bytes": nbytes}), flush=True)
        same = all(bool(torch.equal(u, v)) for u, v in zip(res["per-warp ring"][:2], res["tile ring"][:2]))
        print("  bit-identical:", same, flush=True)
    for pat in [p for p in args.sweep.split(",") if p]:
        for name, fn, nbytes in cs:
            if pat not in name:
                continue
            for warps, ctas, stages, passes in itertools.product((8, 10, 12, 15, 16), (1, 2), (2, 3), (8, 32)):
                set_env(**{pre + "WARPS": warps, pre + "CTAS": ctas, pre + "STAGES": stages, pre + "PASSES": passes})
                med, _ = timeit(fn, warm=2, reps=7)
                gbs = nbytes / med / 1e6
                print("sweep %-34s warps=%2d ctas=%d stages=%d passes=%d  %.3f ms  %.3f of peak" % (
                    name, warps, ctas, stages, passes, med, gbs / PEAK), flush=True)
    set_env()


if __name__ == "__main__":
    main()
