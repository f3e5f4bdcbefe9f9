#!/usr/bin/env python
"""Direct-kernel microbenchmarks: params ~ N(0,1)*2 laid out [B, D_t*P], x ~ N(0,1), timed with
CUDA events (5 warm-up, median of 20), working sets >> L2.  Prints achieved algorithmic GB/s per kernel and,
with --sweep, the tuning grid of the pipelined kernel (env knobs FC_PIPE_*)."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flowconductor_b200 import _cabi, ops  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, warm=5, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def rqs_case(B, D, K, coupling=True, inverse=False):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    d_t = D // 2 if coupling else D
    P = 3 * K - 1
    x = torch.randn(B, D, generator=g, device=dev)
    p = torch.randn(B, d_t * P, generator=g, device=dev) * 2
    tc = torch.arange(0, D, 2, dtype=torch.int32, device=dev) if coupling else None
    cc = torch.arange(1, D, 2, dtype=torch.int32, device=dev) if coupling else None
    args = (x, p, tc, cc, K, _cabi.TAILS_LINEAR, inverse, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1e-3, 1.0 / 16)
    nbytes = B * (4 * d_t * P + 4 * D + 4 * D + 4)  # params + full x row in + full y row out + lad
    return args, nbytes


def set_env(**kw):
    for k in list(os.environ):
        if k.startswith("FC_PIPE") or k.startswith("FC_TILE"):
            del os.environ[k]
    for k, v in kw.items():
        os.environ[k] = str(v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--B", type=int, default=1 << 20)
    ap.add_argument("--sweep-case", type=int, default=0)
    args = ap.parse_args()
    out = []
    cases = [("rqs_fwd cfg2 D=64 K=8 coupling", rqs_case(args.B, 64, 8)),
             ("rqs_inv cfg2 D=64 K=8 coupling", rqs_case(args.B, 64, 8, inverse=True)),
             ("rqs_fwd cfg5 D=256 K=8 coupling", rqs_case(args.B // 4, 256, 8)),
             ("rqs_fwd cfg3 D=16 K=16 autoregressive", rqs_case(args.B, 16, 16, coupling=False))]
    for name, (a, nbytes) in cases:
        for label, env in (("staged", {"FC_PIPE": 0, "FC_TILE": 0}), ("tile ring", {})):
            set_env(**env)
            med, best = timeit(lambda: ops.rqs_layer(*a))
            gbs = nbytes / med / 1e6
            rec = {"kernel": name, "variant": label, "ms_median": med, "ms_best": best, "GB/s": gbs,
                   "frac_of_measured_peak": gbs / PEAK, "bytes": nbytes}
            out.append(rec)
            print(json.dumps(rec), flush=True)
        # the two variants must agree bit for bit (same element arithmetic)
        set_env(FC_PIPE=0, FC_TILE=0)
        y0, l0, _ = ops.rqs_layer(*a)
        set_env()
        y1, l1, _ = ops.rqs_layer(*a)
        print("  bit-identical:", bool(torch.equal(y0, y1) and torch.equal(l0, l1)), flush=True)
    # backward (fc_rqs_backward): x, params, grad_y, grad_lad in; grad_x, grad_params out
    for name, (a, _) in (cases[0], cases[3]):
        x, p, tc, cc = a[:4]
        rest = a[4:]
        gy = torch.randn_like(x)
        gl = torch.randn(x.shape[0], device=x.device)
        nbytes = x.shape[0] * (2 * 4 * p.shape[1] + 3 * 4 * x.shape[1] + 4)
        res = {}
        for label, env in (("staged", {"FC_PIPE_BWD": 0, "FC_TILE_BWD": 0}), ("tile ring", {})):
            set_env(**env)
            os.environ.pop("FC_PIPE_BWD", None)
            for k, v in env.items():
                os.environ[k] = str(v)
            med, best = timeit(lambda: ops.rqs_layer_backward(x, p, gy, gl, tc, cc, *rest))
            res[label] = ops.rqs_layer_backward(x, p, gy, gl, tc, cc, *rest)
            gbs = nbytes / med / 1e6
            rec = {"kernel": name.replace("rqs_fwd", "rqs_bwd"), "variant": label, "ms_median": med, "ms_best": best,
                   "GB/s": gbs, "frac_of_measured_peak": gbs / PEAK, "bytes": nbytes}
            print(json.dumps(rec), flush=True)
        os.environ.pop("FC_PIPE_BWD", None)
        print("  bit-identical:", all(bool(torch.equal(u, v)) for u, v in zip(res["staged"], res["tile ring"])), flush=True)
    # sum-of-sigmoids (cfg 4 shapes: D = 32, n = 10, P = 31) forward / inverse / backward, affine coupling (16 B/element)
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    Bs, Ds, ns = args.B, 32, 10
    xs = torch.randn(Bs, Ds, generator=g, device=dev)
    ps = torch.randn(Bs, Ds * (3 * ns + 1), generator=g, device=dev)
    gys, gls = torch.randn_like(xs), torch.randn(Bs, device=dev)
    set_env()
    for name, fn, nbytes in (
            ("sos_fwd cfg4 D=32 n=10", lambda: ops.sos_layer(xs, ps, ns, 0.0, False, 50, 120.0), Bs * (4 * ps.shape[1] + 8 * Ds + 4)),
            ("sos_inv cfg4 D=32 n=10 (safeguarded Newton)", lambda: ops.sos_layer(xs, ps, ns, 0.0, True, 50, 120.0), Bs * (4 * ps.shape[1] + 8 * Ds + 4)),
            ("sos_bwd cfg4 D=32 n=10", lambda: ops.sos_layer_backward(xs, ps, gys, gls, ns), Bs * (8 * ps.shape[1] + 12 * Ds + 4))):
        med, best = timeit(fn)
        gbs = nbytes / med / 1e6
        print(json.dumps({"kernel": name, "variant": "tile ring", "ms_median": med, "ms_best": best, "GB/s": gbs,
                          "frac_of_measured_peak": gbs / PEAK, "bytes": nbytes}), flush=True)
    xa = torch.randn(args.B * 4, 64, generator=g, device=dev)
    pa = torch.randn(args.B * 4, 64, generator=g, device=dev)
    tca = torch.arange(0, 64, 2, dtype=torch.int32, device=dev)
    cca = torch.arange(1, 64, 2, dtype=torch.int32, device=dev)
    med, best = timeit(lambda: ops.affine_layer(xa, pa, tca, cca, _cabi.AFFINE_BLOCKED, _cabi.SCALE_SIGMOID2, False))
    nbytes = xa.shape[0] * (4 * 64 + 4 * 64 + 4 * 64 + 4)
    print(json.dumps({"kernel": "affine_fwd coupling D=64 (4M rows)", "variant": "tile ring", "ms_median": med, "ms_best": best,
                      "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK, "bytes": nbytes}), flush=True)
    gya, gla = torch.randn_like(xa), torch.randn(xa.shape[0], device=dev)
    med, best = timeit(lambda: ops.affine_layer_backward(xa, pa, gya, gla, tca, cca, _cabi.AFFINE_BLOCKED,
                                                         _cabi.SCALE_SIGMOID2, False))
    nbytes = xa.shape[0] * (4 * 64 * 5 + 4)
    print(json.dumps({"kernel": "affine_bwd coupling D=64 (4M rows)", "variant": "tile ring", "ms_median": med,
                      "ms_best": best, "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK,
                      "bytes": nbytes}), flush=True)
    # piecewise-linear spline, coupling shapes of cfg 2 (D = 64, 32 transformed + 32 copied, K = 8): 4 (K + 2) B/element
    xl = torch.randn(args.B, 64, generator=g, device=dev)
    pl = torch.randn(args.B, 32 * 8, generator=g, device=dev)
    gyl, gll = torch.randn_like(xl), torch.randn(args.B, device=dev)
    lin = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0)
    for name, fn, nbytes in (
            ("linspline_fwd D=64 K=8 coupling", lambda: ops.linspline_layer(xl, pl, tca, cca, *lin),
             args.B * (4 * 64 + 4 * 256 + 4 * 64 + 4)),
            ("linspline_bwd D=64 K=8 coupling", lambda: ops.linspline_layer_backward(xl, pl, gyl, gll, tca, cca, *lin),
             args.B * (4 * 64 * 3 + 4 * 256 * 2 + 4))):
        med, best = timeit(fn)
        print(json.dumps({"kernel": name, "variant": "tile ring", "ms_median": med, "ms_best": best,
                          "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK,
                          "bytes": nbytes}), flush=True)
    # piecewise-quadratic spline, same coupling shapes: K = 8 with linear tails -> P = 15, 4 (P + 2) B/element
    pq = torch.randn(args.B, 32 * 15, generator=g, device=dev)
    quad = (8, _cabi.TAILS_LINEAR, False, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3, 1.0 / 16.0)
    for name, fn, nbytes in (
            ("quadspline_fwd D=64 K=8 coupling", lambda: ops.quadspline_layer(xl, pq, tca, cca, *quad),
             args.B * (4 * 64 + 4 * 480 + 4 * 64 + 4)),
            ("quadspline_bwd D=64 K=8 coupling", lambda: ops.quadspline_layer_backward(xl, pq, gyl, gll, tca, cca, *quad),
             args.B * (4 * 64 * 3 + 4 * 480 * 2 + 4))):
        med, best = timeit(fn)
        print(json.dumps({"kernel": name, "variant": "tile ring", "ms_median": med, "ms_best": best,
                          "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK,
                          "bytes": nbytes}), flush=True)
    # cubic spline, same coupling shapes: K = 8 -> P = 2 K + 2 = 18, 4 (P + 2) B/element (was unmeasured)
    pc = torch.randn(args.B, 32 * 18, generator=g, device=dev)
    for name, fn, nbytes in (
            ("cubicspline_fwd D=64 K=8 coupling", lambda: ops.cubicspline_layer(xl, pc, tca, cca, *quad),
             args.B * (4 * 64 + 4 * 576 + 4 * 64 + 4)),
            ("cubicspline_inv D=64 K=8 coupling (safeguarded Newton)",
             lambda: ops.cubicspline_layer(xl, pc, tca, cca, 8, _cabi.TAILS_LINEAR, True, -3.0, 3.0, -3.0, 3.0, 1e-3, 1e-3,
                                           1.0 / 16.0), args.B * (4 * 64 + 4 * 576 + 4 * 64 + 4)),
            ("cubicspline_bwd D=64 K=8 coupling", lambda: ops.cubicspline_layer_backward(xl, pc, gyl, gll, tca, cca, *quad),
             args.B * (4 * 64 * 3 + 4 * 576 * 2 + 4))):
        med, best = timeit(fn)
        print(json.dumps({"kernel": name, "variant": "tile ring", "ms_median": med, "ms_best": best,
                          "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK,
                          "bytes": nbytes}), flush=True)
    # ActNorm (8 B/element forward; backward reads x and grad_y, writes grad_x: 12 B/element)
    ls, sh = torch.randn(64, device=dev) * 0.3, torch.randn(64, device=dev)
    for name, fn, nbytes in (
            ("actnorm_fwd D=64 (4M rows)", lambda: ops.actnorm_layer(xa, ls, sh, False), xa.shape[0] * (8 * 64 + 4)),
            ("actnorm_bwd D=64 (4M rows)", lambda: ops.actnorm_layer_backward(xa, ls, sh, gya, gla, False),
             xa.shape[0] * (12 * 64 + 4))):
        med, best = timeit(fn)
        print(json.dumps({"kernel": name, "variant": "direct", "ms_median": med, "ms_best": best,
                          "GB/s": nbytes / med / 1e6, "frac_of_measured_peak": nbytes / med / 1e6 / PEAK,
                          "bytes": nbytes}), flush=True)
    if args.sweep:
        a, nbytes = cases[args.sweep_case][1]
        best = None
        for warps, stages, ctas, slot in itertools.product((8, 12, 16), (2, 3, 4), (1, 2), (1, 2, 4)):
            set_env(FC_PIPE_WARPS=warps, FC_PIPE_STAGES=stages, FC_PIPE_CTAS=ctas, FC_PIPE_SLOT_ROWS=slot)
            try:
                med, _ = timeit(lambda: ops.rqs_layer(*a), warm=2, reps=7)
            except Exception as e:  # noqa: BLE001
                print("sweep", warps, stages, ctas, slot, "failed", e)
                continue
            gbs = nbytes / med / 1e6
            print("sweep warps=%d stages=%d ctas=%d slot_rows=%d  %.3f ms  %.0f GB/s  (%.3f of peak)" % (
                warps, stages, ctas, slot, med, gbs, gbs / PEAK), flush=True)
            if best is None or gbs > best[0]:
                best = (gbs, warps, stages, ctas, slot)
        print("BEST", best)
    set_env()


if __name__ == "__main__":
    main()
