"""GPU bring-up check of the fused conditioner kernel (csrc/fc_conditioner.cu): per layer of cfg 2 / cfg 3 / cfg 5,
fused kernel vs the per-layer tensor-core path vs the fp32 / fp64 oracle, then a timing of the full-size log_prob with
both paths.   python scripts/check_conditioner.py [--rows N] [--time] [--only cfg2,cfg3]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from flowconductor_b200 import _cabi, conditioner, workloads  # noqa: E402
from flowconductor_b200.nn import tensorcore  # noqa: E402
from oracle import restated  # noqa: E402  (checker only)


def quantiles(e):
    q = torch.quantile(e.flatten().double()[: 1 << 22], torch.tensor([0.5, 0.99, 0.9999], dtype=torch.float64))
    return "p50 %.2e p99 %.2e p99.99 %.2e max %.2e" % (q[0], q[1], q[2], e.max())


def check(name, rows, dev):
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl)
    flow.load_state_dict(state)
    specs = workloads.oracle_specs(wl)
    state64 = {k: (v.double() if v.is_floating_point() else v) for k, v in state.items()}
    flow = flow.to(dev)
    D = wl["features"]
    h = torch.randn(rows, D, generator=torch.Generator().manual_seed(99))
    ok = True
    with torch.no_grad():
        for li, (layer, spec) in enumerate(zip(flow._transform._transforms, specs)):
            y64, l64 = restated.apply_layer(state64, spec, h.double())
            y32, l32 = restated.apply_layer(state, spec, h)
            res = {}
            for fused in (True, False):
                tensorcore.FUSED_CONDITIONER = fused
                _cabi.STATS.reset()
                y, lad = layer(h.to(dev))
                torch.cuda.synchronize()
                res[fused] = (y.cpu(), lad.cpu(), dict(_cabi.STATS.counts))
            err = conditioner.kernel_error()
            yf, lf, cf = res[True]
            yu, lu, cu = res[False]
            line = "%s layer %d: fused launches %s" % (name, li, cf)
            print(line)
            print("   y   : fused-fp64 %s | perlayer-fp64 %s | oracle32-fp64 %s" % (
                quantiles((yf.double() - y64).abs()), quantiles((yu.double() - y64).abs()),
                quantiles((y32.double() - y64).abs())))
            print("   lad : fused-fp64 %s | perlayer-fp64 %s | oracle32-fp64 %s" % (
                quantiles((lf.double() - l64).abs()), quantiles((lu.double() - l64).abs()),
                quantiles((l32.double() - l64).abs())))
            if err != 0 or not torch.isfinite(yf).all() or (yf.double() - y64).abs().max() > 1e-2:
                print("   !!! kernel error word %d, finite %s" % (err, bool(torch.isfinite(yf).all())))
                ok = False
            h = y32
    tensorcore.FUSED_CONDITIONER = True
    return ok


def timing(name, dev, steps=5):
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl)
    flow.load_state_dict(state)
    flow = flow.to(dev)
    B = min(wl["batch"], 1 << 20)
    x = torch.randn(B, wl["features"], generator=torch.Generator(device=dev).manual_seed(1234), device=dev)
    for fused in (True, False):
        tensorcore.FUSED_CONDITIONER = fused
        with torch.no_grad():
            for _ in range(2):
                lp = flow.log_prob(x)
            torch.cuda.synchronize()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(steps):
                lp = flow.log_prob(x)
            t1.record()
            torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        print("%s log_prob %d rows, fused=%s: %.2f ms/step, %.2f M samples/s, sum %.6e, error word %d" % (
            name, B, fused, ms, B / ms / 1e3, lp.double().sum().item(), conditioner.kernel_error()))
        if fused:
            prof = conditioner.kernel_profile()
            if prof["mma total"]:
                print("   profile (cycles, CTA 0, last launch):", prof)
    tensorcore.FUSED_CONDITIONER = True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1000)
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--only", default="cfg2")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    t = time.time()
    ok = True
    for name in args.only.split(","):
        ok = check(name, args.rows, dev) and ok
    print("check %s in %.1f s" % ("OK" if ok else "FAILED", time.time() - t))
    if args.time and ok:
        for name in args.only.split(","):
            timing(name, dev)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
