#!/usr/bin/env python
"""Launch-bound regime: eager calls vs a replayed CUDA graph (flowconductor_b200.graphs), SURVEY §8(f) n2 / n4.

    python scripts/bench_graphs.py

One JSON line per case: median wall time per call (host clock around a synchronised call, because the quantity of
interest is launch latency as the caller sees it), device time of the replay from CUDA events, and the number of
library kernel launches one eager call issues.
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import _cabi, graphs, workloads  # noqa: E402


def wall_ms(fn, reps=30, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def device_ms(fn, reps=30):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def main():
    dev = torch.device("cuda:0")
    cases = [("cfg1", 500, "log_prob"), ("cfg1", 10000, "log_prob"), ("cfg1", 10000, "inverse"),
             ("cfg3", 4096, "log_prob"), ("cfg3", 4096, "inverse"), ("cfg3", 32768, "inverse"),
             ("cfg2", 4096, "log_prob"), ("cfg2", 4096, "inverse"), ("cfg4", 4096, "log_prob")]
    for name, rows, what in cases:
        wl = workloads.get_workload(name)
        flow = workloads.build_flow(wl, seed=0)
        state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl, seed=1)
        flow.load_state_dict(state)
        flow = flow.to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(7)
        x = torch.randn(rows, wl["features"], generator=g, device=dev)
        ctx = wl.get("context_features")
        c = torch.randn(rows, ctx, generator=g, device=dev) if ctx else None
        if what == "log_prob":
            fn = (lambda a, b: flow.log_prob(a, context=b)) if ctx else flow.log_prob
        else:
            fn = (lambda a, b: flow._transform.inverse(a, context=b)) if ctx else flow._transform.inverse
        args = (x, c) if ctx else (x,)
        with torch.no_grad():
            fn(*args)
            _cabi.STATS.reset()
            fn(*args)
            launches = _cabi.STATS.total()
            eager = wall_ms(lambda: fn(*args))
            eager_dev = device_ms(lambda: fn(*args))
            gc = graphs.capture(fn, *args)
            same = all(torch.equal(a, b) for a, b in zip(graphs._flatten(gc(*args)), graphs._flatten(fn(*args))))
            graphed = wall_ms(lambda: gc(*args))
            graphed_dev = device_ms(lambda: gc(*args))
        print(json.dumps({"workload": name, "rows": rows, "call": what, "library_launches_per_call": launches,
                          "eager_ms": round(eager, 4), "graph_ms": round(graphed, 4),
                          "speedup": round(eager / graphed, 2), "eager_device_ms": round(eager_dev, 4),
                          "graph_device_ms": round(graphed_dev, 4), "bitwise_equal": bool(same),
                          "rows_per_s_graph": round(rows / graphed * 1e3)}), flush=True)


if __name__ == "__main__":
    main()
