#!/usr/bin/env python
"""Autoregressive inverse / sampling: the incremental-MADE kernel (csrc/fc_made_inverse.cu) against the
D-pass inverse (eager and replayed from a CUDA graph) on cfg 3 (MAF-RQS D=16, K=16, 5 layers, H=256).

    python scripts/bench_made_inverse.py [rows ...]

One JSON line per batch size: device time per `flow._transform.inverse` call (CUDA events around 20 calls).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import _cabi, graphs, made_inverse, workloads  # noqa: E402


def device_ms(fn, reps=20, warmup=3):
    for _ in range(warmup):
        fn()
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
    rows_list = [int(a) for a in sys.argv[1:]] or [4096, 32768, 262144]
    wl = workloads.get_workload("cfg3")
    flow = workloads.build_flow(wl, seed=0)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl, seed=1)
    flow.load_state_dict(state)
    flow = flow.to(dev).eval()
    inv = flow._transform.inverse
    for rows in rows_list:
        g = torch.Generator(device=dev).manual_seed(7)
        z = torch.randn(rows, wl["features"], generator=g, device=dev)
        with torch.no_grad():
            made_inverse.ENABLED = True
            x_new, lad_new = inv(z)
            _cabi.STATS.reset()
            inv(z)
            launches_new = _cabi.STATS.total()
            t_new = device_ms(lambda: inv(z))
            gnew = graphs.capture(inv, z)
            t_new_graph = device_ms(lambda: gnew(z))
            made_inverse.ENABLED = False
            x_old, lad_old = inv(z)
            _cabi.STATS.reset()
            inv(z)
            launches_old = _cabi.STATS.total()
            t_old = device_ms(lambda: inv(z), reps=5, warmup=1)
            gold = graphs.capture(inv, z)
            t_old_graph = device_ms(lambda: gold(z), reps=5, warmup=1)
            made_inverse.ENABLED = True
        print(json.dumps({"workload": "cfg3 inverse (sampling direction)", "rows": rows,
                          "incremental_ms": round(t_new, 4), "incremental_graph_ms": round(t_new_graph, 4),
                          "d_pass_ms": round(t_old, 4), "d_pass_graph_ms": round(t_old_graph, 4),
                          "speedup_vs_d_pass_graph": round(t_old_graph / min(t_new, t_new_graph), 2),
                          "samples_per_s": round(rows / (min(t_new, t_new_graph) * 1e-3)),
                          "library_launches": {"incremental": launches_new, "d_pass": launches_old},
                          "max_abs_diff_outputs": float((x_new - x_old).abs().max()),
                          "max_abs_diff_logabsdet": float((lad_new - lad_old).abs().max())}), flush=True)


if __name__ == "__main__":
    main()
