#!/usr/bin/env python
"""Secondary measurements: the BASELINE.json configs that are parity-test cases rather than the bench.py line.

    python scripts/bench_configs.py [--cpu]                 # 1 GPU
    torchrun --nproc-per-node N scripts/bench_configs.py    # N GPUs (cfg 3 training step with gradient all-reduce)

Prints one JSON line per measurement (rank 0).  Timing: CUDA events, 3 warm-up + 5 timed steps, max over ranks.
--cpu adds the oracle port (oracle/restated.py, torch CPU fp32, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import distributed as fdist, workloads  # noqa: E402


def timed(fn, dev, world, steps=5, warmup=3):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        fn()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def build(name, dev, trained_like=True):
    wl = workloads.get_workload(name)
    flow = workloads.build_flow(wl, seed=0)
    state = {k: v.clone() for k, v in flow.state_dict().items()}
    if trained_like:
        state = workloads.trained_like_(state, wl, seed=1)
    flow.load_state_dict(state)
    return wl, flow.to(dev), state


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", default="")
    ap.add_argument("--graph", action="store_true", help="also time the cfg-3 step replayed from a CUDA graph")
    ap.add_argument("--rows", type=int, default=0, help="override the cfg-3 global batch (e.g. 32768: one 8-GPU shard)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    only = set(args.only.split(",")) if args.only else None

    def emit(**kw):
        if rank == 0:
            kw.update(n_gpus=world, dtype="f32", data="synthetic")
            print(json.dumps(kw), flush=True)

    def want(tag):
        return only is None or tag in only

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    # ---- cfg 2: sample (inverse pass of the 8-layer coupling flow), 1M rows per GPU ---------------------------
    if want("cfg2_sample"):
        wl, flow, _ = build("cfg2", dev)
        B = wl["batch"]
        with torch.no_grad():
            ms = timed(lambda: flow.sample(B), dev, world)
        emit(metric="flow_sample_samples_per_sec", workload="cfg2 sample (randn + 8 inverse coupling layers), D=64 K=8 H=256",
             value=world * B / (ms * 1e-3), unit="samples/s", ms_per_step=ms, rows_per_gpu=B)
        del flow

    # ---- cfg 3: MAF-RQS training step, 262144 rows global ------------------------------------------------------
    if want("cfg3_train"):
        wl, flow, state = build("cfg3", dev, trained_like=False)
        fdist.broadcast_parameters(flow)
        Bg = args.rows or wl["batch"]
        B = Bg // world
        x = torch.randn(B, wl["features"], generator=gen, device=dev)
        opt = torch.optim.Adam(flow.parameters(), lr=1e-3, weight_decay=1e-5)

        def step():
            opt.zero_grad(set_to_none=True)
            loss = -flow.log_prob(x).mean()
            loss.backward()
            fdist.allreduce_gradients(flow)
            opt.step()

        ms = timed(step, dev, world)
        emit(metric="train_step_samples_per_sec",
             workload="cfg3 MAF-RQS D=16 K=16 5 layers H=256: zero_grad, -log_prob.mean, backward, flat-bucket gradient "
                      "all-reduce, Adam", value=Bg / (ms * 1e-3), unit="samples/s", ms_per_step=ms, rows_per_gpu=B,
             global_rows=Bg, scaling="strong")
        if args.graph:
            # the same step recorded into one CUDA graph (gradient all-reduce included): the launch-bound regime
            from flowconductor_b200 import graphs
            gopt = torch.optim.Adam(flow.parameters(), lr=1e-3, weight_decay=1e-5, capturable=True)
            gstep = graphs.GraphedTrainStep(flow, gopt, lambda xb: -flow.log_prob(xb).mean(), x,
                                            sync_gradients=fdist.allreduce_gradients)
            ms = timed(lambda: gstep.step(x), dev, world)
            emit(metric="train_step_samples_per_sec", impl="whole step replayed from one CUDA graph",
                 workload="cfg3 training step", value=Bg / (ms * 1e-3), unit="samples/s", ms_per_step=ms,
                 rows_per_gpu=B, global_rows=Bg, scaling="strong")
        if args.cpu and rank == 0:
            from oracle import restated
            specs = workloads.oracle_specs(wl)
            cs = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state.items()}
            xs = torch.randn(8192, wl["features"])
            torch.set_num_threads(os.cpu_count() or 1)
            t0 = time.perf_counter()
            loss = -restated.flow_log_prob(cs, specs, xs).mean()
            loss.backward()
            dt = time.perf_counter() - t0
            emit(metric="train_step_samples_per_sec", impl="cpu oracle port (forward + backward only, 8192 rows)",
                 workload="cfg3", value=8192 / dt, unit="samples/s", cores=os.cpu_count())
        del flow, opt

    # ---- cfg 4: conditional sum-of-sigmoids, 262144 rows per GPU ----------------------------------------------
    if want("cfg4_log_prob"):
        wl, flow, state = build("cfg4", dev)
        B = wl["batch"]
        x = torch.randn(B, wl["features"], generator=gen, device=dev)
        ctx = torch.randn(B, wl["context_features"], generator=gen, device=dev)
        with torch.no_grad():
            ms = timed(lambda: flow.log_prob(x, context=ctx), dev, world)
        emit(metric="flow_log_prob_samples_per_sec", workload="cfg4 conditional sum-of-sigmoids D=32 ctx=8 n=10 H=64 3 layers",
             value=world * B / (ms * 1e-3), unit="samples/s", ms_per_step=ms, rows_per_gpu=B)
        if args.cpu and rank == 0:
            from oracle import restated
            specs = workloads.oracle_specs(wl)
            torch.set_num_threads(os.cpu_count() or 1)
            xs, cs = torch.randn(16384, wl["features"]), torch.randn(16384, wl["context_features"])
            with torch.no_grad():
                t0 = time.perf_counter()
                restated.flow_log_prob(state, specs, xs, cs)
                dt = time.perf_counter() - t0
            emit(metric="flow_log_prob_samples_per_sec", impl="cpu oracle port (16384 rows)", workload="cfg4",
                 value=16384 / dt, unit="samples/s", cores=os.cpu_count())
        del flow

    # ---- cfg 5: D=256 coupling flow, streamed 1M-row chunks generated on device --------------------------------
    if want("cfg5_stream"):
        wl, flow, _ = build("cfg5", dev)
        chunk, chunks = wl["batch"], 4
        total = torch.zeros((), dtype=torch.float64, device=dev)

        def stream():
            for c in range(chunks):
                g = torch.Generator(device=dev).manual_seed(1234 + rank * 1000 + c)
                xc = torch.randn(chunk, wl["features"], generator=g, device=dev)
                with torch.no_grad():
                    total.add_(flow.log_prob(xc).double().sum())

        ms = timed(stream, dev, world, steps=2, warmup=1)
        emit(metric="flow_log_prob_samples_per_sec",
             workload="cfg5 RQ coupling D=256 K=8 8 layers H=256, {} x 1M-row chunks per GPU generated on device "
                      "(100M-row job = 96 such chunks per GPU at 1 GPU)".format(chunks),
             value=world * chunk * chunks / (ms * 1e-3), unit="samples/s", ms_per_step=ms, rows_per_gpu=chunk * chunks)
        del flow
    if world > 1:
        if args.graph:
            # a recorded graph keeps NCCL work objects alive and destroy_process_group() then waits forever: leave
            # through a barrier and a hard exit instead
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
