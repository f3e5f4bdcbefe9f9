#!/usr/bin/env python
"""bench.py — Flow.log_prob samples/s of the RQ-spline coupling flow (BASELINE.json configs[1]:
D=64, K=8 bins, 8 layers, H=256, linear tails at 3.0, batch 1M synthetic Gaussian, fp32).

    python bench.py --gpus N --steps K --warmup W              # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W      # CPU port of the reference (oracle), host cores

A step is one `flow.log_prob` pass over one batch resident in HBM (per GPU: weak scaling — every rank owns
its own 1M-row shard, no data-path collective; one 2-element all-reduce of the log-likelihood sum per step).
Prints ONE JSON line (contract in the task statement): value = whole-job samples/s with inputs resident,
e2e = same through the public host-batch API (`distributed.host_log_prob`: pinned HOST buffers, chunked H2D of the
batch overlapping the kernels + D2H of log_prob, all inside the timed region), roofline = the RQ-spline layer kernel against the measured HBM copy peak, cpu_baseline = the
oracle port on the box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "flow_log_prob_samples_per_sec"
UNIT = "samples/s"
WORKLOAD = "cfg2"
CPU_SAMPLE_ROWS = 262144      # cpu_baseline leg of the default run (~10-30 s of host work)
REF_STEP_ROWS = 65536         # --impl reference: rows per step (bounded so K+W steps end within minutes)
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
FALLBACK_BF16_TFLOPS = 2250.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD)
    ap.add_argument("--batch", type=int, default=None, help="rows per GPU (default: the workload's batch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def measured_tensor_peak():
    """Dense tf32 tensor peak in TFLOP/s = half the measured dense bf16 rate (kind::tf32 UMMAs retire 8 k-values
    per 128xN instruction where kind::f16 retires 16).  The kernels run inside a long step -> sustained figure."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return (float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) / 2,
                    "measured (MEASURED_PEAKS.json bf16_tflops_sustained / 2 = dense tf32)")
        except Exception:
            pass
    return FALLBACK_BF16_TFLOPS / 2, "fallback (B200_PROFILING.md nominal dense bf16 2250 / 2 = dense tf32)"


def profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return {}
    return {}


def build_state(wl):
    from flowconductor_b200 import workloads

    flow = workloads.build_flow(wl, seed=0)
    state = workloads.trained_like_({k: v.clone() for k, v in flow.state_dict().items()}, wl, seed=1)
    flow.load_state_dict(state)
    return flow, state


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU port of the reference (oracle) — used for cpu_baseline and for --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_log_prob_rate(wl, state, rows, repeats=1, warmup=0):
    """samples/s of oracle.restated.flow_log_prob on the host cores (all threads), fp32."""
    from flowconductor_b200 import workloads
    from oracle import restated

    torch.set_num_threads(os.cpu_count() or 1)
    specs = workloads.oracle_specs(wl)
    x = torch.randn(rows, wl["features"], generator=torch.Generator().manual_seed(1234))
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            restated.flow_log_prob(state, specs, x)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return rows * len(times) / sum(times), sum(times) / len(times)


def run_reference(args, wl):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference is pure
    Python/PyTorch, so there is no compiled oracle/_ref).  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    _, state = build_state(wl)
    rate, sec = cpu_log_prob_rate(wl, state, REF_STEP_ROWS, repeats=args.steps, warmup=args.warmup)
    cores = os.cpu_count() or 1
    sample = "{} rows of {} per step, {} steps after {} warm-up, torch CPU fp32".format(
        REF_STEP_ROWS, wl["name"], args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(wl, REF_STEP_ROWS, args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(wl, rows_per_gpu, n_gpus):
    first = wl["layers"][0]
    return {"workload": "{}: PiecewiseRationalQuadraticCouplingTransform flow log_prob, D={}, K={}, {} layers, H={}, "
                        "tails=linear@{}".format(wl["name"], wl["features"], first.get("num_bins"),
                                                 len(wl["layers"]), first.get("hidden_features"),
                                                 first.get("tail_bound")),
            "rows_per_gpu": rows_per_gpu, "global_rows": rows_per_gpu * n_gpus,
            "parallelism": "row-sharded x{} (replicated weights, no data-path collective)".format(n_gpus),
            "weights": "random init seed 0 + trained-like perturbation seed 1 (SURVEY 8d)",
            "l2": "inputs_larger_than_l2 (x 268 MB + 3 GB of spline parameters per layer vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl):
    import torch.distributed as dist

    from flowconductor_b200 import _cabi
    from flowconductor_b200 import distributed as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus {} needs torchrun (one process per GPU)".format(args.gpus))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    flow, state = build_state(wl)
    flow = flow.to(dev).eval()
    B = args.batch or wl["batch"]
    D = wl["features"]
    x = torch.randn(B, D, generator=torch.Generator(device=dev).manual_seed(1234 + rank), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(inp):
        with torch.no_grad():
            lp = flow.log_prob(inp)
            total, count = (lp.double().sum(), lp.numel()) if world == 1 else fdist.reduce_log_likelihood(lp)
        return lp, total

    for _ in range(args.warmup):
        step(x)
    barrier()

    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _cabi.STATS.reset()
    _cabi.STATS.timing = True
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        lp, total = step(x)
    end.record()
    barrier()
    _cabi.STATS.timing = False
    ms = start.elapsed_time(end)
    launches = _cabi.STATS.total()
    n_f, f_ms = _cabi.STATS.elapsed_ms("fc_linear_rqs_apply")   # final conditioner layer + spline (tensor cores)
    n_h, h_ms = _cabi.STATS.elapsed_ms("fc_linear_apply")       # other conditioner layers (tensor cores)
    clocks = sampler.stop() if rank == 0 else None
    ll = float(total.item())
    assert ll == ll, "log-likelihood is NaN"

    # ---- timed region 2: end to end through the public API with host buffers ----------------------
    x_host = x.cpu().pin_memory()
    out_host = torch.empty(B, dtype=torch.float32).pin_memory()

    def e2e_step():
        # public host-batch entry point: chunked, double-buffered H2D overlapping the kernels, D2H of every chunk's result
        fdist.host_log_prob(flow, x_host, out_host)

    for _ in range(min(2, args.warmup)):
        e2e_step()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(args.steps):
        e2e_step()
    e2.record()
    barrier()
    ms_e2e = s2.elapsed_time(e2)

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    if rank == 0:
        value = world * B * args.steps / (ms * 1e-3)
        e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
        first = wl["layers"][0]
        d_t = D // 2
        H = first["hidden_features"]
        p_per = 3 * first["num_bins"] - 1
        # (1) the fused final-layer kernel (north_star item 3): tensor bound.  Algorithmic flops per launch =
        # 2 * rows * H * (D_t * P); the kernel executes 3 tf32 UMMAs per product on N padded to 24 columns
        # per feature, so the tensor pipe does 3 * 2 * rows * H * (D_t * 24).
        tpeak, tpeak_src = measured_tensor_peak()
        alg_flops = 2.0 * B * H * d_t * p_per
        exec_flops = 3 * 2.0 * B * H * d_t * 24
        f_each = (f_ms / n_f) if n_f else None
        roofline = {"bound": "tensor", "kernel": "linear_tf32x3_kernel<EPI=rqs> (fc_linear_rqs_apply: final conditioner "
                                                 "GEMM + RQ spline epilogue)",
                    "achieved": (alg_flops / (f_each * 1e-3) / 1e12) if f_each else None, "peak": tpeak,
                    "unit": "TFLOP/s", "frac": (alg_flops / (f_each * 1e-3) / 1e12 / tpeak) if f_each else None,
                    "traffic": profile_json("traffic_linear_rqs_apply.json").get("dram_bytes_per_launch"),
                    "peak_source": tpeak_src, "algorithmic_flops_per_launch": alg_flops,
                    "tensor_flops_executed_per_launch": exec_flops,
                    "tensor_pipe_frac": (exec_flops / (f_each * 1e-3) / 1e12 / tpeak) if f_each else None,
                    "kernel_ms_per_launch": f_each, "kernel_share_of_step": (f_ms / ms) if n_f else None,
                    "launches_timed": n_f,
                    "note": "3xTF32 (fp32-faithful): 3 tf32 UMMAs per fp32 product, so frac <= 1/3 * 23/24"}
        # (2) the other conditioner layers (same kernel, store epilogue): 4 x (H x H) + 1 x (D x H) per flow layer
        n_layers = len(wl["layers"])
        hid_alg = 2.0 * B * (4 * H * H + d_t * H) * n_layers * args.steps
        hid_exec = 3 * 2.0 * B * (4 * H * H + D * H) * n_layers * args.steps  # first layer reads the full-width rows
        roofline_hidden = {"bound": "tensor", "kernel": "linear_tf32x3_kernel<EPI=store> (fc_linear_apply)",
                           "achieved": (hid_alg / (h_ms * 1e-3) / 1e12) if n_h else None, "peak": tpeak,
                           "unit": "TFLOP/s", "frac": (hid_alg / (h_ms * 1e-3) / 1e12 / tpeak) if n_h else None,
                           "tensor_pipe_frac": (hid_exec / (h_ms * 1e-3) / 1e12 / tpeak) if n_h else None,
                           "kernel_share_of_step": (h_ms / ms) if n_h else None, "launches_timed": n_h}
        # (3) the stand-alone element-wise RQ-spline layer kernel (north_star items 1-2; the path taken whenever the
        # parameters are materialised: training, unsupported conditioners): HBM bound.  Not on the inference step
        # above, so it is timed here on its own: 10 launches over a materialised [B, D_t * P] parameter tensor.
        # algorithmic bytes per sample per launch: x + params + y + lad + identity copy
        bytes_per_sample = 4 * d_t + 4 * d_t * p_per + 4 * d_t + 4 + 8 * (D - d_t)
        peak, peak_src = measured_peak()
        layer0 = flow._transform._transforms[0]
        prm = torch.randn(B, d_t * p_per, device=dev)
        _cabi.STATS.reset()
        with torch.no_grad():
            for _ in range(3):
                layer0._coupling_layer(x, prm, False)
            torch.cuda.synchronize()
            _cabi.STATS.reset()
            _cabi.STATS.timing = True
            for _ in range(10):
                layer0._coupling_layer(x, prm, False)
            torch.cuda.synchronize()
            _cabi.STATS.timing = False
        n_k, k_ms = _cabi.STATS.elapsed_ms("fc_rqs_apply")
        del prm
        achieved = bytes_per_sample * B / ((k_ms / max(n_k, 1)) * 1e-3) / 1e9 if n_k else None
        roofline_hbm = {"bound": "hbm", "kernel": "pipelined_apply_kernel<RqsOp<8>, true> (fc_rqs_apply), timed alone",
                        "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": (achieved / peak) if achieved else None,
                        "traffic": profile_json("traffic_rqs_apply.json").get("dram_bytes_per_launch"),
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_sample * B,
                        "kernel_ms_per_launch": (k_ms / n_k) if n_k else None, "launches_timed": n_k}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, B, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline, "roofline_hidden_layers": roofline_hidden, "roofline_elementwise": roofline_hbm,
            "log_likelihood_sum": ll,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, sec = cpu_log_prob_rate(wl, state, CPU_SAMPLE_ROWS, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "{} rows of {} log_prob, one pass ({:.1f} s), oracle/restated.py on "
                                              "torch CPU fp32, all host threads".format(CPU_SAMPLE_ROWS, wl["name"],
                                                                                        sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    from flowconductor_b200 import workloads

    wl = workloads.get_workload(args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
