#!/usr/bin/env python
"""bench.py — Flow.log_prob samples/s of the RQ-spline coupling flow (BASELINE.json configs[1]: D=64, K=8 bins, 8 layers,
H=256, linear tails at 3.0, batch 1M synthetic Gaussian, fp32), plus the other two bench-size configs on request.

    python bench.py --gpus N --steps K --warmup W                       # this repo's CUDA path, cfg 2 (the metric's config)
    python bench.py --impl reference --steps K --warmup W               # the UNMODIFIED reference on the host cores
    python bench.py --workload cfg5 [--rows R]                          # D=256 flow, R (default 100M) rows streamed in 1M-row
                                                                        # chunks generated on device, sharded over the ranks
    python bench.py --workload cfg3_train [--graph]                     # MAF-RQS training step with gradient all-reduce

cfg 2: a step is one `flow.log_prob` pass over one batch resident in HBM (per GPU: weak scaling — every rank owns its own
1M-row shard, no data-path collective; one 2-element all-reduce of the log-likelihood sum per step).  Prints ONE JSON line
(contract in the task statement): value = whole-job samples/s with inputs resident; e2e = the same through the public
host-batch API (`distributed.host_log_prob`: pinned HOST buffers, chunked H2D overlapping the kernels + D2H of log_prob,
all inside the timed region); roofline = the fused conditioner + spline kernel against the measured dense fp16 tensor rate
(and roofline_elementwise = the stand-alone RQ-spline layer kernel against the measured HBM copy peak); cpu_baseline = the
unmodified reference (oracle/_ref) on the box's host cores on a bounded sample; gpu_eager_baseline = the unmodified
reference in eager mode on this GPU (the real "before" number).
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
# --impl reference: rows per step.  The reference's CPU path runs at ~1.5-3 k samples/s per 8 cores: the
# 1M-row step of our arm would take 5-10 minutes EACH, so the step is a bounded sample of the same workload (the
# per-sample rate is flat in the batch size: 33.0 k/s at 262144 rows vs 34.1 k/s at 65536).
REF_STEP_ROWS = {"cfg2": 65536, "cfg4": 16384, "cfg5": 16384, "cfg3_train": 8192}
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
FALLBACK_BF16_TFLOPS = 1400.0  # sustained dense bf16 (same guide)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=WORKLOAD, choices=["cfg2", "cfg4", "cfg5", "cfg3_train"])
    ap.add_argument("--batch", type=int, default=None, help="cfg2: rows per GPU (default: the workload's batch)")
    ap.add_argument("--rows", type=int, default=None, help="cfg5 / cfg3_train: GLOBAL rows per step (default 100M / 262144)")
    ap.add_argument("--graph", action="store_true", help="cfg3_train: replay the whole step from one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
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
    """Dense fp16 / bf16 tensor peak in TFLOP/s (the fused kernel issues kind::f16 UMMAs).  The kernel runs inside a long
    step -> the sustained figure."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return (float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    "measured (MEASURED_PEAKS.json bf16_tflops_sustained: dense 16-bit tensor rate under the power cap)")
        except Exception:
            pass
    return FALLBACK_BF16_TFLOPS, "fallback (B200_PROFILING.md sustained dense bf16)"


def profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return {}
    return {}


def build_state(wl, trained_like=True):
    from flowconductor_b200 import workloads

    flow = workloads.build_flow(wl, seed=0)
    state = {k: v.clone() for k, v in flow.state_dict().items()}
    if trained_like:
        state = workloads.trained_like_(state, wl, seed=1)
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
# the reference (test infrastructure: oracle/): unmodified package from oracle/_ref (or /root/reference) when it can be
# imported, else the restatement oracle/restated.py
# ------------------------------------------------------------------------------------------------
def reference_flow(wl, state, device="cpu"):
    """The workload built from the UNMODIFIED reference's own classes with our state_dict, or None."""
    try:
        from oracle import locate

        if not locate.have_reference():
            return None
        from oracle import make_golden

        flow = make_golden.build_reference_flow(wl)
        flow.load_state_dict(state)
        return flow.to(device).eval()
    except Exception as exc:  # noqa: BLE001  (the reference is optional test infrastructure)
        sys.stderr.write("reference import failed: {}\n".format(exc))
        return None


def cpu_log_prob_rate(wl, state, rows, repeats=1, warmup=0):
    """(samples/s, seconds per pass, kind) of the reference's log_prob on the host cores (all threads), fp32."""
    from flowconductor_b200 import workloads

    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(rows, wl["features"], generator=torch.Generator().manual_seed(1234))
    ctx = wl.get("context_features")
    c = torch.randn(rows, ctx, generator=torch.Generator().manual_seed(4321)) if ctx else None
    ref = reference_flow(wl, state)
    if ref is not None:
        kind = "reference"

        def run():
            ref.log_prob(x, context=c)
    else:
        from oracle import restated

        kind = "port"
        specs = workloads.oracle_specs(wl)

        def run():
            restated.flow_log_prob(state, specs, x, c)
    times = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            run()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    return rows * len(times) / sum(times), sum(times) / len(times), kind


def cpu_train_step_rate(wl, state, rows, repeats, warmup):
    """cfg 3 training step of the reference on the host cores: zero_grad, -log_prob.mean, backward, Adam."""
    torch.set_num_threads(os.cpu_count() or 1)
    ref = reference_flow(wl, state)
    kind = "reference"
    if ref is None:
        from flowconductor_b200 import workloads
        from oracle import restated

        kind = "port"
        specs = workloads.oracle_specs(wl)
        params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in state.items()}
        plist = [p for p in params.values() if p.requires_grad]

        def loss_fn(xb):
            return -restated.flow_log_prob(params, specs, xb).mean()
    else:
        ref.train()
        plist = list(ref.parameters())

        def loss_fn(xb):
            return -ref.log_prob(xb).mean()
    opt = torch.optim.Adam(plist, lr=1e-3, weight_decay=1e-5)
    x = torch.randn(rows, wl["features"], generator=torch.Generator().manual_seed(1234))
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss_fn(x).backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return rows * len(times) / sum(times), sum(times) / len(times), kind


def run_reference(args, wl):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  Rank 0 only; the
    other ranks of a torchrun launch exit without work — this is NOT an N-GPU run, whatever --gpus says."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    rows = REF_STEP_ROWS[args.workload]
    if args.workload == "cfg3_train":
        _, state = build_state(wl, trained_like=False)
        rate, sec, kind = cpu_train_step_rate(wl, state, rows, args.steps, args.warmup)
        metric = "train_step_samples_per_sec"
    else:
        _, state = build_state(wl)
        rate, sec, kind = cpu_log_prob_rate(wl, state, rows, repeats=args.steps, warmup=args.warmup)
        metric = METRIC
    cores = os.cpu_count() or 1
    sample = ("{} rows of {} per step ({} steps after {} warm-up), {} on torch CPU fp32, all {} host threads; bounded sample of "
              "the GPU arm's workload: the full-size step would take minutes each on the host".format(
                  rows, wl["name"], args.steps, args.warmup,
                  "the unmodified reference (oracle/_ref)" if kind == "reference" else "oracle/restated.py", cores))
    cfg = config_dict(args, wl, rows, 1)
    cfg["parallelism"] = "host CPU only, rank 0 (launched with --gpus {}: the GPUs are idle in this arm)".format(args.gpus)
    line = {
        "impl": "reference", "metric": metric, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, wl, rows_per_gpu, n_gpus):
    first = [l for l in wl["layers"] if l["kind"] != "permutation"][0]
    n_layers = len([l for l in wl["layers"] if l["kind"] != "permutation"])
    if args.workload == "cfg4":
        what = ("{}: ConditionalSumOfSigmoidsTransform flow log_prob (hypernetwork ResidualNet on the context), D={}, "
                "context={}, n_sigmoids={}, {} layers, H={}".format(wl["name"], wl["features"], wl["context_features"],
                                                                     first.get("n_sigmoids"), n_layers,
                                                                     first.get("hidden_features")))
        l2 = "a step touches x + context + outputs only (42 MB at 262144 rows): L2 is flushed between timed steps"
    elif args.workload == "cfg3_train":
        what = ("{}: MaskedPiecewiseRationalQuadraticAutoregressiveTransform (MAF-RQS) training step (zero_grad, "
                "-log_prob.mean, backward, gradient all-reduce, Adam), D={}, K={}, {} layers, H={}".format(
                    wl["name"], wl["features"], first.get("num_bins"), n_layers, first.get("hidden_features")))
        l2 = "activations of one step (GBs) exceed the 126 MB L2"
    else:
        what = ("{}: PiecewiseRationalQuadraticCouplingTransform flow log_prob, D={}, K={}, {} layers, H={}, "
                "tails=linear@{}".format(wl["name"], wl["features"], first.get("num_bins"), n_layers,
                                         first.get("hidden_features"), first.get("tail_bound")))
        l2 = "inputs_larger_than_l2 (x is {} MB per 1M rows vs 126 MB L2)".format(wl["features"] * 4)
    return {"workload": what, "rows_per_gpu": rows_per_gpu, "global_rows": rows_per_gpu * n_gpus,
            "parallelism": "row-sharded x{} (replicated weights, no data-path collective)".format(n_gpus),
            "weights": "random init seed 0" + ("" if args.workload == "cfg3_train"
                                               else " + trained-like perturbation seed 1 (SURVEY 8d)"),
            "l2": l2}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def setup_dist(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus {} needs torchrun (one process per GPU)".format(args.gpus))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local_rank, dev


def conditioner_flops(wl):
    """(algorithmic, executed-on-the-tensor-pipe) flops per row of ONE flow layer's conditioner (2 * (D_id H +
    2 nb H^2 + H D_t P); executed: 3 fp16 products per fp32 product on the padded shapes — the first layer reads the
    full-width row padded to 64, the final layer 24 columns per feature in 96-column tiles)."""
    first = [l for l in wl["layers"] if l["kind"] != "permutation"][0]
    D, H, nb, K = wl["features"], first["hidden_features"], first["num_blocks"], first["num_bins"]
    coupling = first["kind"] == "prq_coupling"
    d_t = D // 2 if coupling else D
    d_in = D - d_t if coupling else D
    p_per = 3 * K - 1
    ppad = {8: 24, 16: 48}[K]
    feats = 96 // ppad
    alg = 2.0 * (d_in * H + 2 * nb * H * H + H * d_t * p_per)
    k0_pad = (D + 63) // 64 * 64
    n_final = (d_t + feats - 1) // feats * 96
    executed = 3 * 2.0 * (k0_pad * H + 2 * nb * H * H + H * n_final)
    return alg, executed


def gpu_eager_baseline(wl, state, x, c=None):
    """The UNMODIFIED reference in eager mode on this GPU (fp32, TF32 off — torch's default), same weights, same rows."""
    ref = None
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        ref = reference_flow(wl, state, device=x.device)
        if ref is None:
            return {"unavailable": "reference package not importable (oracle/_ref missing)"}
        rows = x.shape[0]
        with torch.no_grad():
            while True:
                try:
                    ref.log_prob(x[:rows], context=None if c is None else c[:rows])
                    break
                except torch.cuda.OutOfMemoryError:
                    torch.cuda.empty_cache()
                    rows //= 2
                    if rows < 1024:
                        raise
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            reps = 3
            for _ in range(reps):
                ref.log_prob(x[:rows], context=None if c is None else c[:rows])
            e.record()
            torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        return {"value": rows / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "rows": rows, "kind": "reference",
                "how": "unmodified reference package (oracle/_ref) on cuda, eager mode, fp32, allow_tf32=False, same "
                       "state_dict, {} passes after 1 warm-up".format(reps)}
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": "{}: {}".format(type(exc).__name__, str(exc)[:200])}
    finally:
        del ref
        torch.cuda.empty_cache()


def run_cfg2(args, wl):
    import torch.distributed as dist

    from flowconductor_b200 import _cabi
    from flowconductor_b200 import distributed as fdist

    world, rank, local_rank, dev = setup_dist(args)
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
            total = lp.double().sum() if world == 1 else fdist.reduce_log_likelihood(lp)[0]
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
    counts = dict(_cabi.STATS.counts)
    n_c, c_ms = _cabi.STATS.elapsed_ms("fc_conditioner_rqs_apply")  # whole conditioner + spline, one launch per flow layer
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
        p_per = 3 * first["num_bins"] - 1
        # (1) the dominant kernel: the whole conditioner + spline of one flow layer (north_star item 3 widened to SURVEY
        # 8f n1): tensor bound.  Algorithmic flops per launch = rows * 2 (D_id H + 4 H^2 + H D_t P).
        tpeak, tpeak_src = measured_tensor_peak()
        alg_row, exec_row = conditioner_flops(wl)
        alg_flops, exec_flops = alg_row * B, exec_row * B
        c_each = (c_ms / n_c) if n_c else None
        roofline = {"bound": "tensor",
                    "kernel": "conditioner_f16x3_kernel<CondRqs<8>, 24, 2, false> (fc_conditioner_rqs_apply: initial layer, 2 residual "
                              "blocks, final layer and RQ spline of one flow layer in one persistent tcgen05 kernel)",
                    "achieved": (alg_flops / (c_each * 1e-3) / 1e12) if c_each else None, "peak": tpeak,
                    "unit": "TFLOP/s", "frac": (alg_flops / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None,
                    "traffic": profile_json("traffic_conditioner_rqs_apply.json").get("dram_bytes_per_launch"),
                    "peak_source": tpeak_src, "algorithmic_flops_per_launch": alg_flops,
                    "tensor_flops_executed_per_launch": exec_flops,
                    "tensor_pipe_frac": (exec_flops / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None,
                    "kernel_ms_per_launch": c_each, "kernel_share_of_step": (c_ms / ms) if n_c else None,
                    "launches_timed": n_c,
                    "note": "3xFP16 (fp32-faithful): 3 fp16 UMMAs per fp32 product on padded shapes, so frac <= "
                            "{:.3f}".format(alg_row / exec_row)}
        # (2) the stand-alone element-wise RQ-spline layer kernel (north_star items 1-2; the path taken whenever the
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
        # (3) the sampling direction of the same flow (BASELINE configs[1]: "log_prob + sample"): noise -> 8 inverse layers,
        # device-timed on this rank over the same number of rows; reported beside the headline, not part of `value`
        with torch.no_grad():
            flow._transform.inverse(x)
            torch.cuda.synchronize()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for _ in range(3):
                flow._transform.inverse(x)
            e3.record()
            torch.cuda.synchronize()
        ms_inv = s3.elapsed_time(e3) / 3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, B, world),
            "sample_direction": {"value": B / (ms_inv * 1e-3), "unit": "samples/s per GPU", "ms_per_pass": ms_inv,
                                 "what": "CompositeTransform.inverse of the same flow over the same rows (Flow.sample minus "
                                         "the base draw), rank 0"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "gpu_launches_by_entry_point": counts,
            "clocks": clocks,
            "roofline": roofline, "roofline_elementwise": roofline_hbm,
            "log_likelihood_sum": ll,
        }
        if world == 1 and not args.no_eager_baseline:
            line["gpu_eager_baseline"] = gpu_eager_baseline(wl, state, x)
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, kind = cpu_log_prob_rate(wl, state, CPU_SAMPLE_ROWS, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "{} rows of {} log_prob, one pass ({:.1f} s), {} on torch CPU fp32, all host "
                                              "threads".format(CPU_SAMPLE_ROWS, wl["name"], sec,
                                                               "the unmodified reference (oracle/_ref)"
                                                               if kind == "reference" else "oracle/restated.py")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args, wl):
    """BASELINE configs[3]: conditional sum-of-sigmoids flow with a hypernetwork conditioner, D = 32, context 8, 262 144 rows
    per GPU; one launch per flow layer (fc_conditioner_sos_apply).  Rows are sharded over the ranks (weak scaling)."""
    import torch.distributed as dist

    from flowconductor_b200 import _cabi
    from flowconductor_b200 import distributed as fdist

    world, rank, local_rank, dev = setup_dist(args)
    _cabi.lib()
    flow, state = build_state(wl)
    flow = flow.to(dev).eval()
    B = args.batch or wl["batch"]
    D, C = wl["features"], wl["context_features"]
    x = torch.randn(B, D, generator=torch.Generator(device=dev).manual_seed(1234 + rank), device=dev)
    c = torch.randn(B, C, generator=torch.Generator(device=dev).manual_seed(4321 + rank), device=dev)
    flush = torch.empty((1 << 26,), dtype=torch.float32, device=dev)  # 256 MB > the 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.no_grad():
            lp = flow.log_prob(x, context=c)
            total = lp.double().sum() if world == 1 else fdist.reduce_log_likelihood(lp)[0]
        return lp, total

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _cabi.STATS.reset()
    _cabi.STATS.timing = True
    ms = 0.0
    barrier()
    for _ in range(args.steps):  # every step timed on its own, the L2 flushed in between (the working set fits in L2)
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        lp, total = step()
        e.record()
        torch.cuda.synchronize()
        ms += s.elapsed_time(e)
    barrier()
    _cabi.STATS.timing = False
    launches = _cabi.STATS.total()
    counts = dict(_cabi.STATS.counts)
    n_c, c_ms = _cabi.STATS.elapsed_ms("fc_conditioner_sos_apply")
    clocks = sampler.stop() if rank == 0 else None
    ll = float(total.item())
    assert ll == ll, "log-likelihood is NaN"

    x_host, c_host = x.cpu().pin_memory(), c.cpu().pin_memory()
    out_host = torch.empty(B, dtype=torch.float32).pin_memory()
    for _ in range(min(2, args.warmup)):
        fdist.host_log_prob(flow, x_host, out_host, context_host=c_host)
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s2.record()
    for _ in range(args.steps):
        fdist.host_log_prob(flow, x_host, out_host, context_host=c_host)
    e2.record()
    barrier()
    t = torch.tensor([ms, s2.elapsed_time(e2)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        first = wl["layers"][0]
        H, nb, n = first["hidden_features"], first["num_blocks"], first["n_sigmoids"]
        P = 3 * n + 1
        alg_row = 2.0 * (C * H + 2 * nb * H * H + H * D * P)
        h_pad, k_real = (128 if H <= 128 else 256), (H + 63) // 64 * 64
        exec_row = 3 * 2.0 * (64 * h_pad + 2 * nb * k_real * h_pad + k_real * ((D + 1) // 2 * 96))
        tpeak, tpeak_src = measured_tensor_peak()
        c_each = (c_ms / n_c) if n_c else None
        roofline = {"bound": "tensor",
                    "kernel": "conditioner_f16x3_kernel<CondSos<10>, 48, 1, true> (fc_conditioner_sos_apply: ResidualNet on the "
                              "context + sum-of-sigmoids transform of one flow layer in one persistent tcgen05 kernel)",
                    "achieved": (alg_row * B / (c_each * 1e-3) / 1e12) if c_each else None, "peak": tpeak, "unit": "TFLOP/s",
                    "frac": (alg_row * B / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None, "traffic": None,
                    "peak_source": tpeak_src, "algorithmic_flops_per_launch": alg_row * B,
                    "tensor_flops_executed_per_launch": exec_row * B,
                    "tensor_pipe_frac": (exec_row * B / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None,
                    "kernel_ms_per_launch": c_each, "kernel_share_of_step": (c_ms / ms) if n_c else None, "launches_timed": n_c,
                    "note": "the only dense contraction of the step, hence the tensor roofline; the kernel itself is bound by the "
                            "sum-of-sigmoids arithmetic of its row threads (~800 instructions per element, DESIGN 4.12), not "
                            "by the tensor pipe"}
        line = {
            "metric": METRIC, "value": world * B * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, wl, B, world),
            "e2e": {"value": world * B * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * (D + C) * 4,
                    "d2h_bytes_per_step": B * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "gpu_launches_by_entry_point": counts, "clocks": clocks, "roofline": roofline,
            "log_likelihood_sum": ll,
        }
        # the sampling direction (numerical inverse, transforms/no_analytic_inv/base.py:23-83): conditioner outputs from one
        # fc_conditioner_store_apply per layer, then the safeguarded-Newton inverse kernel; outside the timed region above
        with torch.no_grad():
            flow._transform.inverse(x, context=c)
            torch.cuda.synchronize()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for _ in range(3):
                flow._transform.inverse(x, context=c)
            e3.record()
            torch.cuda.synchronize()
        ms_inv = s3.elapsed_time(e3) / 3
        line["sample_direction"] = {"value": B / (ms_inv * 1e-3), "unit": "samples/s per GPU", "ms_per_pass": ms_inv,
                                    "what": "CompositeTransform.inverse of the same flow over the same rows and contexts, rank 0"}
        if world == 1 and not args.no_eager_baseline:
            line["gpu_eager_baseline"] = gpu_eager_baseline(wl, state, x, c)
        if world == 1 and not args.no_cpu_baseline:
            rows = REF_STEP_ROWS["cfg4"] * 4
            rate, sec, kind = cpu_log_prob_rate(wl, state, rows, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "{} rows of cfg4 log_prob, one pass ({:.1f} s), {} on torch CPU fp32, all host "
                                              "threads".format(rows, sec, "the unmodified reference (oracle/_ref)"
                                                               if kind == "reference" else "oracle/restated.py")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg5(args, wl):
    """BASELINE configs[4]: batch-sharded log_prob of the D=256 flow over `--rows` (default 100M) rows in total, streamed in
    1M-row chunks generated on the device (seed 1234 + rank * 1000 + chunk), running fp64 sum, one 2-element
    all-reduce per step.  The rows are SHARDED over the ranks: strong scaling."""
    import torch.distributed as dist

    from flowconductor_b200 import _cabi
    from flowconductor_b200 import distributed as fdist

    world, rank, local_rank, dev = setup_dist(args)
    _cabi.lib()
    flow, state = build_state(wl)
    flow = flow.to(dev).eval()
    D, chunk = wl["features"], wl["batch"]
    total_rows = args.rows or wl["total_rows"]
    lo, hi = fdist.shard_bounds(total_rows, rank, world)
    my_rows = hi - lo
    bounds = [(c0, min(my_rows, c0 + chunk)) for c0 in range(0, my_rows, chunk)]
    xbuf = torch.empty(chunk, D, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        acc = torch.zeros((), dtype=torch.float64, device=dev)
        with torch.no_grad():
            for ci, (c0, c1) in enumerate(bounds):
                g = torch.Generator(device=dev).manual_seed(1234 + rank * 1000 + ci)
                xc = xbuf[: c1 - c0]
                xc.normal_(generator=g)
                acc += flow.log_prob(xc).double().sum()
            packed = torch.stack((acc, torch.tensor(float(my_rows), dtype=torch.float64, device=dev)))
            if world > 1:
                dist.all_reduce(packed)
        return packed

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _cabi.STATS.reset()
    _cabi.STATS.timing = True
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        packed = step()
    end.record()
    barrier()
    _cabi.STATS.timing = False
    ms = start.elapsed_time(end)
    launches = _cabi.STATS.total()
    n_c, c_ms = _cabi.STATS.elapsed_ms("fc_conditioner_rqs_apply")
    clocks = sampler.stop() if rank == 0 else None
    ll, count = packed.tolist()
    assert ll == ll and int(count) == total_rows, (ll, count, total_rows)

    # e2e: the same rows from pinned HOST memory (a 4M-row host buffer re-read until the shard is covered), chunked H2D
    # overlapping the kernels, D2H of every chunk's log_prob
    host_rows = min(my_rows, 4 * chunk)
    x_host = torch.randn(host_rows, D).pin_memory()
    out_host = torch.empty(host_rows, dtype=torch.float32).pin_memory()
    passes = (my_rows + host_rows - 1) // host_rows

    def e2e_step():
        for _ in range(passes):
            fdist.host_log_prob(flow, x_host, out_host, chunk_rows=262144)

    e2e_step()
    barrier()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(1, min(args.steps, 2))
    s2.record()
    for _ in range(e2e_steps):
        e2e_step()
    e2.record()
    barrier()
    ms_e2e = s2.elapsed_time(e2)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        tpeak, tpeak_src = measured_tensor_peak()
        alg_row, exec_row = conditioner_flops(wl)
        c_each = (c_ms / n_c) if n_c else None
        rows_per_launch = my_rows / max(len(bounds), 1)
        line = {
            "metric": METRIC, "value": total_rows * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args, wl, my_rows, 1), global_rows=total_rows, chunk_rows=chunk,
                           parallelism="{} rows sharded over {} rank(s), streamed in {}-row chunks generated on device; "
                                       "no data-path collective".format(total_rows, world, chunk)),
            "e2e": {"value": passes * host_rows * world * e2e_steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": passes * host_rows * D * 4, "d2h_bytes_per_step": passes * host_rows * 4,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "conditioner_f16x3_kernel<CondRqs<8>, 24, 2, false> (fc_conditioner_rqs_apply)",
                         "achieved": (alg_row * rows_per_launch / (c_each * 1e-3) / 1e12) if c_each else None,
                         "peak": tpeak, "unit": "TFLOP/s",
                         "frac": (alg_row * rows_per_launch / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None,
                         "traffic": None, "peak_source": tpeak_src,
                         "tensor_pipe_frac": (exec_row * rows_per_launch / (c_each * 1e-3) / 1e12 / tpeak) if c_each else None,
                         "kernel_ms_per_launch": c_each, "kernel_share_of_step": (c_ms / ms) if n_c else None,
                         "launches_timed": n_c},
            "log_likelihood_sum": ll,
        }
        if world == 1 and not args.no_cpu_baseline:
            rows = REF_STEP_ROWS["cfg5"]
            rate, sec, kind = cpu_log_prob_rate(wl, state, rows, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "{} rows of cfg5 log_prob, one pass ({:.1f} s), torch CPU fp32".format(rows, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg3_train(args, wl):
    """BASELINE configs[2]: MAF-RQS training step — zero_grad; loss = -log_prob(x).mean(); backward; flat-bucket gradient
    all-reduce (NCCL); Adam(lr=1e-3, weight_decay=1e-5) — as in examples/toy_2d.py:57-67.  The 262144-row global batch is
    SHARDED over the ranks: strong scaling."""
    import torch.distributed as dist

    from flowconductor_b200 import _cabi, graphs
    from flowconductor_b200 import distributed as fdist

    world, rank, local_rank, dev = setup_dist(args)
    _cabi.lib()
    flow, state = build_state(wl, trained_like=False)
    flow = flow.to(dev).train()
    fdist.broadcast_parameters(flow)
    Bg = args.rows or wl["batch"]
    B = Bg // world
    x = torch.randn(B, wl["features"], generator=torch.Generator(device=dev).manual_seed(1234 + rank), device=dev)
    opt = torch.optim.Adam(flow.parameters(), lr=1e-3, weight_decay=1e-5, capturable=args.graph)

    def eager_step():
        opt.zero_grad(set_to_none=True)
        loss = -flow.log_prob(x).mean()
        loss.backward()
        fdist.allreduce_gradients(flow)
        opt.step()
        return loss

    if args.graph:
        gstep = graphs.GraphedTrainStep(flow, opt, lambda xb: -flow.log_prob(xb).mean(), x,
                                        sync_gradients=fdist.allreduce_gradients)

        def step():
            return gstep.step(x)
    else:
        step = eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    _cabi.STATS.reset()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for _ in range(args.steps):
        loss = step()
    end.record()
    barrier()
    ms = start.elapsed_time(end)
    launches = _cabi.STATS.total()
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the step fed from pinned host memory, loss read back on the host
    x_host = x.cpu().pin_memory()
    s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s2.record()
    for _ in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        float(step().item())
    e2.record()
    barrier()
    ms_e2e = s2.elapsed_time(e2)
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    if rank == 0:
        line = {
            "metric": "train_step_samples_per_sec", "value": Bg * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_dict(args, wl, B, 1), global_rows=Bg,
                           parallelism="data-parallel x{}: global batch sharded, one flat-bucket NCCL all-reduce of all "
                                       "gradients per step{}".format(world, "; whole step replayed from one CUDA graph"
                                                                     if args.graph else "")),
            "e2e": {"value": Bg * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * wl["features"] * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches if not args.graph else None, "clocks": clocks, "loss": float(loss.item()),
            "roofline": None,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, sec, kind = cpu_train_step_rate(wl, state, REF_STEP_ROWS["cfg3_train"], 1, 0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": "{} rows, one training step ({:.1f} s), torch CPU fp32".format(
                                        REF_STEP_ROWS["cfg3_train"], sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        if args.graph:
            # a recorded graph keeps NCCL work objects alive and destroy_process_group() then waits forever: leave through
            # a barrier and a hard exit instead
            dist.barrier()
            torch.cuda.synchronize()
            sys.stdout.flush()
            os._exit(0)
        dist.destroy_process_group()


def main():
    args = parse()
    from flowconductor_b200 import workloads

    wl = workloads.get_workload("cfg3" if args.workload == "cfg3_train" else args.workload)
    if args.impl == "reference":
        run_reference(args, wl)
    elif args.workload == "cfg5":
        run_cfg5(args, wl)
    elif args.workload == "cfg3_train":
        run_cfg3_train(args, wl)
    elif args.workload == "cfg4":
        run_cfg4(args, wl)
    else:
        run_cfg2(args, wl)


if __name__ == "__main__":
    main()
