"""Summarise an .ncu-rep (read here with `ncu -i`) into the small CSV / JSON files kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/linear_tc.ncu-rep            # key metrics of every captured launch
    python scripts/ncu_summary.py --launches gpurun_out/launches_tc.csv   # launch list -> share table
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("# kernel: {}".format(d.get("Kernel Name", "?")))
        print("metric,unit,value")
        for h in hdr:
            if any(h == k or h.endswith("." + k) or h.endswith(k) for k in KEYS):
                print("{},{},{}".format(h, u[h], d[h]))
        print()


def launches(path):
    agg = defaultdict(lambda: [0, 0.0])
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        k = r["Kernel Name"]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("share_pct,launches,avg_ns,total_ns,kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('{:.2f},{},{:.0f},{:.0f},"{}"'.format(100 * t / tot, n, t / n, t, k[:140]))


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
