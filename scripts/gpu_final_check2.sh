mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-200
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
for w in cfg4 cfg5 cfg3_train; do timeout 900 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"; done
python - <<'PY'
import json
for f in ("bench_final","bench_cfg4","bench_cfg5","bench_cfg3_train"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).readline()); print(f, round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), (d.get('roofline') or {}).get('frac'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e: print(f, "ERR", e)
PY
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernel_microbench.jsonl 2> gpurun_out/kernel_microbench.err; echo "microbench rc=$?"
timeout 600 python bench.py --workload cfg3_train --graph > gpurun_out/bench_cfg3_train_graph.json 2> gpurun_out/bench_cfg3_train_graph.err; echo "cfg3 graph rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_train_graph.json').readline()); print('graph', d['value'], d['ms_per_step'])"
