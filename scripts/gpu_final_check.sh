mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log | cut -c1-200
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').readline()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])"
