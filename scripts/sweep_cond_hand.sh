mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioner.py -m gpu -q --timeout 300 > gpurun_out/cond_tests.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED" gpurun_out/cond_tests.log | tail -5
for h in 0 1 2 3 4; do
  echo "== FC_COND_HAND=$h"
  FC_COND_HAND=$h timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('cfg2 ms_per_step', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'clk', d['clocks']['sm_mhz'])"
  FC_COND_HAND=$h timeout 300 python scripts/bench_configs.py --only cfg4_log_prob 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('cfg4 ms', d['ms_per_step'])"
done
