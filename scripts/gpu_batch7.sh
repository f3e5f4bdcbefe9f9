python scripts/bench_made_inverse.py 4096 32768 2>/dev/null | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_conditioner.py -m gpu -q --timeout 300 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('cfg2 ms_per_step', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'clk', d['clocks']['sm_mhz'])"
python scripts/bench_configs.py --only cfg4_log_prob 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('cfg4 ms', d['ms_per_step'])"
