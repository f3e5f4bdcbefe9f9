for rep in 1 2; do
for v in a cur; do
  if [ $v = a ]; then export FC_LIB=/root/repo/flowconductor_b200/lib/variants/libflowcon_b200_a.so; else unset FC_LIB; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$v cfg2 ms_per_step', d['ms_per_step'], 'kernel ms', d['roofline']['kernel_ms_per_launch'], 'clk', d['clocks']['sm_mhz'])"
done; done
unset FC_LIB
python scripts/bench_configs.py --only cfg4_log_prob 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('cfg4 ms', d['ms_per_step'])"
