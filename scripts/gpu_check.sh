#!/bin/bash
# Run on the B200 box through gpurun: GPU parity tests, smoke, a short bench, and an ncu launch list.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit: $?" >> gpurun_out/pytest_gpu.log
grep -E '^(FAILED|ERROR|E  +Assertion)|passed|failed' gpurun_out/pytest_gpu.log | cut -c1-600 | tail -40
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit: $?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python scripts/bench_kernels.py ${SWEEP:-} > gpurun_out/kernels.log 2>&1; echo "kernels exit: $?"; grep -v "^sweep" gpurun_out/kernels.log | tail -20
