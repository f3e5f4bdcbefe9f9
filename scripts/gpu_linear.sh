#!/bin/bash
# GPU iteration on the tensor-core linear kernels: correctness + error profile per accumulation chunk, then bench
set -u
mkdir -p gpurun_out
for ck in ${CHUNKS:-32}; do
  echo "=== FC_LINEAR_CHUNK_K=$ck"
  FC_LINEAR_CHUNK_K=$ck timeout 300 python scripts/check_linear.py --bench > gpurun_out/check_linear_c$ck.log 2>&1
  echo "exit $?"; cat gpurun_out/check_linear_c$ck.log
done
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit: $?"
grep -E '^(FAILED|ERROR|E  +Assertion)|passed|failed' gpurun_out/pytest_gpu.log | cut -c1-600 | tail -15
