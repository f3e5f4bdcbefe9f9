#!/bin/bash
# GPU iteration on the tensor-core linear kernels: correctness + error profile, then bench; CTA-pair vs single-CTA
set -u
mkdir -p gpurun_out
for c in ${MODES:-3 1}; do
  echo "=== FC_LINEAR_MODE=$c"
  FC_LINEAR_MODE=$c timeout 180 python scripts/check_linear.py --bench > gpurun_out/check_linear_ctas$c.log 2>&1
  echo "exit $?"; cat gpurun_out/check_linear_ctas$c.log | tail -32
done
if [ "${PYTEST:-1}" = "1" ]; then
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit: $?"
grep -E '^(FAILED|ERROR|E  +Assertion)|passed|failed' gpurun_out/pytest_gpu.log | cut -c1-600 | tail -15
fi
