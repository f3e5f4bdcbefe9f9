#!/bin/bash
# quick GPU iteration: parity tests + accuracy + kernel microbench (+ optional ncu of the RQS kernel)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit: $?"
grep -E '^(FAILED|ERROR|E  +Assertion)|passed|failed' gpurun_out/pytest_gpu.log | cut -c1-500 | tail -15
python scripts/accuracy_report.py > gpurun_out/accuracy.log 2>&1; cat gpurun_out/accuracy.log
python scripts/bench_kernels.py ${SWEEP:-} > gpurun_out/kernels.log 2>&1; echo "kernels exit: $?"; grep -v '^sweep' gpurun_out/kernels.log | cut -c1-250; grep '^sweep' gpurun_out/kernels.log | sort -k8 -n -t' ' | sort -k7 -g | head -8
if [ "${NCU:-0}" = "1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:pipelined_apply -s 3 -c 1 -f -o gpurun_out/rqs_pipe python scripts/bench_kernels.py > gpurun_out/ncu_run.log 2>&1
  echo "ncu exit: $?"
fi
