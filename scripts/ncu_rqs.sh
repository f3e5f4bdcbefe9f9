#!/bin/bash
# ncu capture of the RQ-spline layer kernel (run under gpurun; ONE ncu use per call).
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_kernels.py --B 1048576"
$CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pipelined_apply -s 3 -c 1 -f -o gpurun_out/rqs_pipe $CMD > gpurun_out/ncu_run.log 2>&1
echo "ncu exit: $?"; tail -3 gpurun_out/ncu_run.log; ls -la gpurun_out/*.ncu-rep
