#!/bin/bash
# Round profile run of the tensor-core inference path (under gpurun): full bench line, ncu launch list of the same
# command, one full capture each of the hidden-layer GEMM and the fused final-layer GEMM + spline kernel.
set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_tc.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list exit $?"
python scripts/prof_linear.py > gpurun_out/prof_linear.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:linear_tf32x3 -s 2 -c 2 -f -o gpurun_out/linear_tc \
    python scripts/prof_linear.py > gpurun_out/ncu_linear.log 2>&1
echo "ncu full exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cut -c1-300 gpurun_out/bench_reference.json
