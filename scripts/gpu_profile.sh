#!/bin/bash
# Round profile run (under gpurun): full bench line, ncu launch list of the same command, and one full capture of
# the dominant hand-written kernel.  Outputs in gpurun_out/; summaries are copied into profiles/ by hand.
set -u
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"
cat gpurun_out/bench_full.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list exit $?"
python scripts/bench_kernels.py > gpurun_out/kernels.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pipelined_apply -s 3 -c 1 -f -o gpurun_out/rqs_pipe \
    python scripts/bench_kernels.py > gpurun_out/ncu_run.log 2>&1
echo "ncu full exit $?"
grep -v staged gpurun_out/kernels.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>&1; cat gpurun_out/bench_reference.json | cut -c1-400
