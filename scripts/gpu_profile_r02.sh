#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): full default bench line, the reference arm, ncu launch list of the bench
# command, one full capture of the fused conditioner kernel, the other configs, the kernel microbenchmarks.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>&1; echo "reference exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/bench_short.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_cfg2_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu launch list exit $?"
python scripts/prof_conditioner.py > gpurun_out/prof_cond.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conditioner_f16x3 -s 2 -c 1 -f -o gpurun_out/r02_conditioner_final \
    python scripts/prof_conditioner.py > gpurun_out/ncu_cond.log 2>&1
echo "ncu full exit $?"
python scripts/bench_configs.py > gpurun_out/r02_bench_other_configs.jsonl 2> gpurun_out/bench_configs.err; echo "configs exit $?"
python scripts/bench_made_inverse.py 4096 32768 262144 > gpurun_out/r02_made_inverse_bench.jsonl 2>&1; echo "inverse exit $?"
python scripts/bench_kernels.py > gpurun_out/r02_kernel_microbench.txt 2>&1; echo "kernels exit $?"
