mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python scripts/bench_kernels.py > gpurun_out/kernel_microbench.jsonl 2> gpurun_out/kernel_microbench.err; echo "microbench rc=$?"; grep -c kernel gpurun_out/kernel_microbench.jsonl
timeout 600 python bench.py --workload cfg3_train > gpurun_out/bench_cfg3_train.json 2> gpurun_out/bench_cfg3_train.err; echo "cfg3_train rc=$?"; cut -c1-400 gpurun_out/bench_cfg3_train.json
