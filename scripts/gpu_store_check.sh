mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioner.py -m gpu -q --timeout 600 > gpurun_out/pytest_store.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|^E  " gpurun_out/pytest_store.log | cut -c1-300 | tail -12
timeout 400 python scripts/bench_store_fused.py > gpurun_out/store_fused.jsonl 2> gpurun_out/store_fused.err; echo "bench rc=$?"; cat gpurun_out/store_fused.jsonl; tail -3 gpurun_out/store_fused.err
