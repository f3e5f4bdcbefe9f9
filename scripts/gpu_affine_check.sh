mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioner.py -m gpu -q --timeout 600 -k affine > gpurun_out/pytest_cond.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_cond.log | cut -c1-300 | tail -12
timeout 300 python scripts/bench_affine_fused.py > gpurun_out/affine_fused.jsonl 2> gpurun_out/affine_fused.err; echo "affine bench rc=$?"; cat gpurun_out/affine_fused.jsonl; tail -3 gpurun_out/affine_fused.err
timeout 600 python scripts/fuzz_conditioner.py 12 150 > gpurun_out/fuzz_cond.log 2>&1; echo "fuzz rc=$?"; grep -c "affine.*fused" gpurun_out/fuzz_cond.log; tail -2 gpurun_out/fuzz_cond.log | cut -c1-300
