mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conditioner.py tests/test_gpu_parity.py tests/test_gpu_guard.py -m gpu -q --timeout 600 > gpurun_out/pytest_pad.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_pad.log | cut -c1-300 | tail -12
timeout 600 python scripts/fuzz_conditioner.py 21 200 > gpurun_out/fuzz_cond_pad.log 2>&1; echo "fuzz rc=$?"; grep -c MISMATCH gpurun_out/fuzz_cond_pad.log; grep MISMATCH gpurun_out/fuzz_cond_pad.log | head -5 | cut -c1-250; tail -1 gpurun_out/fuzz_cond_pad.log
