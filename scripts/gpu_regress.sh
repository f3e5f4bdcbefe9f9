mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 900 python scripts/fuzz_conditioner.py 31 250 > gpurun_out/fuzz_cond.log 2>&1; echo "fuzz cond rc=$?"; grep -c MISMATCH gpurun_out/fuzz_cond.log; grep MISMATCH gpurun_out/fuzz_cond.log | head -5 | cut -c1-250; tail -1 gpurun_out/fuzz_cond.log; grep -c "fused" gpurun_out/fuzz_cond.log
timeout 600 python scripts/fuzz_made_inverse.py 5 80 > gpurun_out/fuzz_made.log 2>&1; echo "fuzz made rc=$?"; grep -c MISMATCH gpurun_out/fuzz_made.log; tail -1 gpurun_out/fuzz_made.log | cut -c1-200
timeout 600 python scripts/fuzz_training.py 3 40 > gpurun_out/fuzz_train.log 2>&1; echo "fuzz train rc=$?"; grep -c MISMATCH gpurun_out/fuzz_train.log; tail -1 gpurun_out/fuzz_train.log | cut -c1-200
