mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cubic.py tests/test_gpu_tile_ring.py tests/test_gpu_made_inverse.py -m gpu -q --timeout 600 > gpurun_out/pytest_quad.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_quad.log | cut -c1-300 | tail -8
timeout 300 python scripts/bench_tile_ring.py 2>&1 | grep -E "quad|cubic|sos" | cut -c1-150
timeout 300 python scripts/bench_tile_ring.py --backward 2>&1 | grep -E "quad|cubic|sos" | cut -c1-150
