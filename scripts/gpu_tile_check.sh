mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_guard.py tests/test_gpu_cubic.py -m gpu -q -x --timeout 900 > gpurun_out/pytest_tile.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^ERROR|Error" gpurun_out/pytest_tile.log | cut -c1-300 | tail -8
timeout 400 python scripts/bench_tile_ring.py > gpurun_out/tile_fwd.log 2>&1; echo "fwd rc=$?"; grep -E "tile ring|identical: False" gpurun_out/tile_fwd.log | cut -c1-160
timeout 600 python scripts/bench_tile_ring.py --backward --sweep rqs_bwd,affine_bwd,quadspline_bwd,linspline_bwd > gpurun_out/tile_bwd.log 2>&1; echo "bwd rc=$?"; grep -v "^sweep" gpurun_out/tile_bwd.log | cut -c1-170
