mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_made_inverse.py -m gpu -q -s --timeout 300 > gpurun_out/made_inv_tests.log 2>&1; echo "inverse pytest rc=$?"; grep -E "passed|failed|^FAILED|inverse outputs:|^E  " gpurun_out/made_inv_tests.log | cut -c1-260 | tail -14
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "flow_matches_reference or graph" 2>&1 | tail -4
