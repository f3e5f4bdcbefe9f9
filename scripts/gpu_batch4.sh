mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
for i in 1 2 3; do timeout 600 python -m pytest tests/test_reference_suite.py -m gpu -q --timeout 600 2>&1 | tail -1; done
