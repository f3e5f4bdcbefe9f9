mkdir -p gpurun_out
timeout 200 python scripts/ncu_conditioner_variants.py; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conditioner_f16x3 -o gpurun_out/r02_cond_variants python scripts/ncu_conditioner_variants.py > gpurun_out/ncu_variants.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_variants.log
