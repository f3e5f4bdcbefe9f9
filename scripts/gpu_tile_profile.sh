mkdir -p gpurun_out
timeout 300 python scripts/bench_tile_ring.py --backward --sweep sos_bwd,cubicspline_bwd > gpurun_out/tile_bwd2.log 2>&1; echo "bwd sweep rc=$?"
timeout 120 python scripts/ncu_tile_ring.py; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled_ -o gpurun_out/r02_tile_ring python scripts/ncu_tile_ring.py > gpurun_out/ncu_tile.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_tile.log
ls -la gpurun_out/r02_tile_ring.ncu-rep
