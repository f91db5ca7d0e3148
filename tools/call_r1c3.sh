#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_step_tiles --launch-skip 6 -c 1 -o gpurun_out/r1c_prof_tiles -f \
    python bench.py --size 128 --steps 2 --warmup 3 --no-cpu > gpurun_out/r1c_ncu_tiles.log 2>&1; echo "ncu tiles rc=$?"
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_node_fields --launch-skip 1 -c 1 -o gpurun_out/r1c_prof_nodes -f \
    python tools/bench_io.py --size 445 --no-ref > gpurun_out/r1c_ncu_nodes.log 2>&1; echo "ncu nodes rc=$?"
ls -la gpurun_out/*.ncu-rep
