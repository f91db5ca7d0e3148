#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 120 python tools/ab_variants.py "$@" --out gpurun_out/$name.json > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; grep tile_cells gpurun_out/$name.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' T',d['tile_cells'],'var',d['variant'],'ms',round(d['ms_per_step'],4),'graph_ms',round(d['graph_ms_per_step'],4),'same',d['identical_to_v0'])"; }
run r1c_step_o2 --workload step --size 445 --steps 100 --order 2 --flux roe --variants 0,16,8
run r1c_step_o1roe --workload step --size 445 --steps 100 --order 1 --flux roe --variants 0,16,8
run r1c_box_o1 --workload box --size 128 --steps 20 --order 1 --variants 0,16,8 --tiles 512,384,256
run r1c_box_o2_3cta --workload box --size 128 --steps 20 --variants 0,16 --tiles 512,384
