#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/ab_variants.py --size 128 --steps 20 --block-threads 128 --variants 8,64,0 --tiles 192,240 --out gpurun_out/r1c_nt128.json > gpurun_out/r1c_nt128.log 2>&1; echo "rc=$?"
grep tile_cells gpurun_out/r1c_nt128.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(' T',d['tile_cells'],'var',d['variant'],'ms',round(d['ms_per_step'],4),'graph_ms',round(d['graph_ms_per_step'],4),'same',d['identical_to_v0'])"
