#!/bin/bash
# tile-size sweep of the fused kernel: tools/sweep.sh <n> "<T list>" "<NT list>"
n=${1:-128}
for T in ${2:-96 128 192 256}; do for NT in ${3:-128 256}; do
python bench.py --size $n --steps 10 --warmup 3 --no-cpu --tile-cells $T --block-threads $NT 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('T',d['config']['tile_cells'],'NT',d['config']['block_threads'],'ms/step',round(d['ms_per_step'],3),'Gcells/s',round(d['value']/1e9,3),'frac',round(d['roofline']['frac'],3),d['roofline']['kernels_ms'],'GiB',round(d['device_gib'],1))"
done; done
