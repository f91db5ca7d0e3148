#!/bin/bash
# tools/sweep2.sh <n> "<renumber list>" "<T:NT list>"
n=$1
for r in $2; do for c in $3; do T=${c%%:*}; NT=${c##*:}
python bench.py --size $n --steps 10 --warmup 3 --no-cpu --renumber $r --tile-cells $T --block-threads $NT 2>/dev/null | python tools/brief.py "n=$n ren=$r T=$T NT=$NT"
done; done
