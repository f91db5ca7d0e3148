#!/bin/bash
# tools/scale.sh "<n list>" "<kernel list>" [extra bench args]
for n in $1; do for k in $2; do
python bench.py --size $n --steps 10 --warmup 3 --no-cpu --kernel $k $3 2>/dev/null | python tools/brief.py "n=$n $k"
done; done
