#!/bin/bash
# GPU call: extension parity tests, then the existing suite, then a few short bench lines
mkdir -p gpurun_out
python -m pytest tests/test_extension_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/ext_tests.log
python -m pytest tests -x -q -m gpu --deselect tests/test_extension_gpu.py 2>&1 | tail -15 > gpurun_out/pytest_gpu2.log
b() { name=$1; shift; timeout 300 python bench.py --no-cpu "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; }
b lim_tiles --size 128 --steps 10 --limiter bj --gradient lsq --cfl 0.4 --shock 1
b lim_split --size 128 --steps 10 --limiter bj --gradient lsq --cfl 0.4 --shock 1 --kernel split
b lsq_tiles --size 128 --steps 10 --gradient lsq
b base128 --size 128 --steps 10
b step_o1 --workload step --size 445 --flux ausm --order 1 --graph 1 --steps 50
for c in 128:128 192:128 256:256 384:256; do T=${c%%:*}; NT=${c##*:}
b step_o1_T$T --workload step --size 445 --flux ausm --order 1 --graph 1 --steps 50 --tile-cells $T --block-threads $NT
done
b step_o2_lim --workload step --size 445 --flux ausm --order 2 --limiter venkat --limiter-k 1 --graph 1 --steps 50
cat gpurun_out/ext_tests.log | tail -30
cat gpurun_out/pytest_gpu2.log | tail -5
