#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_parity_gpu.py tests/test_extension_gpu.py -q -m gpu -k "viscous or sphere or implicit_step_matches" 2>&1 | tail -15 > gpurun_out/pytest_gpu6.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; tail -1 gpurun_out/$name.err; }
b sphere_vt --workload sphere --size 42 --viscous 1 --steps 10 --no-cpu
b sphere_vt192 --workload sphere --size 42 --viscous 1 --steps 10 --no-cpu --tile-cells 192
b sphere_vs --workload sphere --size 42 --viscous 1 --steps 10 --no-cpu --kernel split
tail -15 gpurun_out/pytest_gpu6.log
