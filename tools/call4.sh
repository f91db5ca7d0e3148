#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_extension_gpu.py tests/test_parity_gpu.py tests/test_lusgs_gpu.py -q -m gpu -k "implicit or sphere or lusgs or sweep or device" 2>&1 | tail -15 > gpurun_out/pytest_gpu5.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; tail -1 gpurun_out/$name.err; }
b imp128b --size 128 --steps 3 --warmup 3 --implicit 1 --no-cpu
b sphere_v --workload sphere --size 42 --viscous 1 --steps 10 --no-cpu
b sphere_i --workload sphere --size 42 --steps 10 --no-cpu
tail -15 gpurun_out/pytest_gpu5.log
