#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_extension_gpu.py tests/test_lusgs_gpu.py -q -m gpu -k "implicit or lusgs or sweep or device" 2>&1 | tail -25 > gpurun_out/pytest_gpu4.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; tail -2 gpurun_out/$name.err; }
b imp128 --size 128 --steps 3 --warmup 3 --implicit 1 --cpu-n 32
tail -25 gpurun_out/pytest_gpu4.log
