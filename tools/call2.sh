#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu3.log
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; }
b lusgs64 --workload lusgs --size 64 --steps 5 --no-cpu
b lusgs128 --workload lusgs --size 128 --steps 5 --cpu-n 32
b lim_bj --no-cpu --size 128 --steps 10 --limiter bj --gradient lsq --cfl 0.4 --shock 1
b lim_vk --no-cpu --size 128 --steps 10 --limiter venkat --limiter-k 1 --cfl 0.4 --shock 1
b lim_vk_T256 --no-cpu --size 128 --steps 10 --limiter venkat --limiter-k 1 --cfl 0.4 --shock 1 --tile-cells 256
b lim_split --no-cpu --size 128 --steps 10 --limiter bj --gradient lsq --cfl 0.4 --shock 1 --kernel split
b step_roe2_lim --no-cpu --workload step --size 445 --flux roe --order 2 --limiter bj --graph 1 --steps 50
tail -25 gpurun_out/pytest_gpu3.log
tail -3 gpurun_out/lusgs128.err
