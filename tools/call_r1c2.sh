#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/bench_io.py --size 445 --out gpurun_out/r1c_io_bench.json > gpurun_out/r1c_io_bench.log 2>&1; echo "io bench rc=$?"; tail -1 gpurun_out/r1c_io_bench.log | cut -c1-1800
timeout 150 python tools/ab_variants.py --size 128 --steps 20 --out gpurun_out/r1c_ab128.json > gpurun_out/r1c_ab128.log 2>&1; echo "ab128 rc=$?"; tail -9 gpurun_out/r1c_ab128.log
timeout 240 python tools/ab_variants.py --size 203 --steps 20 --out gpurun_out/r1c_ab203.json > gpurun_out/r1c_ab203.log 2>&1; echo "ab203 rc=$?"; tail -9 gpurun_out/r1c_ab203.log
