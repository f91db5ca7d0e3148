#!/bin/bash
# Round-1 third session, the one GPU call (12 GPU-minutes left): most important first, every step logs on its own.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r1c_gpu.txt 2>&1
timeout 240 python -m pytest tests/test_output_gpu.py tests/test_reference_host_gpu.py -x -q -m gpu > gpurun_out/r1c_pytest_new.log 2>&1; echo "new tests rc=$?"; tail -3 gpurun_out/r1c_pytest_new.log
timeout 120 python tools/bench_io.py --size 445 --out gpurun_out/r1c_io_bench.json > gpurun_out/r1c_io_bench.log 2>&1; echo "io bench rc=$?"; tail -1 gpurun_out/r1c_io_bench.log | cut -c1-1500
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r1c_pytest_gpu.log 2>&1; echo "full gpu suite rc=$?"; tail -3 gpurun_out/r1c_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r1c_bench203.json 2> gpurun_out/r1c_bench203.err; echo "bench rc=$?"; python tools/brief.py r1c_bench203 < gpurun_out/r1c_bench203.json; tail -2 gpurun_out/r1c_bench203.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1c_ref.json 2> gpurun_out/r1c_ref.err; echo "ref arm rc=$?"; cut -c1-400 gpurun_out/r1c_ref.json
