#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
python -m pytest tests/test_multigpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_mgpu2.log
tail -15 gpurun_out/pytest_mgpu2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --size 64 --steps 3 --warmup 3 --implicit 1 --no-cpu > gpurun_out/imp64_2gpu.json 2> gpurun_out/imp64_2gpu.err
python tools/brief.py imp64_2gpu < gpurun_out/imp64_2gpu.json; tail -3 gpurun_out/imp64_2gpu.err
