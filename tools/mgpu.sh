#!/bin/bash
# tools/mgpu.sh <ngpus> <size> [extra bench args]  -- multi-rank bench through torchrun
N=$1; n=$2; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
  bench.py --gpus $N --size $n --steps 10 --warmup 3 "$@" 2>gpurun_out/mgpu_${N}_${n}.err | tee gpurun_out/mgpu_${N}_${n}.json | python tools/brief.py "N=$N n=$n"
grep -E "\[bench\]|Error|error" gpurun_out/mgpu_${N}_${n}.err | tail -6
