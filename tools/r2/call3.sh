#!/bin/bash
# round 2, call 3: own-cell weight derived (8 weights per face), staged packet stream (cp.async landing slots)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/r2c_pytest_gpu.log
MSTGPU_TILE_STAGED=1 timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_multigpu.py -x -q -m gpu -k "not full_size" > gpurun_out/r2c_pytest_staged.log 2>&1; echo "staged parity rc=$?"; tail -2 gpurun_out/r2c_pytest_staged.log
C="256:512:0:1,256:416:0:2,256:400:0:2,256:384:0:2,128:208:0:2,128:192:0:2,128:240:8:1"
timeout 300 python tools/ab_variants.py --size 128 --steps 20 --configs $C --out gpurun_out/r2c_ab128.json 2> gpurun_out/r2c_ab128.err | cut -c1-290
timeout 700 python tools/ab_variants.py --size 203 --steps 20 --configs $C --out gpurun_out/r2c_ab203.json 2> gpurun_out/r2c_ab203.err | cut -c1-290
tail -3 gpurun_out/r2c_ab203.err
