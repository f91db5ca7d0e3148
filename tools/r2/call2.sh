#!/bin/bash
# round 2, call 2: Roe flux contracted centrally first (fewer registers): parity suite, then CTA shapes A/B
mkdir -p gpurun_out
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/r2b_pytest_gpu.log
C="256:512:0,320:512:0,384:480:0,128:240:8,128:192:64,256:512:0"
timeout 300 python tools/ab_variants.py --size 128 --steps 20 --configs $C --out gpurun_out/r2b_ab128.json 2> gpurun_out/r2b_ab128.err | cut -c1-260
timeout 600 python tools/ab_variants.py --size 203 --steps 20 --configs $C --out gpurun_out/r2b_ab203.json 2> gpurun_out/r2b_ab203.err | cut -c1-260
