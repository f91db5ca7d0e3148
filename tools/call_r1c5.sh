#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/ab_variants.py --workload step --size 445 --steps 200 --variants 0,16,8,0 --tiles 0,256 --out gpurun_out/r1c_step_variants.json > gpurun_out/r1c_step_variants.log 2>&1; echo "rc=$?"; tail -10 gpurun_out/r1c_step_variants.log
