#!/bin/bash
mkdir -p gpurun_out
timeout 150 python tools/ab_variants.py --size 128 --steps 20 --variants 0 --tiles 512,440,544,552,560 --out gpurun_out/r1c_tiles128.json > gpurun_out/r1c_tiles128.log 2>&1; echo "tiles128 rc=$?"; tail -6 gpurun_out/r1c_tiles128.log
timeout 200 python tools/ab_variants.py --size 203 --steps 20 --variants 0 --tiles 512,552,560 --out gpurun_out/r1c_tiles203.json > gpurun_out/r1c_tiles203.log 2>&1; echo "tiles203 rc=$?"; tail -4 gpurun_out/r1c_tiles203.log
