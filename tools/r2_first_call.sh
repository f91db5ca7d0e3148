#!/bin/bash
# First GPU call of the next round: what round 1 built but could not time (its GPU budget was spent).
#   gpurun --timeout 900 -- 'bash tools/r2_first_call.sh'
mkdir -p gpurun_out
b() { name=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; tail -1 gpurun_out/$name.err; }
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
# 1. fused LU-SGS iteration: mode 0 vs mode 1, the sweeps alone and inside the implicit step
for m in 0 1; do
  MSTGPU_LUSGS_MODE=$m b r2_lusgs128_mode$m --workload lusgs --size 128 --steps 3 --warmup 3 --no-cpu
  MSTGPU_LUSGS_MODE=$m b r2_imp96_mode$m --size 96 --steps 3 --warmup 3 --implicit 1 --no-cpu
done
# 2. 128-thread CTAs, 4 per SM: tile sizes around the 56 KB shared-memory class limit
timeout 200 python tools/ab_variants.py --size 203 --steps 20 --block-threads 128 --variants 8 --tiles 224,240,248 --out gpurun_out/r2_nt128_203.json 2>&1 | grep tile_cells | cut -c1-200
# 3. the two bench workloads added at the end of round 1 (GPU arm never run)
b r2_sod --workload sod --steps 200 --warmup 5 --graph 1
timeout 100 python bench.py --impl reference --workload sod --steps 20 --warmup 3 > gpurun_out/r2_sod_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2_sod_ref.json
