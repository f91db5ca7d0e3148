#!/bin/bash
# Round-1 (second session) final GPU call: full GPU suite, the bench lines of every BASELINE config,
# ncu launch lists and --set full captures of the dominant kernels.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu_final.log
tail -3 gpurun_out/pytest_gpu_final.log
b() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; python tools/brief.py "$name" < gpurun_out/$name.json; tail -1 gpurun_out/$name.err; }
b r1b_bench203                                   # config 4, the headline line
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1b_ref.json 2> gpurun_out/r1b_ref.err
b r1b_tvd203 --no-cpu --limiter bj --gradient lsq --cfl 0.4 --shock 1 --steps 10     # config 4 with the limiter (SURVEY 8d input)
b r1b_sphere_visc --workload sphere --size 42 --viscous 1 --steps 20 --cpu-n 16      # config 3
b r1b_step_ausm1 --workload step --size 445 --flux ausm --order 1 --graph 1 --steps 50 --no-cpu   # config 2
b r1b_step_roe2lim --workload step --size 445 --flux roe --order 2 --limiter bj --graph 1 --steps 50 --no-cpu
b r1b_lusgs203 --workload lusgs --size 203 --steps 3 --warmup 3 --cpu-n 32           # config 5: the sweeps at 50 M rows
b r1b_imp160 --size 160 --steps 3 --warmup 3 --implicit 1 --no-cpu                   # config 5: implicit step, 24.6 M tets
# ---- ncu: launch lists (kernel share of the step) ----
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches_implicit.csv \
    python bench.py --size 64 --steps 1 --warmup 3 --implicit 1 --no-cpu > gpurun_out/ncu_l2.log 2>&1
# ---- ncu --set full: fused kernel (default), limited + viscous variant, LU-SGS sweep level ----
ncu --set full --clock-control none --import-source on -k regex:k_step_tiles --launch-skip 6 -c 1 -o gpurun_out/r1b_prof_tiles -f \
    python bench.py --size 128 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_f1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_tiles --launch-skip 6 -c 1 -o gpurun_out/r1b_prof_tiles_lim -f \
    python bench.py --size 128 --steps 2 --warmup 3 --no-cpu --limiter venkat --limiter-k 1 > gpurun_out/ncu_f2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_level --launch-skip 40 -c 2 -o gpurun_out/r1b_prof_sweep -f \
    python bench.py --workload lusgs --size 96 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_f3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ux --launch-skip 10 -c 1 -o gpurun_out/r1b_prof_ux -f \
    python bench.py --workload lusgs --size 96 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_f4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
python __graft_entry__.py smoke 2>&1 | tail -3
