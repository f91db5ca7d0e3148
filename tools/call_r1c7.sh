#!/bin/bash
# final validation of the session: full GPU suite with the rebuilt library, config-2 bench line, smoke
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r1c_pytest_gpu_final.log 2>&1; echo "full gpu suite rc=$?"; tail -3 gpurun_out/r1c_pytest_gpu_final.log
timeout 100 python bench.py --workload step --size 445 --flux ausm --order 1 --graph 1 --steps 50 --no-cpu > gpurun_out/r1c_step_ausm1.json 2> gpurun_out/r1c_step_ausm1.err; python tools/brief.py r1c_step_ausm1 < gpurun_out/r1c_step_ausm1.json
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -3
