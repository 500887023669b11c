#!/bin/bash
# round 2, step y: full GPU suite; ncu full capture of the uniform-policy k_rollout_w after the induction-variable change
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --set full --clock-control none --import-source on -k regex:k_rollout_w -s 5 -c 1 -o gpurun_out/prof_w_unif2 python tools/ab_rate.py cfg2 4096 4096 512 > gpurun_out/b_prof_w_unif2.log 2>&1
tail -2 gpurun_out/b_prof_w_unif2.log
timeout 300 python tools/ab_rate.py cfg2 4096 4096 512 given 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 65536 32 64 2>&1 | grep steps/s
