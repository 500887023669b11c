#!/bin/bash
# round 2, step w: ncu of the GIVEN-action variant of k_rollout_w (why is it 15 % slower than the uniform-policy variant?)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rollout_w -s 25 -c 1 -o gpurun_out/prof_w_given python tools/ab_rate.py cfg2 4096 4096 512 given > gpurun_out/b_prof_w_given.log 2>&1
tail -3 gpurun_out/b_prof_w_given.log
ls -la gpurun_out/*.ncu-rep
