#!/bin/bash
# round 2, step ab: cfg4 with few envs per GPU: thread-per-env (auto) vs the two-level warp-group kernel k_rollout_mdp<8,P2>
mkdir -p gpurun_out
{
for n in 8192 16384 32768; do
  timeout 600 python tools/ab_rate.py cfg4 $n 32 64 2>&1 | grep steps/s
  timeout 600 python tools/ab_rate.py cfg4 $n 32 64 search=two_level 2>&1 | grep "steps/s\|Error\|error" | tail -2
done
} | tee gpurun_out/r2_ab.txt
