#!/bin/bash
# round 2, step ac: WIDE rows (24-entry blocks, one dependent DRAM access per step) vs the 8-entry block lines on cfg4
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "thread_per_env or bench_workload or fuzz or ragged or sharding or checkpoint" 2>&1 | tail -5
{
for n in 65536 131072 8192; do
  for w in 0 1; do
    AMDP_TP_WIDE=$w timeout 600 python tools/ab_rate.py cfg4 $n 32 64 2>&1 | grep steps/s | sed "s/^/wide=$w /"
  done
done
} | tee gpurun_out/r2_ac.txt
