#!/bin/bash
# round 2, step ad: wide rows vs block lines with fewer tasks (is the first level L2-resident?)
mkdir -p gpurun_out
{
for t in 4 8 16; do
  for w in 0 1; do
    AMDP_TP_WIDE=$w timeout 600 python tools/ab_rate.py cfg4 65536 $t 64 2>&1 | grep steps/s | sed "s/^/wide=$w /"
  done
done
} | tee gpurun_out/r2_ad.txt
