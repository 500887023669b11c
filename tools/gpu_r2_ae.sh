#!/bin/bash
# round 2, step ae: three-level rows, next tick's `top` node prefetched into L1 as soon as s' is known (variant) vs base
mkdir -p gpurun_out
{
for lib in "" build/variants/libamdp_tp_toppf.so; do
  echo "== lib=${lib:-base}"
  AMDP_B200_LIB=$lib timeout 600 python tools/ab_rate.py cfg4 65536 32 64 2>&1 | grep steps/s
  AMDP_B200_LIB=$lib timeout 600 python tools/ab_rate.py cfg4 131072 32 64 2>&1 | grep steps/s
  AMDP_B200_LIB=$lib timeout 600 python tools/ab_rate.py cfg4 8192 32 64 2>&1 | grep steps/s
done
} | tee gpurun_out/r2_ae.txt
