#!/bin/bash
# round 2, step aa: two-level rows, draws under the block load (variant) vs before the search (base)
mkdir -p gpurun_out
{
for lib in "" build/variants/libamdp_tp_drawblk.so; do
  echo "== lib=${lib:-base}"
  AMDP_B200_LIB=$lib timeout 600 python tools/ab_rate.py cfg5 131072 131072 128 2>&1 | grep steps/s
  AMDP_B200_LIB=$lib timeout 600 python tools/ab_rate.py cfg2 65536 4096 128 2>&1 | grep steps/s
done
} | tee gpurun_out/r2_aa.txt
