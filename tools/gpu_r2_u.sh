#!/bin/bash
# round 2, step u: (1) k_rollout_w with L1-allocating row / payload loads (gather2: no_allocate 16-byte misses are the slow kind);
# (2) k_rollout_tp with fewer env lanes per warp (a warp waits for the slowest of its lanes' requests)
mkdir -p gpurun_out
{
echo "== k_rollout_w, cfg2: base"; timeout 300 python tools/ab_rate.py cfg2 4096 4096 512 given 2>&1 | grep steps/s
for v in row1 pay1 both1 pay2 row2 both2; do
  echo "== k_rollout_w, cfg2: variant $v"; AMDP_B200_LIB=build/variants/libamdp_w_$v.so timeout 300 python tools/ab_rate.py cfg2 4096 4096 512 2>&1 | grep steps/s
done
echo "== k_rollout_tp lanes per warp"
timeout 600 python tools/ab_rate.py cfg4 65536 32 64 AMDP_TP_LPW=32,16 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 32768 32 64 AMDP_TP_LPW=32,16,8 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 16384 32 64 AMDP_TP_LPW=32,16,8,4 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 8192 32 64 AMDP_TP_LPW=32,16,8,4,2 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg2 65536 4096 128 AMDP_TP_LPW=32,16 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg3 65536 4096 64 AMDP_TP_LPW=32,16 2>&1 | grep steps/s
} | tee gpurun_out/r2_u.txt
