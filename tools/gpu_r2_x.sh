#!/bin/bash
# k_rollout_w: timing (uniform / given), instruction counts per tick, parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "warp_per_env_redux or fuzz or reference_sampled or bench_workloads or given or cp_async or policy" 2>&1 | tail -3
{
timeout 300 python tools/ab_rate.py cfg2 4096 4096 512 given 2>&1 | grep steps/s
ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:k_rollout_w -s 21 -c 6 python tools/ab_rate.py cfg2 4096 4096 512 given 2>/dev/null | grep -E "k_rollout_w|smsp__inst_executed|gpu__time_duration" | paste - - - | awk '{print $1, $2, $3, $(NF-5), $(NF-3), $(NF-2), $NF}' | sort | uniq -c
} | tee gpurun_out/r2_x.txt
