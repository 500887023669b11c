#!/bin/bash
# round 2, step v: full GPU suite after the GIVEN / lanes-per-warp changes; GIVEN timing; draw-under-pv variant on cfg4
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
{
timeout 300 python tools/ab_rate.py cfg2 4096 4096 512 given 2>&1 | grep steps/s
echo "== cfg4 draw under pv"
timeout 600 python tools/ab_rate.py cfg4 65536 32 64 2>&1 | grep steps/s
AMDP_B200_LIB=build/variants/libamdp_tp_drawpv.so timeout 600 python tools/ab_rate.py cfg4 65536 32 64 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 131072 32 64 2>&1 | grep steps/s
AMDP_B200_LIB=build/variants/libamdp_tp_drawpv.so timeout 600 python tools/ab_rate.py cfg4 131072 32 64 2>&1 | grep steps/s
timeout 600 python tools/ab_rate.py cfg4 8192 32 64 2>&1 | grep steps/s
} | tee gpurun_out/r2_v.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-other --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('value %.3e  e2e %.3e (%.1f GB/s d2h, %.3f ms/step)' % (d['value'], e['value'], e['d2h_GBps_per_gpu'], e['ms_per_step']))"
