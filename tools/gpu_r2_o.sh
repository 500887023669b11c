#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tasks.py tests/test_gpu_bench_workloads.py tests/test_gpu_policy.py -x -q -k "warp or redux or cfg2 or reference_sampled or fast_rollout or fuzz or policy or packed" 2>&1 | tail -6
run() { label=$1; shift; timeout 600 env "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label: %.4e steps/s  kern %.4f ms  tick %.1f ns' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['kernel_ms_per_launch']*1e6/512))"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other --no-e2e"
for n in 512 4096; do
for i in 1 2; do
run "lag2   $n" X=1 $B --envs $n --tasks $n
run "nolag2 $n" AMDP_B200_LIB=build/variants/libamdp_nolag2.so $B --envs $n --tasks $n
done
done
run "lag2   table" X=1 $B --policy table
run "nolag2 table" AMDP_B200_LIB=build/variants/libamdp_nolag2.so $B --policy table
