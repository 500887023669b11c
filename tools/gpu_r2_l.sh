#!/bin/bash
# round 2, call L: A/B of the MAPREG variant of k_rollout_w (cfg2) and of the wide-register POMDP variant of k_rollout_tp (cfg3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_tasks.py tests/test_gpu_bench_workloads.py -x -q 2>&1 | tail -4
run() { label=$1; shift; timeout 600 env "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$label: %.4e steps/s  kern %.4f ms  frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"; }
B="python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-other --no-e2e"
for i in 1 2; do
run "cfg2 mapreg=1" AMDP_W_MAPREG=1 $B
run "cfg2 mapreg=0" AMDP_W_MAPREG=0 $B
done
run "cfg2 table mapreg=1" AMDP_W_MAPREG=1 $B --policy table
run "cfg2 table mapreg=0" AMDP_W_MAPREG=0 $B --policy table
for i in 1 2; do
run "cfg3 wide" AMDP_TP_MINB=w $B --workload cfg3 --ticks 64
run "cfg3 lo  " AMDP_TP_MINB=l $B --workload cfg3 --ticks 64
done
run "cfg3 default" $B --workload cfg3 --ticks 64
