#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>>gpurun_out/sweep.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e steps/s  %.3f ms  frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))"
}
for lib in "" build/libamdp_siv3a.so; do
export AMDP_B200_LIB=$lib; echo "#### lib=$lib"
run --workload cfg2 --envs 65536 --ticks 128 --search thread
run --workload cfg2 --envs 131072 --ticks 128 --search thread
run --workload cfg4 --tasks 32 --ticks 64 --search thread
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072 --search thread
done
export AMDP_B200_LIB=

