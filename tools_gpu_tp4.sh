#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_env or packed" > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
tail -3 gpurun_out/pytest_tp.log
run() {
  lib=$1; shift
  p=xenoverse_b200/libamdp_b200.so; [ $lib != base ] && p=build/variants/libamdp_$lib.so
  AMDP_B200_LIB=$PWD/$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib $*: %.3e steps/s  kern %.4f ms frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"
}
for lib in base tpearly; do
run $lib --workload cfg4 --tasks 32 --ticks 64
run $lib --workload cfg4 --tasks 32 --ticks 64 --envs 131072
run $lib --workload cfg2 --envs 65536 --ticks 128
run $lib --workload cfg5 --ticks 128
done
