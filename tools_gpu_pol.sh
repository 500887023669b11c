#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
run() {
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*: %.3e steps/s  kern %.4f ms frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"
}
run --workload cfg2
run --workload cfg2 --policy table
run --workload cfg2 --envs 65536 --ticks 128
run --workload cfg2 --envs 65536 --ticks 128 --policy table
run --workload cfg5 --ticks 128
run --workload cfg5 --ticks 128 --policy table
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072 --policy table
