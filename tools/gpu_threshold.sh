#!/bin/bash
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*: %.3e steps/s  kern %.4f ms frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"
}
mkdir -p gpurun_out
for e in 4096 8192 16384; do
  run --workload cfg4 --tasks 32 --ticks 64 --envs $e --search warp
  run --workload cfg4 --tasks 32 --ticks 64 --envs $e --search thread
done
for e in 8192 16384; do
  run --workload cfg2 --ticks 256 --envs $e --search warp
  run --workload cfg2 --ticks 256 --envs $e --search thread
done
run --workload cfg5 --ticks 128 --envs 8192 --tasks 8192 --search warp
run --workload cfg5 --ticks 128 --envs 8192 --tasks 8192 --search thread
