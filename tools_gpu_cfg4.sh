#!/bin/bash
mkdir -p gpurun_out
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*: %.3e steps/s  kern %.4f ms frac %.3f  %s bytes %.1f GB' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel'], d['config']['table_bytes_per_gpu']/1e9))"
}
for t in 4 8 16 32; do run --workload cfg4 --tasks $t --ticks 64; done
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072
run --workload cfg4 --tasks 8 --ticks 64 --envs 131072
