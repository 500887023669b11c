#!/bin/bash
# usage: tools/gpu_quick.sh "<pytest -k expr>" — selected GPU tests, then the thread-per-env workloads
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -4
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*: %.3e steps/s  kern %.4f ms frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"
}
run --workload cfg5 --ticks 128
run --workload cfg3 --ticks 64
run --workload cfg2 --envs 65536 --ticks 128
