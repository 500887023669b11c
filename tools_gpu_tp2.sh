#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "thread_per_env" > gpurun_out/pytest_tp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tp.log
tail -4 gpurun_out/pytest_tp.log
run() {
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other --no-e2e "$@" 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*: %.3e steps/s  kern %.4f ms frac %.3f  %s bytes %.1f GB' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel'], d['config']['table_bytes_per_gpu']/1e9))"
}
for t in 4 32; do run --workload cfg4 --tasks $t --ticks 64; done
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072
