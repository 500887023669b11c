#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "warp_per_env or filter_fallback" > gpurun_out/pytest_w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_w.log
tail -3 gpurun_out/pytest_w.log
run() {
  echo "== $*"
  timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-other "$@" 2>>gpurun_out/sweep.err | tee -a gpurun_out/sweep_w.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e steps/s  %.3f ms  kern %.3f ms frac %.3f  kernel %s e2e %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e'] and '%.3e' % d['e2e']['value']))"
}
run --workload cfg2
run --workload cfg2
