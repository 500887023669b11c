#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() {
  echo "== $*"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>>gpurun_out/sweep.err | tee -a gpurun_out/sweep_tp.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e steps/s  %.3f ms  frac %.3f  B_alg %d  bytes %d' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['algorithmic_bytes_per_env_step'], d['config']['table_bytes_per_gpu']))"
}
run --workload cfg2 --envs 65536 --ticks 128 --search thread
run --workload cfg2 --envs 131072 --ticks 128 --search thread
run --workload cfg2 --envs 32768 --ticks 128 --search thread
run --workload cfg2 --envs 16384 --ticks 128 --search thread
run --workload cfg5 --ticks 128 --search thread
run --workload cfg4 --tasks 32 --ticks 64 --search thread
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072 --search thread
