#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "== $*"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-other "$@" 2>>gpurun_out/sweep.err | tee -a gpurun_out/sweep_tp.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e steps/s  %.3f ms  frac %.3f  B_alg %d  bytes %d %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['algorithmic_bytes_per_env_step'], d['config']['table_bytes_per_gpu'], d['roofline']['kernel']))"
}
for m in lo hi; do export AMDP_TP_MINB=$m; echo "### MINB=$m"
run --workload cfg5 --ticks 128
run --workload cfg2 --envs 131072 --ticks 128
run --workload cfg2 --envs 98304 --ticks 128
run --workload cfg2 --envs 65536 --ticks 128
run --workload cfg4 --tasks 32 --ticks 64 --envs 131072
done
unset AMDP_TP_MINB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/b_launches.log 2>&1
tail -3 gpurun_out/launches_cfg2.csv | cut -c1-200
