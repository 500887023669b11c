#!/bin/bash
# One GPU-box call that establishes the state of the round: parity tests, smoke, bench lines, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -c 3000 gpurun_out/bench_cfg2.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
for args in "--workload cfg2 --envs 65536 --ticks 128" "--workload cfg5 --ticks 128" "--workload cfg3 --ticks 64" "--workload cfg4 --tasks 32 --ticks 64" "--workload cfg4 --tasks 32 --ticks 64 --envs 131072"; do
  echo "== $args"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $args 2>>gpurun_out/sweep.err | tee -a gpurun_out/sweep.jsonl | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['algorithmic_bytes_per_env_step'], d['config']['table_bytes_per_gpu'], d['config']['upload_s'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
tail -5 gpurun_out/launches_cfg2.csv
