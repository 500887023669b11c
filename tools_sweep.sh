for args in "--workload cfg2 --envs 65536 --ticks 128" "--workload cfg5 --ticks 128" "--workload cfg3 --ticks 64" "--workload cfg4 --tasks 32 --ticks 64"; do
  echo "== $args"; python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $args 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['algorithmic_bytes_per_env_step'], d['config']['table_bytes_per_gpu'], d['config']['upload_s'])"
done
