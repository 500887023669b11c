#!/bin/bash
# round-end style validation: all GPU tests, smoke, default bench, launch list of the bench command, reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print('value %.3e frac %.3f kernel %s e2e %.3e cpu %.3e launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches']))
for k,v in d['other_workloads'].items():
    print(k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan'))) if 'value' in v else print(k, {a: b for a, b in v.items() if a != 'workload'})
print(d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/b_ncu.log 2>&1
grep -c k_rollout gpurun_out/launches.csv
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
