#!/bin/bash
# round 2, call B: fixed/new tests, then the default bench line (driver-style: --steps 20 --warmup 3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reference_tasks.py tests/test_gpu_round2_fixes.py -x -q 2>&1 | tail -15
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_b.json').read().strip().splitlines()[-1])
print('value %.3e frac %.3f kernel %s e2e %.3e launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['value'], d['gpu_launches']))
print('cpu', {k: v for k, v in d['cpu_baseline'].items() if k != 'sample'})
print('lat', d['roofline'].get('latency_ceiling'))
for k,v in d['other_workloads'].items():
    print(k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan')), v.get('bound')) if 'value' in v else print(k, {a: b for a, b in v.items() if a != 'workload'})
print(d['clocks'])
PY
