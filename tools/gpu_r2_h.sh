#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_task_sampler.py -x -q 2>&1 | tail -15
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_h.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.3e frac %.3f kernel %s launches %d' % (d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['gpu_launches']))
print('e2e %.3e (%.1f GB/s) packed8 %.3e ceiling' % (e['value'], e['d2h_GBps_per_gpu'], e['packed8_synchronous']['value']), e['d2h_ceiling'])
for k,v in d['other_workloads'].items():
    print(k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan')), v.get('bound')) if 'value' in v else print(k, {a: b for a, b in v.items() if a not in ('workload','reference_cpu','kernels')})
print(d['clocks'])
PY
