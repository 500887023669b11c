#!/bin/bash
# round-end style validation on one GPU: all GPU tests, smoke, the driver's bench command, the reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.3e frac %.3f kernel %s launches %d ms/step %.4f' % (d['value'], d['roofline']['frac'], d['roofline']['kernel'], d['gpu_launches'], d['ms_per_step']))
print('e2e %.3e (%.1f GB/s) packed8 %.3e ceiling %.1f' % (e['value'], e['d2h_GBps_per_gpu'], e['packed8_synchronous']['value'], e['d2h_ceiling']['aggregate_GBps']))
print('cpu', {k: v for k, v in d['cpu_baseline'].items() if k != 'sample'})
for k,v in d['other_workloads'].items():
    print(k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan')), v.get('bound')) if 'value' in v else print(k, {a: b for a, b in v.items() if a not in ('workload','reference_cpu','kernels')})
print(d['clocks'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-300
