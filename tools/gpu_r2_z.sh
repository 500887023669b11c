#!/bin/bash
# round 2, step z: full GPU suite (transition_gt label plane, ABI 3), default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_z.json 2> gpurun_out/bench_z.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_z.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.3e  kernel %.4f ms  frac %.3f | e2e %.3e (%.1f GB/s d2h, %.3f ms/step)' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], e['value'], e['d2h_GBps_per_gpu'], e['ms_per_step']))
for k, w in d.get('other_workloads', {}).items():
    if isinstance(w, dict):
        print(k, {x: (round(v, 4) if isinstance(v, float) and v < 100 else v) for x, v in w.items() if x in ('value', 'env_steps_per_s', 'roofline_frac', 'kernel', 'frac')})
PY
