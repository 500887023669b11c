#!/bin/bash
NP=${NP:-4}
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $NP --steps 20 --warmup 3 > gpurun_out/bench_p_$NP.json 2> gpurun_out/bench_p_$NP.err; echo "rc=$? in $(( $(date +%s) - t0 )) s"
grep -v -E "OMP_NUM|^\*+$|^$" gpurun_out/bench_p_$NP.err | tail -3 | cut -c1-300
python - "$NP" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/bench_p_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
e=d['e2e']
print('value %.3e ms/step %.4f kern %.4f launches %d' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches']))
print('e2e %.3e (%.1f GB/s/gpu) ceiling %.1f frac %.2f' % (e['value'], e['d2h_GBps_per_gpu'], e['d2h_ceiling']['aggregate_GBps'], e['d2h_ceiling']['frac_of_ceiling']))
print('check', d.get('exchange_check'), d['clocks'])
for k,v in (d.get('other_workloads') or {}).items():
    print('  ', k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan'))) if 'value' in v else print('  ', k, v)
PY
t0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $NP --steps 3 --warmup 1 | cut -c1-200; echo "reference arm in $(( $(date +%s) - t0 )) s"
