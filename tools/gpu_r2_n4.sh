#!/bin/bash
# 4-GPU bench exactly as the driver launches it (round-2 final build)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
echo "rc=$?"; tail -3 gpurun_out/bench_n4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
e=d['e2e']
print('N=4 value %.3e ms/step %.4f kernel_ms %.4f | e2e %.3e' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], e['value']))
print(d.get('exchange_check'))
for k,v in d.get('other_workloads',{}).items():
    print(k, {a:(('%.3e'%b) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','roofline_frac','kernel','error')})
PY
