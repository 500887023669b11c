#!/bin/bash
# round 2, call C (2 GPUs): the peer-memory statistics exchange across processes (cudaIpc), vs NCCL, driver-style runs
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo2.txt 2>&1; lscpu | grep -i -E "numa|socket|^cpu\(s\)" >> gpurun_out/topo2.txt
run() { tag=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" > gpurun_out/bench_c_$tag.json 2> gpurun_out/bench_c_$tag.err; echo "$tag rc=$?"; tail -2 gpurun_out/bench_c_$tag.err | cut -c1-300
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_c_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    print('  value %.3e ms/step %.4f kern %.4f e2e %s launches %d  stats: %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e'] and '%.3e' % d['e2e']['value'], d['gpu_launches'], d['config']['statistics_per_step'][:60]))
    print('  check', d.get('exchange_check'), d['clocks'])
    for k,v in (d.get('other_workloads') or {}).items():
        print('  ', k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan'))) if 'value' in v else print('  ', k, v)
except Exception as e:
    print('  parse failed', e)
PY
}
run peer20 --steps 20 --warmup 3 --no-other
run nccl20 --steps 20 --warmup 3 --no-other --allreduce nccl --no-e2e
run none20 --steps 20 --warmup 3 --no-other --allreduce none --no-e2e
run peer20b --steps 20 --warmup 3 --no-other --no-e2e
run peer200 --steps 200 --warmup 3 --no-e2e
timeout 300 python bench.py --steps 20 --warmup 3 --no-other --no-e2e --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 value %.3e ms/step %.4f' % (d['value'], d['ms_per_step']), d['clocks'])"
