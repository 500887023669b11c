#!/bin/bash
# round 2, call E (8 GPUs): scaling of the device-timed value (peer exchange vs NCCL vs none) and of the e2e path; D2H ceiling
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|^cpu\(s\)|model name"; free -g | head -2; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done ) > gpurun_out/topo8.txt 2>&1
NP=${NP:-8}
tr() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29512 "$@"; }
tr tools/d2h_ceiling.py 2>/dev/null | grep d2h
run() { tag=$1; shift; tr bench.py --gpus $NP "$@" > gpurun_out/bench_e_$tag.json 2> gpurun_out/bench_e_$tag.err; echo "$tag rc=$?"; grep -v -E "OMP_NUM|^\*+$|^$" gpurun_out/bench_e_$tag.err | tail -2 | cut -c1-300
python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/bench_e_%s.json' % sys.argv[1]).read().strip().splitlines()[-1])
    e=d['e2e']
    print('  value %.3e ms/step %.4f kern %.4f launches %d  stats: %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches'], d['config']['statistics_per_step'][:50]))
    if e: print('  e2e %.3e (%.1f GB/s/gpu) packed8sync %.3e' % (e['value'], e['d2h_GBps_per_gpu'], e['packed8_synchronous']['value']), e['host_buffers'])
    print('  check', d.get('exchange_check'), d['clocks'])
    for k,v in (d.get('other_workloads') or {}).items():
        print('  ', k, '%.3e' % v['value'], v.get('kernel'), 'frac %.3f' % v.get('roofline_frac', float('nan'))) if 'value' in v else print('  ', k, v)
except Exception as e:
    print('  parse failed', e)
PY
}
run peer20 --steps 20 --warmup 3
run nccl20 --steps 20 --warmup 3 --no-other --allreduce nccl --no-e2e
run none20 --steps 20 --warmup 3 --no-other --allreduce none --no-e2e
run peer20lag1 --steps 20 --warmup 3 --no-other --no-e2e --lag 1
run peer20wc --steps 20 --warmup 3 --no-other --e2e-wc
run peer20nonuma --steps 20 --warmup 3 --no-other --no-numa-bind
run peer200 --steps 200 --warmup 3 --no-other --no-e2e
