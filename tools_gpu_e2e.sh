#!/bin/bash
mkdir -p gpurun_out
for c in 32 64 128 256 512; do
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other --e2e-chunk $c 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c: dev %.3e  e2e %.3e  ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
for k in 2048; do
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other --e2e-chunk 128 --e2e-ticks $k 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ticks $k chunk 128: dev %.3e  e2e %.3e  ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
./tools_prof.sh tp5 --workload cfg4 --tasks 32 --ticks 64 --envs 131072 --no-other
