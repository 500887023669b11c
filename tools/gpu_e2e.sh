#!/bin/bash
# end-to-end (host buffers) sweep of the staging chunk size
mkdir -p gpurun_out
for c in 16 32 64 128 256; do
  timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other --e2e-chunk $c 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c: dev %.3e  e2e %.3e  ms %.3f' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed" 2>&1 | tail -2
