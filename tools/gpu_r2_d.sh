#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2_fixes.py -x -q 2>&1 | tail -8
python tools/d2h_ceiling.py
for extra in "" "--e2e-wc" "--e2e-chunk 64" "--e2e-chunk 256 --e2e-ticks 1024"; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-other --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('[$extra] value %.3e e2e %.3e (%.1f GB/s d2h, %.3f ms/step) packed8 sync %.3e (%.1f GB/s)' % (d['value'], e['value'], e['d2h_GBps_per_gpu'], e['ms_per_step'], e['packed8_synchronous']['value'], e['packed8_synchronous']['d2h_GBps_per_gpu']), e['host_buffers'])
"
done
