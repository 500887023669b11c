#!/bin/bash
export AMDP_B200_LIB=build/libamdp_ldg256.so
for g in 0 32 64 128; do
  echo "### L2GRAN=$g"
  L2GRAN=$g python tools_l2gran.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload cfg2 --envs 65536 --ticks 128 --search thread 2>&1 | grep -E "set L2|metric" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'): d = json.loads(l); print('%.3e' % d['value'])
    else: print(l.strip())"
done
