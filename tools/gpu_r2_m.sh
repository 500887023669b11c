#!/bin/bash
run() { label=$1; shift; timeout 600 env "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); n=d['config']['envs_per_gpu']; print('$label: %.4e steps/s  kern %.4f ms  tick %.1f ns  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['kernel_ms_per_launch']*1e6/512, d['roofline']['kernel']))"; }
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other --no-e2e --search warp"
for n in 512 1024 2048 3072 4096 4144; do run "envs $n" X=1 $B --envs $n --tasks $n; done
