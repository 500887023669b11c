#!/bin/bash
# usage: tools/gpu_var.sh "<bench args>" name1 name2 ...   (name "base" = the in-tree library)
mkdir -p gpurun_out
args=$1; shift
for v in "$@"; do
  lib=build/variants/libamdp_$v.so; [ $v = base ] && lib=xenoverse_b200/libamdp_b200.so
  for rep in 1 2; do
  AMDP_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-other --no-e2e $args 2>>gpurun_out/var.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v: %.3e steps/s  kern %.4f ms frac %.3f  %s' % (d['value'], d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'], d['roofline']['kernel']))"
  done
done
