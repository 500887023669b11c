#!/bin/bash
mkdir -p gpurun_out
prof() { tag=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -f -o /tmp/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_prof_$tag.log 2>&1
  tail -1 gpurun_out/b_prof_$tag.log | cut -c1-120
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv > gpurun_out/prof_${tag}_source.csv
}
prof tp3c_cfg2_64k --workload cfg2 --envs 65536 --ticks 128 --search thread
