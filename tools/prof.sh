#!/bin/bash
# usage: tools/prof.sh <tag> [bench args...]  — ncu full capture of the rollout kernel + short summary
tag=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_rollout -s 3 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/b_prof_$tag.log 2>&1
tail -1 gpurun_out/b_prof_$tag.log | cut -c1-300
