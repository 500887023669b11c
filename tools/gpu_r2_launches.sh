#!/bin/bash
# the ncu launch list of the driver's bench command (per-launch gpu__time_duration; cold-cache, serialised)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-other --no-cpu-baseline > gpurun_out/b_launches.log 2>&1
tail -c 600 gpurun_out/b_launches.log
wc -l gpurun_out/r2_launches_cfg2.csv
