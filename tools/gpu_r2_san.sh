#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r2_sanitizer_racecheck.txt
