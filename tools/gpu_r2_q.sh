#!/bin/bash
mkdir -p gpurun_out
python tools/sanitize_small.py 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2_sanitizer_racecheck.txt
