#!/bin/bash
# sanity of the current tree on a fresh box: GPU parity tests, warp-primitive latencies, default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
./tools/warplat > gpurun_out/warplat.txt 2>&1; cat gpurun_out/warplat.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_default.json
