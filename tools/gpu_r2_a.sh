#!/bin/bash
# round 2, call A: new parity tests + smoke + host topology facts
mkdir -p gpurun_out
( nvidia-smi topo -m; lscpu | grep -i -E "numa|model name|socket|^cpu\(s\)"; nproc; free -g | head -2; for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q 0x0302 $d/class; then echo $d $(cat $d/numa_node) $(cat $d/local_cpulist); fi; done ) > gpurun_out/topo.txt 2>&1
timeout 1500 python -m pytest tests/test_gpu_bench_workloads.py tests/test_gpu_reference_tasks.py tests/test_gpu_round2_fixes.py -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_bench_workloads.py --deselect tests/test_gpu_reference_tasks.py --deselect tests/test_gpu_round2_fixes.py 2>&1 | tail -5
