#!/bin/bash
# N-GPU bench exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 50 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 1 --impl reference > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -c 400 gpurun_out/bench_ref_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --workload cfg4 --tasks 256 --ticks 64 --no-e2e > gpurun_out/bench_cfg4_n$N.json 2> gpurun_out/bench_cfg4_n$N.err
tail -c 1200 gpurun_out/bench_cfg4_n$N.json; tail -3 gpurun_out/bench_cfg4_n$N.err
