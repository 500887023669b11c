#!/bin/bash
# round 2, call K: ncu evidence — launch list of the bench command, full captures of the kernels behind the round-2 numbers
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. launch list of the driver-style bench command (kernel's SHARE of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other > gpurun_out/b_ncu.log 2>&1
echo "launch list rows: $(grep -c k_ gpurun_out/r2_launches_cfg2.csv)"
# 2. headline kernel
timeout 600 $NCU -k regex:k_rollout_w -s 3 -c 1 -o /tmp/prof_r2_w_cfg2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_w_cfg2.ncu-rep $((4096*512)) > gpurun_out/r2_w_cfg2.txt 2>&1; head -12 gpurun_out/r2_w_cfg2.txt
# 3. statistics kernel (single rank: k_stats without exchange)
timeout 600 $NCU -k regex:k_stats -s 3 -c 1 -o /tmp/prof_r2_stats python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_stats.ncu-rep > gpurun_out/r2_stats.txt 2>&1; head -6 gpurun_out/r2_stats.txt
# 4. cfg4 at 65536 envs and 131072 envs, 32 tasks
for e in 65536 131072; do
timeout 900 $NCU -k regex:k_rollout_tp -s 3 -c 1 -o /tmp/prof_r2_tp3_cfg4_$e python bench.py --workload cfg4 --envs $e --tasks 32 --ticks 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_tp3_cfg4_$e.ncu-rep $((e*64/32)) > gpurun_out/r2_tp3_cfg4_$e.txt 2>&1; head -8 gpurun_out/r2_tp3_cfg4_$e.txt
done
# 5. the new kernels: multi-token rollout, value iteration per task, task sampler
cat > /tmp/drv.py <<'PY'
import sys, torch, numpy
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
what = sys.argv[1]
if what == "mt":
    print(bench.measure_mtpomdp(0, dev, steps=2)["value"])
elif what == "vi":
    print(bench.measure_vi("cfg2", 4096, 128, 5, dev, 0, sweeps=64)["achieved_GBps"])
elif what == "sampler":
    from xenoverse_b200.task_sampler import sample_mdp_candidates, check_valuefunction_device
    c = sample_mdp_candidates(592, 128, 5, seed=1)
    r = check_valuefunction_device(c, 128, numpy.full(592, 400.0))
    torch.cuda.synchronize(); print(int(r["accept"].sum()))
elif what == "learn":
    print(bench.measure_learner(0, dev, "q", 4096, 4096, 128, 5, 64, steps=2)["value"])
PY
timeout 600 $NCU -k regex:k_rollout_mt -s 3 -c 1 -o /tmp/prof_r2_mt python /tmp/drv.py mt > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_mt.ncu-rep $((65536*64/32)) > gpurun_out/r2_mt.txt 2>&1; head -8 gpurun_out/r2_mt.txt
timeout 600 $NCU -k regex:k_vi_task -s 1 -c 1 -o /tmp/prof_r2_vi_task python /tmp/drv.py vi > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_vi_task.ncu-rep > gpurun_out/r2_vi_task.txt 2>&1; head -10 gpurun_out/r2_vi_task.txt
timeout 900 $NCU -k "regex:k_sample_mdp|k_filter_values|k_square_batch" -c 4 -o /tmp/prof_r2_sampler python /tmp/drv.py sampler > /dev/null 2>&1
ncu -i /tmp/prof_r2_sampler.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keys=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','smsp__inst_executed.sum']
ix=[h.index(k) for k in keys if k in h]
for r in rows[2:]: print(' | '.join(r[i][:60] for i in ix))
" > gpurun_out/r2_sampler.txt; cat gpurun_out/r2_sampler.txt
timeout 600 $NCU -k regex:k_rollout_learn -s 3 -c 1 -o /tmp/prof_r2_learn_q python /tmp/drv.py learn > /dev/null 2>&1
python tools/ncu_summary.py /tmp/prof_r2_learn_q.ncu-rep $((4096*64)) > gpurun_out/r2_learn_q.txt 2>&1; head -8 gpurun_out/r2_learn_q.txt
ls -la /tmp/prof_r2_*.ncu-rep | awk '{print $5, $9}'
