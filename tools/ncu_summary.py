#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + the hottest SASS lines of the loop body.
usage: tools/ncu_summary.py <report.ncu-rep> [ticks_per_launch*warps]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else None
def page(name):
    # either a .ncu-rep (needs ncu) or the prefix of <prefix>_raw.csv / <prefix>_source.csv exported on the GPU box
    if rep.endswith(".ncu-rep"):
        return subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return open(f"{rep}_{name}.csv").read()
raw = page("raw")
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size', 'lts__t_bytes.sum', 'dram__bytes_read.sum.per_second',
        'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg']
for i, h in enumerate(hdr):
    if h in keys:
        print(f"{h:70s} {units[i]:12s} {vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(vals[i] or 0) > 0.15:
        print(f"{h:90s} {vals[i]}")
src = page("source")
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['Instructions Executed']]) for r in rows[2:])
smp = sum(int(r[ix['# Samples']]) for r in rows[2:])
print("total warp-instr", tot, "samples", smp, "per unit", tot / per if per else "")
if per:
    print("--- loop body (>=0.4 exec per tick) : addr exec/tick samples long_sb wait short_sb sass")
    for r in rows[2:]:
        n = int(r[ix['Instructions Executed']])
        if n >= 0.4 * per or int(r[ix['# Samples']]) > 0.01 * smp:
            print(r[ix['Address']][-4:], f"{n/per:5.2f}", f"{int(r[ix['# Samples']]):6d}", r[ix['stall_long_sb']], r[ix['stall_wait']], r[ix['stall_short_sb']], r[ix['Source']].strip()[:80])
