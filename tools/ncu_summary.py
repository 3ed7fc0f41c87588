"""Summarise an ncu report exported with --page raw --csv and --page source --csv."""
import csv, re, sys, collections
raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors.sum', 'lts__t_bytes.sum.per_second', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
for w in want:
    idx = [i for i, h in enumerate(hdr) if h == w]
    if idx:
        i = idx[0]
        print(f"{w} [{units[i]}] = {[r[i] for r in data]}")
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ci, si, smp, ti = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples'), h.index('Thread Instructions Executed')
agg, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break
    try:
        n, s, t = int(r[ci]), int(r[smp]), int(r[ti])
    except Exception:
        continue
    op = re.sub(r'^@!?U?P\d+\s+', '', r[si].strip())
    base = (op.split()[0] if op else '?').split('.')[0]
    agg[base] += n; samp[base] += s; thr[base] += t; tot += n
print("total warp-instr", tot, "avg threads", sum(thr.values()) / max(tot, 1))
for k, v in agg.most_common(24):
    print(f"  {k:10s} {v / tot * 100:6.2f}%  avg-threads {thr[k] / max(v, 1):5.1f}  samples {samp[k]}")
