#!/usr/bin/env python
"""Print the warp-stall breakdown (pc sampling) and pipe utilisations of an .ncu-rep."""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, v = rows[0], rows[1], rows[2]
st = []
for i, n in enumerate(h):
    if 'pcsamp_warps_issue_stalled' in n and 'not_issued' not in n:
        try: st.append((float(v[i]), n.replace('smsp__pcsamp_warps_issue_stalled_', '')))
        except ValueError: pass
tot = sum(x for x, _ in st)
print('stalls:', ', '.join('%s %.1f%%' % (n, 100 * x / tot) for x, n in sorted(st, reverse=True)[:10]))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
          'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum',
          'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'launch__registers_per_thread',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
          'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct']:
    if k in h: print('  %-80s %s %s' % (k, v[h.index(k)], u[h.index(k)]))
