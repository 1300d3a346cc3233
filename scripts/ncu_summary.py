#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into the few numbers DESIGN.md/profiles/ cite."""
import csv, io, re, subprocess, sys, collections

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum.per_second', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'launch__waves_per_multiprocessor']

def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main(rep, top=25):
    rows = page(rep, 'raw')
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('== kernel:', vals[hdr.index('Kernel Name')][:80])
        for i, h in enumerate(hdr):
            if h in WANT or 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct'):
                try:
                    v = float(vals[i])
                except ValueError:
                    continue
                if 'warp_issue_stalled' in h and v < 2.0:
                    continue
                print('  %-90s %14.3f %s' % (h, v, units[i]))
    src = page(rep, 'source')
    h = src[1]; data = src[2:]
    isrc, ie, ismp = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
    tot_i = sum(int(r[ie]) for r in data if r[ie].isdigit())
    tot_s = sum(int(r[ismp]) for r in data if r[ismp].isdigit())
    print('== source: %d SASS lines, %d warp-instructions, %d samples' % (len(data), tot_i, tot_s))
    ops = collections.Counter()
    for r in data:
        if r[ie].isdigit():
            m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
            if m: ops[m.group(2).split('.')[0]] += int(r[ie])
    print('  executed opcode mix:', ', '.join('%s %.1f%%' % (k, 100.0 * v / tot_i) for k, v in ops.most_common(14)))
    print('  top stall lines (samples, executed, SASS):')
    for r in sorted((r for r in data if r[ismp].isdigit()), key=lambda r: -int(r[ismp]))[:top]:
        print('   %6s %10s  %s' % (r[ismp], r[ie], r[isrc][:100]))

if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
