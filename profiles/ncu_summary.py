#!/usr/bin/env python
"""Summarise an ncu report: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [launches.csv]"""
import csv, io, re, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__occupancy_limit_barriers', 'launch__occupancy_limit_blocks',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__shared_mem_config_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print('=' * 100)
        print(re.sub(r'\(.*', '', r[idx['Kernel Name']]))
        for w in WANT:
            if w in idx:
                print('  %-82s %s %s' % (w, r[idx[w]], units[idx[w]]))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2])))
        for i, r in enumerate(rows):
            if 'Kernel Name' in r:
                hdr = r
                start = i + 1
                break
        ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
        print('=' * 100)
        print('launch list (gpu__time_duration.sum, ns; cold-cache, serialised)')
        for r in rows[start:]:
            if len(r) > vi:
                print('  %4s %-90s %s' % (r[0], re.sub(r'\(.*', '', r[ki])[:90], r[vi]))


if __name__ == '__main__':
    main()
