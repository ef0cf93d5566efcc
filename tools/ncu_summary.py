#!/usr/bin/env python3
"""Summarise an .ncu-rep (read offline with `ncu -i`): per-launch key metrics and stall mix."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.avg',
        'smsp__warps_eligible.avg.per_cycle_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']


def main():
    rep = sys.argv[1]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[ix['Kernel Name']]
        short = name.split('(')[0][-70:]
        if '--all' not in sys.argv and short in seen:
            continue
        seen.add(short)
        print('=====', name[:160])
        for k in KEYS:
            if k in ix:
                print(f'  {k:75s} {r[ix[k]]:>16s} {units[ix[k]]}')
        st = []
        for h, i in ix.items():
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    st.append((float(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        print('  stalls/issue: ' + ', '.join(f'{n}={v:.2f}' for v, n in sorted(st, reverse=True)[:8]))


if __name__ == '__main__':
    main()
