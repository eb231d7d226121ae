#!/usr/bin/env python
"""Summarise an `ncu --set full` report as a markdown table (one row per captured launch).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor %'),
    ('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'xu %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu %'),
    ('launch__registers_per_thread', 'regs'),
    ('smsp__inst_executed.sum', 'warp-instr'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem conflicts'),
    ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
    ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_sb'),
    ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
    ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
    ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS if m in ix]
    print('| kernel | ' + ' | '.join(f'{n} ({units[ix[m]]})' if units[ix[m]] else n for m, n in cols) + ' |')
    print('|---|' + '---|' * len(cols))
    for r in rows[2:]:
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '')[:48]
        vals = []
        for m, _ in cols:
            v = r[ix[m]]
            try:
                f = float(v.replace(',', ''))
                v = f'{f:.4g}'
            except ValueError:
                pass
            vals.append(v)
        print(f'| {name} | ' + ' | '.join(vals) + ' |')


if __name__ == '__main__':
    main(sys.argv[1])
