#!/usr/bin/env python
"""Opcode histogram (executed warp-instructions and stall samples) of one kernel of an ncu report.

    python tools/ncu_ops.py report.ncu-rep <kernel regex> [per-unit divisor]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, kernel, div=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', f'regex:{kernel}'],
                         capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    seen, ops, samples, tot = set(), collections.Counter(), collections.Counter(), 0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[0] == 'Address' or r[0] in seen:
            continue
        seen.add(r[0])
        try:
            n, sm = int(r[ix['Instructions Executed']]), int(r[ix['# Samples']])
        except ValueError:
            continue
        parts = r[ix['Source']].split()
        op = parts[1] if parts[0].startswith('@') else parts[0]
        op = '.'.join(op.split('.')[:2]) if op.startswith(('MUFU', 'LDS', 'STS', 'LDG', 'STG', 'SYNCS')) else op.split('.')[0]
        ops[op] += n
        samples[op] += sm
        tot += n
    print(f'total warp-instructions {tot}' + (f' = {tot / div:.1f} per unit' if div else ''))
    for op, n in ops.most_common(28):
        extra = f'  {n / div:7.2f} / unit' if div else ''
        print(f'{op:14s} {100 * n / tot:5.1f}%{extra}  samples {samples[op]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
