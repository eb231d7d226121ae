#!/usr/bin/env python
"""Hottest SASS instructions of one kernel of an ncu report (per-instruction warp-stall samples).

    python tools/ncu_hot.py report.ncu-rep <kernel regex> [top N] [view: sass|source]
"""
import csv
import io
import subprocess
import sys


def main(path, kernel, top=40, view='sass'):
    cmd = ['ncu', '-i', path, '--page', 'source', '--csv', '--kernel-name', f'regex:{kernel}']
    if view == 'source':
        cmd += ['--print-source', 'cuda,sass']
    raw = subprocess.run(cmd, capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    rows = list(csv.reader(io.StringIO('\n'.join(lines[start:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[0].startswith('Kernel') or r[0] == 'Address':
            continue
        try:
            n = int(r[ix['# Samples']])
        except ValueError:
            continue
        data.append((n, r))
    total = sum(n for n, _ in data) or 1
    tot_inst = sum(int(r[ix['Instructions Executed']] or 0) for _, r in data)
    print(f'total samples {total}, instructions executed {tot_inst}')
    agg = {s: 0 for s in stalls}
    for n, r in data:
        for s in stalls:
            try:
                agg[s] += int(r[ix[s]])
            except ValueError:
                pass
    print('stall mix:', ', '.join(f'{s[6:]} {100 * v / total:.1f}%' for s, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for n, r in sorted(data, key=lambda t: -t[0])[:top]:
        why = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
        print(f'{100 * n / total:5.1f}%  {r[ix["Instructions Executed"]]:>10}  {r[ix["Source"]][:90]:<90}  '
              + ' '.join(f'{w}:{c}' for c, w in why if c))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else 'sass')
