#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libbeer_b200.so (the Blackwell-native mnemonics), as a markdown table:

    python tools/sass_histogram.py > profiles/rNN_sass_histogram.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDGSTS', 'SYNCS', 'ELECT', 'R2UR', 'CREDUX', 'MUFU']


def main():
    lib = os.path.join(ROOT, 'beer_b200', 'lib', 'libbeer_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True,
                           text=True).stdout.splitlines()
    hist, order, cur, k = {}, [], None, 0
    for line in sass.splitlines():
        if 'Function :' in line:
            cur = re.sub(r'\(.*', '', names[k]).replace('void ', '').replace('beer::', '').replace('(anonymous namespace)::', '')
            k += 1
            hist[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)', line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    print('# SASS opcode histogram of libbeer_b200.so (cuobjdump -sass, per kernel; Blackwell-native mnemonics)\n')
    print('UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG / '
          'UTMASTG =\ncp.async.bulk.tensor (tensor-map TMA load / store), LDGSTS = cp.async, SYNCS = mbarrier, ELECT = elect.sync, '
          'CREDUX = redux.sync.\nOnly kernels with at least one tensor-core or TMA instruction are listed.\n')
    print('| kernel | ' + ' | '.join(COLS) + ' |')
    print('|---|' + '---|' * len(COLS))
    for name in sorted(order):
        h = hist[name]
        if not any(h[c] for c in ('UTCHMMA', 'UBLKCP', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM')):
            continue
        print(f'| {name} | ' + ' | '.join(str(h[c]) for c in COLS) + ' |')


if __name__ == '__main__':
    main()
