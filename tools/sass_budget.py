"""Static per-loop instruction budget of one kernel from its SASS (no GPU needed):

    python tools/sass_budget.py beer_b200/build/scan.o 'hmm_fb_lr_kernelILi4ELi1ELb0'

Loops are found from backward branches (BRA to a lower address); for every innermost loop body the
opcodes are counted by class.  A body that holds conditional code (the unit counts, the optional
outputs) is an upper bound of what one iteration issues; the dynamic count (ncu smsp__inst_executed)
is the cross-check."""
import collections
import re
import subprocess
import sys

CLASSES = [
    ('mufu', r'^MUFU'),
    ('redux/shfl', r'^(REDUX|SHFL|CREDUX)'),
    ('fp32', r'^(FADD|FMUL|FFMA|FMNMX|FSEL|FSETP|FSET|FMNMX3)'),
    ('fp64', r'^(DADD|DMUL|DFMA|F2F|DSETP)'),
    ('int/logic', r'^(IADD|IADD3|IMAD|LEA|LOP3|SHF|ISETP|SEL|MOV|IMNMX|VIADD|VIMNMX|PLOP3|PRMT|I2F|F2I|S2R|CS2R|UMOV|ULEA|UIADD3|UISETP|ULOP3|USHF|USEL|UIMAD|R2UR|UPLOP3|S2UR|R2P|P2R|NOP|UFLO|VOTE|VOTEU|POPC|BREV|FLO|I2FP)'),
    ('shared ld/st', r'^(LDS|STS|LDSM)'),
    ('global ld/st', r'^(LDG|STG|LDGSTS|LDGDEPBAR|DEPBAR|ATOM|RED|ATOMG|LD|ST|LDC|ULDC|LDCU|CCTL|MEMBAR|ERRBAR|FENCE)'),
    ('control', r'^(BRA|BSSY|BSYNC|EXIT|WARPSYNC|BAR|CALL|RET|BRX|JMP|YIELD|NANOSLEEP|BMOV|BREAK|BPT)'),
]


def classify(op):
    for name, pat in CLASSES:
        if re.match(pat, op):
            return name
    return 'other:' + op.split('.')[0]


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    sass = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True, check=True).stdout
    instrs, inside = [], False
    for line in sass.splitlines():
        if 'Function :' in line:
            inside = pattern in line
            if inside:
                print(line.strip())
            continue
        if not inside:
            continue
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
        if m:
            text = m.group(2).strip()
            pred = re.match(r'^@!?U?P\w+\s+', text)
            body = text[pred.end():] if pred else text
            instrs.append((int(m.group(1), 16), body, bool(pred)))
    addr_index = {a: i for i, (a, _, _) in enumerate(instrs)}
    loops = []
    for i, (a, text, _) in enumerate(instrs):
        m = re.match(r'^BRA(?:\.\w+)*\s+(?:\w+,\s*)?`?\(?\.?L?_?x?_?\w*\)?', text)
        if text.startswith('BRA'):
            t = re.search(r'0x([0-9a-f]+)', text)
            if t and int(t.group(1), 16) <= a and int(t.group(1), 16) in addr_index:
                loops.append((addr_index[int(t.group(1), 16)], i))
    # innermost loops only
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    print(f'{len(instrs)} instructions, {len(loops)} loops, {len(inner)} innermost')
    for lo, hi in sorted(set(inner)):
        n = hi - lo + 1
        if n < 24:
            continue
        hist = collections.Counter(classify(t.split()[0]) for _, t, _ in instrs[lo:hi + 1])
        npred = sum(1 for _, _, p in instrs[lo:hi + 1] if p)
        mufu = collections.Counter(t.split()[0] for _, t, _ in instrs[lo:hi + 1] if t.startswith('MUFU'))
        print(f'loop 0x{instrs[lo][0]:x}..0x{instrs[hi][0]:x}: {n} instructions ({npred} predicated)')
        for k, v in hist.most_common():
            print(f'    {k:14s} {v}')
        if mufu:
            print('    ' + ', '.join(f'{k} {v}' for k, v in mufu.items()))


if __name__ == '__main__':
    main()
