"""SASS evidence for the hot kernels:  python tools/sass_summary.py > profiles/<tag>_sass_summary.txt
Per kernel of qcmrf_b200/libqcmrf_b200.so (sm_100a cubin): instruction count and a histogram of the mnemonics that
matter on this path -- 128-bit global loads/stores (LDG/STG.E.*128), shared-memory traffic (LDS/STS), warp shuffles
(SHFL), the TMA bulk-copy engine (UBLKCP) and mbarrier synchronisation (SYNCS), system-scope flag accesses -- followed
by the full listing of the kernels named in FULL."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'qcmrf_b200', 'libqcmrf_b200.so')
HOT = ['k_expand_low', 'k_lowq', 'k_block<', 'k_block_gather_tma', 'k_block_gather_inplace', 'k_expand_tree', 'k_init<',
       'k_diag_multi', 'k_sample', 'k_chunk_sums', 'k_mrf']
FULL = ['k_lowq<float, 2, 3, 0>', 'k_expand_low<float, 2, 2, 8>']
KEYS = [('LDG.*128', r'^LDG\..*128'), ('STG.*128', r'^STG\..*128'), ('LDG other', r'^LDG'), ('STG other', r'^STG'),
        ('LDS', r'^LDS'), ('STS', r'^STS'), ('SHFL', r'^SHFL'), ('UBLKCP (TMA bulk copy)', r'^UBLKCP'), ('SYNCS (mbarrier)', r'^SYNCS'),
        ('BAR', r'^BAR'), ('FFMA/FMUL/FADD', r'^(FFMA|FMUL|FADD)'), ('DFMA/DMUL/DADD', r'^(DFMA|DMUL|DADD)'),
        ('LD/ST .SYS (flags)', r'^(LD|ST|LDG|STG)\..*SYS'), ('NANOSLEEP', r'^NANOSLEEP'), ('LDL/STL (spills)', r'^(LDL|STL)')]


def _short(name):
    s = re.sub(r'^void ', '', name)
    s = re.sub(r'qcm::', '', s)
    s = re.sub(r'\((int|bool)\)', '', s)
    return s


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    dem = subprocess.run(['cu++filt'], input=sass, capture_output=True, text=True).stdout or sass
    kernels, name, body = [], None, []
    for line in dem.splitlines():
        m = re.match(r'\s*Function : (.*)', line)
        if m:
            if name:
                kernels.append((name, body))
            name, body = m.group(1).strip(), []
        elif name is not None:
            body.append(line)
    if name:
        kernels.append((name, body))
    print('# cuobjdump -sass %s | cu++filt  (sm_100a), kernels matching %s' % (os.path.relpath(LIB, ROOT), HOT))
    for name, body in kernels:
        short = _short(name)
        if not any(h in short for h in HOT):
            continue
        ops = []
        for l in body:
            m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
            if m:
                ops.append(m.group(1))
        hist = collections.OrderedDict()
        for label, pat in KEYS:
            hist[label] = 0
        for o in ops:
            for label, pat in KEYS:
                if re.match(pat, o):
                    hist[label] += 1
                    break
        print('\n%s\n    %d instructions; %s' % (short.split('(')[0], len(ops), ', '.join('%s %d' % (k, v) for k, v in hist.items() if v)))
    for name, body in kernels:
        short = _short(name)
        if any(short.startswith(f) for f in FULL):
            print('\n\n======== full listing: %s ========' % short.split('(')[0])
            for l in body:
                m = re.match(r'\s*/\*([0-9a-f]{4})\*/\s+(.*?)\s*;?\s*/\*', l)
                if m:
                    print('%s  %s' % (m.group(1), m.group(2)))


if __name__ == '__main__':
    main()
