#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that tell a tcgen05 / TMA kernel from a legacy one (B200_PROFILING.md: UTC*MMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, HMMA = mma.sync) in lib/libptta_b200.so.
Runs without a GPU:   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tta_depth_completion_b200', 'lib', 'libptta_b200.so')
PAT = [('UTCHMMA', r'\bUTC\w*MMA\b'), ('LDTM', r'\bLDTM\b'), ('STTM', r'\bSTTM\b'), ('UTMALDG', r'\bUTMALDG\b'), ('UTMASTG', r'\bUTMASTG\b'),
       ('UBLKCP', r'\bUBLKCP\b'), ('HMMA', r'\bHMMA\b'), ('LDGSTS', r'\bLDGSTS\b'), ('FFMA', r'\bFFMA\b')]
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
counts, order, cur = collections.defaultdict(lambda: collections.Counter()), [], None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', cur).replace('void ', '').replace('ptta::', '')
        order.append(cur)
        continue
    if cur is None:
        continue
    for name, pat in PAT:
        if re.search(pat, line):
            counts[cur][name] += 1
print('%-52s' % ('kernel (%s)' % os.path.basename(LIB)) + ''.join('%9s' % n for n, _ in PAT))
for k in sorted(order, key=lambda k: (-counts[k]['UTCHMMA'], -counts[k]['HMMA'], k)):
    c = counts[k]
    if not (c['UTCHMMA'] or c['HMMA'] or c['UTMALDG'] or c['UTMASTG'] or c['LDTM'] or c['UBLKCP']):
        continue
    print('%-52s' % k[:52] + ''.join('%9d' % c[n] for n, _ in PAT))
tc = [k for k in order if counts[k]['UTCHMMA']]
legacy = [k for k in order if counts[k]['HMMA'] and not counts[k]['UTCHMMA']]
print('\n%d kernels issue tcgen05.mma (UTC*MMA); %d kernels use the legacy mma.sync path (HMMA) only: %s' % (len(tc), len(legacy), ', '.join(sorted(set(re.sub(r'<.*', '', k) for k in legacy)))))
