#!/usr/bin/env python
"""Per-kernel census of the Blackwell-native SASS opcodes in cc_b200/libccb200.so (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKCP = bulk copy,
FFMA2 = packed fp32 FMA.   python tools/sass_census.py > profiles/rNN_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'cc_b200', 'libccb200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, check=True).stdout.decode()
WANT = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'SYNCS', 'FFMA2', 'FFMA', 'HMMA', 'LDGSTS']
cur, counts, arch = None, collections.OrderedDict(), set()
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE).stdout.decode().strip()
        cur = re.sub(r'\(.*$', '', cur.replace('void ', ''))
        counts[cur] = collections.Counter()
        continue
    m = re.search(r'arch = (sm_\w+)', line)
    if m:
        arch.add(m.group(1))
    m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur]['_total'] += 1
        for w in WANT:
            if op == w or (w.endswith('MMA') and op.startswith(w)) or (w in ('UTMALDG', 'UTMASTG', 'LDTM', 'STTM') and op.startswith(w)):
                counts[cur][w] += 1
                break
print('# SASS census of %s (arch %s): instruction counts per kernel' % (os.path.relpath(so, ROOT), ','.join(sorted(arch))))
print('# %-58s %7s ' % ('kernel', 'total') + ' '.join('%8s' % w for w in WANT))
tot = collections.Counter()
for k, c in counts.items():
    print('%-60s %7d ' % (k[:60], c['_total']) + ' '.join('%8s' % (c[w] or '.') for w in WANT))
    tot.update(c)
print('%-60s %7d ' % ('TOTAL (%d kernels)' % len(counts), tot['_total']) + ' '.join('%8d' % tot[w] for w in WANT))
