#!/usr/bin/env python
"""Per-kernel counts of the SASS opcodes that prove the Blackwell path (B200_PROFILING.md: UTCHMMA/UTCQMMA = tcgen05.mma,
UTMALDG/UTMASTG = TMA, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, SYNCS = mbarrier) from
`cuobjdump -sass libmudiff_b200.so`.   python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'mu-diff_b200', 'libmudiff_b200.so')
OPS = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'LDTM', 'STTM', 'UTCCP', 'SYNCS', 'HMMA', 'FFMA', 'FFMA2', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'SHFL']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = demangle(m.group(1))
        counts[cur] = collections.Counter()
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        counts[cur][op] += 1
        counts[cur]['_total'] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (sm_100a)  -  opcode counts per kernel")
print("# " + " ".join(f"{o:>8s}" for o in ['total'] + OPS) + "  kernel")
for k, c in sorted(counts.items(), key=lambda kv: -(kv[1]['UTCHMMA'] * 1000 + kv[1]['UTMALDG'])):
    name = re.sub(r'\s+', ' ', k)
    print("  " + " ".join(f"{c[o]:8d}" for o in ['_total'] + OPS) + "  " + (name[:150]))
