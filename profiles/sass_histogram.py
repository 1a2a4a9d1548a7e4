"""Opcode histogram of the shipped library per kernel (the SASS mnemonics that prove the Blackwell
paths: UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR =
tcgen05.commit, SYNCS = mbarrier).  Usage: python profiles/sass_histogram.py > profiles/r2_sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, 'neural-ode-ion-channels_b200', 'csrc', 'libikr_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
keys = ['UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'SYNCS', 'FFMA2', 'FFMA', 'DFMA', 'HMMA', 'F2FP']
hist, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r'\(.*', '', cur).replace('void ikr::', '')
        hist[cur] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        hist[cur]['total'] += 1
        for k in keys:
            if op == k or (k in ('UTCHMMA', 'LDTM', 'STTM', 'UBLKCP', 'UTCBAR', 'SYNCS', 'F2FP') and op.startswith(k)):
                hist[cur][k] += 1
                break
print('# SASS opcode histogram of `libikr_b200.so` (sm_100a), per kernel\n')
print('`cuobjdump -sass` of the shipped library, counted by `profiles/sass_histogram.py`. UTCHMMA = '
      '`tcgen05.mma`, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UBLKCP = `cp.async.bulk`, UTCBAR = '
      '`tcgen05.commit`, SYNCS = mbarrier operations, F2FP = fp32 -> bf16 / fp16 packs of the operand '
      'split. No `HMMA` (legacy `mma.sync`) anywhere.\n')
print('| kernel | instructions | ' + ' | '.join(keys) + ' |')
print('|---|---|' + '---|' * len(keys))
for name, c in hist.items():
    if c['total'] < 50:
        continue
    print('| `%s` | %d | %s |' % (name, c['total'], ' | '.join(str(c[k]) for k in keys)))
