#!/usr/bin/env python3
"""SASS instruction count per source line bucket for one kernel (nvdisasm --print-line-info output)."""
import re, sys, collections
path, kern = sys.argv[1], sys.argv[2]
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cnt = collections.Counter(); cur = None; on = False
for line in open(path):
    if line.startswith('.text.'):
        on = kern in line
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and cur: cnt[cur] += 1
print('total', sum(cnt.values()))
b = collections.Counter()
for (f, l), c in cnt.items(): b[(f, l // bucket * bucket)] += c
for k, c in sorted(b.items()): 
    if c >= 20: print(k, c)
