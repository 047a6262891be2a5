#!/usr/bin/env python3
"""SASS instructions per source function / per 10-line bucket of one kernel.
usage: python tools/sass_funcs.py <lib.so> <kernel> <source file (basename)>"""
import collections, os, re, subprocess, sys, tempfile
lib, kern, srcname = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "tables" not in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines, cur, on = [], None, False
for line in sass.split("\n"):
    if re.match(r"\s*\.section\s+\.text\.", line): on = kern in line; cur = None; continue
    if re.match(r"\s*\.section\s", line): on = False; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line): lines.append(cur)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "mp3-enc-bsd_b200", "csrc", srcname)).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:SIMT_(?:FN|NOINLINE)|__device__ __forceinline__|template.*)\s+[\w:<> ]+?\s+\*?(\w+)\(", l)
    if m: funcs.append((i, m.group(1)))
def fn(ln):
    name = "?"
    for s, n in funcs:
        if s <= ln: name = n
    return name
c = collections.Counter(); b = collections.Counter()
for ln in lines:
    if ln and ln[0] == srcname: c[fn(ln[1])] += 1; b[ln[1] // 10 * 10] += 1
    else: c[ln[0] if ln else "?"] += 1
print(len(lines), "instructions = %.1f KB" % (len(lines) * 16 / 1024))
for k, v in c.most_common(16): print("  %-28s %5d  %.1f KB" % (k, v, v * 16 / 1024))
print("10-line buckets >= 40 instructions:")
for k in sorted(b):
    if b[k] >= 40: print("  %4d %5d  %s" % (k, b[k], src[k][:100]))
