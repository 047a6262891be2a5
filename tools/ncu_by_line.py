#!/usr/bin/env python3
"""Attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<kernel> > k.csv
    cuobjdump -xelf all libmp3gpu.so; nvdisasm --print-line-info mp3gpu.sm_100a.cubin > all.sass
    python tools/ncu_by_line.py k.csv all.sass <kernel> [bucket]

The n-th SASS instruction of the kernel in the ncu page is the n-th instruction of the kernel's .text section in the
nvdisasm listing (same binary), whose '//## File ..., line N' markers give the source line (inlined callee lines are
reported at their own file:line).  Prints per line bucket: stall samples, instructions executed, shared wavefronts
(and the excessive part), global L1 tag requests, L2 sectors."""
import collections
import csv
import re
import sys


def main():
    ncu_csv, sass, kern = sys.argv[1:4]
    bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    lines, cur, on = [], None, False
    for line in open(sass, errors="ignore"):
        if line.startswith("\t.section\t.text.") or line.startswith(".section\t.text.") or re.match(r"\s*\.section\s+\.text\.", line):
            on = kern in line
            cur = None
            continue
        if re.match(r"\s*\.section\s", line):
            on = False
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
            lines.append(cur)
    rows = list(csv.reader(open(ncu_csv, errors="ignore")))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    col = {n: hdr.index(n) for n in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive",
                                     "L1 Tag Requests Global", "L2 Theoretical Sectors Global")}
    body = []
    for r in rows[h + 1:]:
        if r and r[0] == "Kernel Name":      # the page repeats per captured instance: keep the first
            break
        if len(r) == len(hdr):
            body.append(r)
    if len(body) != len(lines):
        print("warning: %d SASS rows in the ncu page, %d instructions in the listing" % (len(body), len(lines)), file=sys.stderr)
    agg = collections.defaultdict(lambda: [0] * len(col))
    for r, ln in zip(body, lines):
        key = (ln[0], ln[1] // bucket * bucket) if ln else ("?", 0)
        for k, (n, c) in enumerate(col.items()):
            try:
                agg[key][k] += int(float(r[c]))
            except ValueError:
                pass
    tot = [sum(v[k] for v in agg.values()) for k in range(len(col))]
    print("%-28s %9s %12s %12s %12s %12s %12s" % ("file:line", "samples", "inst", "smem_wave", "smem_excess", "gl_tag_req", "l2_sectors"))
    print("%-28s %9d %12d %12d %12d %12d %12d" % (("TOTAL",) + tuple(tot)))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(__import__("os").environ.get("TOPN","45"))]:
        print("%-28s %9d %12d %12d %12d %12d %12d" % (("%s:%d" % key,) + tuple(v)))


if __name__ == "__main__":
    main()
