#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libmp3gpu.so — instruction count, registers / stack, and the histogram of the
opcodes that matter for the Blackwell story (DMMA = FP64 tensor cores, UBLKCP / SYNCS = bulk asynchronous copy + mbarrier,
DFMA / DMUL / DADD, FFMA, LDS / STS, LDCU = uniform constant loads, SHFL, REDUX, MUFU, BAR).  CPU only."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mp3-enc-bsd_b200", "libmp3gpu.so")
KEY = ["DMMA", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "FFMA", "FMUL", "FADD", "LDS", "STS", "LDG", "STG", "LDC", "LDCU", "SHFL",
       "REDUX", "MUFU", "BAR", "MEMBAR", "CCTL", "ATOMG", "ATOMS", "NANOSLEEP", "I2F", "F2F", "LDL", "STL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = m.groups()
    hist, order = {}, []
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur]["_total"] += 1
            hist[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
    for name, pretty in zip(order, demangle):
        h = hist[name]
        reg = usage.get(name, ("?", "?", "?"))
        print("== %s" % pretty.split("(")[0])
        print("   %d SASS instructions, %s registers, %s B stack" % (h["_total"], reg[0], reg[1]))
        print("   " + "  ".join("%s %d" % (k, h[k]) for k in KEY if h[k]))


if __name__ == "__main__":
    main()
