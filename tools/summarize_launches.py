#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total ms, share
(only the kernels of libmp3gpu, i.e. names starting with k_)."""
import csv, sys, collections
def main(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 14 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].split("::")[-1].replace("void ", "")
        if not name.startswith("k_"):
            continue
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1; t[1] += float(r[14]) * 1e-6
    s = sum(v[1] for v in tot.values()) or 1
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (n, ms) in tot.items():
        print(f"| {k} | {n} | {ms:.3f} | {100*ms/s:.1f} % |")
if __name__ == "__main__":
    main(sys.argv[1])
