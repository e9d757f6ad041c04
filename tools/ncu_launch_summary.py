#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
time and share (cold-cache, serialised launches: compare SHARES, not absolutes)."""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", name)
    m = re.match(r"(?:void\s+)?([\w:]+(?:<[^(]{0,60})?)", name)
    return (m.group(1) if m else name)[:90]


def main(path, only_ours=False):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else (v if u.startswith("m") else v * 1e3))
        k = short(r[ki])
        agg[k][0] += 1
        agg[k][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.3f} ms total (serialised, cold cache)")
    print(f"{'ms':>12} {'launches':>9} {'share':>7}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"{v[1]:12.3f} {v[0]:9d} {100 * v[1] / tot:6.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
