"""Aggregate an ncu source-page CSV (SASS view) by CUDA source line using nvdisasm line info.

usage: ncu_by_line.py <sass.csv from `ncu -i rep --page source --csv`> <nvdisasm -g -c output> <kernel substring> [top]
"""
import collections
import csv
import re
import sys

csv_path, sass_path, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
line_of, cur, inside, offs = {}, None, False, []
for ln in open(sass_path):
    if ln.startswith(".text."):
        inside = kernel in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csv_path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
base = None
by_line = collections.Counter()
by_line_ops = collections.defaultdict(collections.Counter)
stall = collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    addr = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None:
        base = addr
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    key = line_of.get(addr - base, ("?", 0))
    by_line[key] += n
    stall[key] += s
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    by_line_ops[key][m.group(2) if m else "?"] += n
    tot += n
stot = sum(stall.values())
print(f"total warp instructions {tot:.4g}, samples {stot:.0f}")
for key, n in by_line.most_common(top):
    ops = ", ".join(f"{k} {100 * v / n:.0f}%" for k, v in by_line_ops[key].most_common(4))
    print(f"{key[0]:22s}:{key[1]:<5d} inst {100 * n / tot:5.1f}%  samples {100 * stall[key] / max(stot, 1):5.1f}%   {ops}")
