#!/usr/bin/env python
"""Print the headline metrics of an .ncu-rep (first kernel) and, with --source, the hottest
source lines by warp-stall samples.  Usage: tools/ncu_report.py file.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread ", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ", "sass__inst_executed_local", "smsp__inst_executed.sum ",
        "smsp__average_warps_issue_stalled", "l1tex__throughput.avg.pct", "lts__throughput.avg.pct",
        "sm__inst_executed_pipe_lsu", "smsp__inst_executed_pipe_fp64", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")]
        print(f"## {name[:100]}")
        for h, u, v in zip(hdr, units, vals):
            hh = h + " "
            if any(k in hh for k in KEYS):
                try:
                    fv = float(v.replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in h and fv < 0.05:
                    continue
                print(f"  {h:92s} {v:>16s} {u}")


def source(path, top):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    data = []
    for r in rows:
        if "Source" in r and hdr is None:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(r)
    if not hdr:
        print(out[:2000])
        return
    si = hdr.index("Source")
    samp = [i for i, h in enumerate(hdr) if h.startswith("# Samples") or h == "Warp Stall Sampling (All Samples)" or "Samples" in h]
    ie = [i for i, h in enumerate(hdr) if h.startswith("Instructions Executed")]
    print("columns:", [hdr[i] for i in samp[:3]], [hdr[i] for i in ie[:2]])
    key = samp[0]

    def num(x):
        try:
            return float(x.replace(",", ""))
        except ValueError:
            return 0.0
    tot = sum(num(r[key]) for r in data) or 1
    data.sort(key=lambda r: -num(r[key]))
    for r in data[:top]:
        print(f"{100 * num(r[key]) / tot:6.2f}%  {r[ie[0]] if ie else '':>12s}  {r[si][:130]}")


if __name__ == "__main__":
    p = sys.argv[1]
    if "--source" in sys.argv:
        source(p, int(sys.argv[sys.argv.index("--source") + 1]))
    else:
        raw(p)
