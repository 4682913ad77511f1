#!/usr/bin/env python
"""Digest an .ncu-rep (read here, no GPU): per-kernel headline metrics, stall mix and the hottest source lines.
usage: python tools/ncu_digest.py gpurun_out/x.ncu-rep [--lines 25] [--csv out.csv]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]
out_rows = []
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=====", name[:90])
    rec = {"kernel": name}
    for k in KEYS:
        if k in hdr:
            print(f"  {k:85s} {r[hdr.index(k)]} {rows[1][hdr.index(k)]}")
            rec[k] = r[hdr.index(k)]
    st = [(h, float(r[i])) for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h and r[i]]
    for h, v in sorted(st, key=lambda t: -t[1])[:9]:
        short = h.split("issue_stalled_")[1].split("_per_issue")[0]
        print(f"  stall {short:28s} {v:6.2f}")
        rec["stall_" + short] = v
    out_rows.append(rec)
if "--csv" in sys.argv:
    path = sys.argv[sys.argv.index("--csv") + 1]
    keys = sorted({k for r in out_rows for k in r})
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=["kernel"] + [k for k in keys if k != "kernel"])
        w.writeheader()
        w.writerows(out_rows)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
func, file, h2, agg = None, None, None, {}
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        func = r[1].split("(")[0][-40:]
        continue
    if r[0] == "Line No":
        h2 = r
        continue
    if h2 is None or not r[0].isdigit():
        continue
    d = dict(zip(h2, r))
    f = lambda k: float(d[k]) if d.get(k, "").replace(".", "").isdigit() else 0.0
    a = agg.setdefault((func, file, int(r[0])), [0, 0, 0, 0, r[1][:100]])
    a[0] += f("Instructions Executed"); a[1] += f("# Samples"); a[2] += f("L1 Wavefronts Shared"); a[3] += f("L1 Tag Requests Global")
for fn in sorted({k[0] for k in agg}):
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    ti, ts = sum(v[0] for _, v in items) or 1, sum(v[1] for _, v in items) or 1
    tw, tg = sum(v[2] for _, v in items) or 1, sum(v[3] for _, v in items) or 1
    print(f"--- {fn}: hottest lines by stall samples")
    for k, v in sorted(items, key=lambda kv: -kv[1][1])[:nlines]:
        print(f"  {k[1][:22]:22s}:{k[2]:4d} inst {100*v[0]/ti:5.1f}% samp {100*v[1]/ts:5.1f}% smem {100*v[2]/tw:5.1f}% gtag {100*v[3]/tg:5.1f}% | {v[4]}")
