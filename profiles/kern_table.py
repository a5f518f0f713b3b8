#!/usr/bin/env python
"""One line per kernel launch of an ncu report: duration, warp instructions, issue-active %, DRAM MB, top stalls.
usage: python profiles/kern_table.py rep.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
def col(name):
    return h.index(name) if name in h else None
ki = col("Kernel Name")
want = [("gpu__time_duration.sum", "us"), ("smsp__inst_executed.sum", "Minst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("launch__registers_per_thread", "regs")]
stalls = [c for c in h if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio")]
units = rows[1]
print(f"{'kernel':44s} " + " ".join(f"{n:>8s}" for _, n in want) + "  top stalls (cycles per issued instruction)")
for r in rows[2:]:
    vals = []
    for c, n in want:
        i = col(c)
        try:
            v = float(r[i].replace(",", ""))
            u = units[i]
            if n == "us": v = v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
            if n == "Minst": v /= 1e6
            if n in ("rdMB", "wrMB"): v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            vals.append(f"{v:8.1f}")
        except Exception:
            vals.append(f"{'-':>8s}")
    st = []
    for c in stalls:
        try:
            st.append((float(r[h.index(c)]), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except Exception:
            pass
    st = sorted(st, reverse=True)[:4]
    print(f"{r[ki][:44]:44s} " + " ".join(vals) + "  " + ", ".join(f"{n} {v:.1f}" for v, n in st if n != "selected"))
