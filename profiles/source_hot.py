#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples, from an ncu report.
usage: python profiles/source_hot.py rep.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = {}
cur = None
for r in rows:
    if len(r) > 5 and r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != '':
        cur = (int(r[0]), r[1].strip()[:110])
        agg.setdefault(cur, [0, 0])
        continue
    try:
        agg[cur][0] += int(r[hdr.index('Instructions Executed')])
        agg[cur][1] += int(r[hdr.index('# Samples')])
    except Exception:
        pass
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instructions {ti}  samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("BYINST","0")) ^ 1])[:top]:
    print(f"{k[0]:5d} inst {100*v[0]/max(ti,1):5.1f}%  samples {100*v[1]/max(ts,1):5.1f}%  {k[1]}")
