#!/usr/bin/env python
"""Instruction distribution by CUDA source line from an ncu report: python tools/insn_dist.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
out = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]; ii = hdr.index("Instructions Executed"); ti = hdr.index("Thread Instructions Executed")
data = []
for r in rows[hi + 1:]:
    if len(r) <= ii or r[2] != "-":
        continue
    try:
        data.append((float(r[ii] or 0), float(r[ti] or 0), r[0], r[1]))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("# warp instructions executed by source line; total %.4g, mean active lanes %.1f" % (tot, sum(d[1] for d in data) / tot))
for v, t, ln, txt in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%5.1f%% lanes=%4.1f %5s %s" % (100 * v / tot, t / max(v, 1), ln, txt.strip()[:120]))
