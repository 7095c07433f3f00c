#!/usr/bin/env python3
"""Summarise an ncu report's source page by code region: share of warp instructions, average active
lanes, stall samples.  Usage: tools/ncu_regions.py report.ncu-rep [bytes_per_region]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; gran = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0x200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ia, isrc, ii, it, isamp = (hdr.index(x) for x in ["Address", "Source", "Instructions Executed", "Predicated-On Thread Instructions Executed", "# Samples"])
base = int(data[0][ia], 16)
tot_i = sum(int(r[ii]) for r in data); tot_t = sum(int(r[it]) for r in data); tot_s = sum(int(r[isamp]) for r in data)
print(f"warp inst {tot_i:.3e}  thread inst {tot_t:.3e}  avg active lanes {tot_t/tot_i:.2f}  samples {tot_s}")
seg = {}
for r in data:
    off = int(r[ia], 16) - base
    s = seg.setdefault(off // gran, [0, 0, 0, 0, 0])
    n = int(r[ii]); s[0] += n; s[1] += int(r[it]); s[2] += int(r[isamp])
    op = r[isrc].split()
    op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
    if op[:2] in ("DA", "DM", "DF", "DS"): s[3] += n
    if op.startswith("MUFU"): s[4] += n
for k in sorted(seg):
    s = seg[k]
    if s[0] > tot_i * 0.003:
        print(f"{k*gran:#7x} inst {100*s[0]/tot_i:5.1f}%  lanes {s[1]/max(s[0],1):4.1f}  samples {100*s[2]/tot_s:5.1f}%  fp64 {100*s[3]/tot_i:5.1f}%  mufu {100*s[4]/tot_i:4.2f}%")
if len(sys.argv) > 3:
    lo, hi = int(sys.argv[3], 0), int(sys.argv[4], 0)
    for r in data:
        off = int(r[ia], 16) - base
        if lo <= off < hi:
            print(f"{off:#7x} {int(r[ii]):>12} {int(r[it])/max(int(r[ii]),1):5.1f} {int(r[isamp]):>6}  {r[isrc].strip()}")
