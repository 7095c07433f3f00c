#!/usr/bin/env python3
"""Per-source-line stall breakdown of one profiled launch (ncu --set full --import-source on): where the long-scoreboard, wait,
no-instruction ... samples sit.  Usage: tools/ncu_stalls.py report.ncu-rep lib.so <mangled-kernel-substring> [stall] [top-N]"""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
stall = sys.argv[4] if len(sys.argv) > 4 else "stall_long_sb"
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "engine" in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
line_of = {}; cur = None; inside = False
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kern in l; continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if r and re.match(r"^0x[0-9a-f]+$", r[0])]
ia, isrc, ist, iall = hdr.index("Address"), hdr.index("Source"), hdr.index(stall), hdr.index("# Samples")
base = int(data[0][ia], 16)
agg = collections.Counter(); ops = collections.defaultdict(collections.Counter); tot = 0; allsamp = 0
for r in data:
    k = line_of.get(int(r[ia], 16) - base, ("?", 0)); v = int(r[ist]); agg[k] += v; tot += v; allsamp += int(r[iall])
    if v: ops[k][r[isrc].split()[1] if r[isrc].split()[0].startswith("@") else r[isrc].split()[0]] += v
print(f"{stall}: {tot} of {allsamp} samples ({100.0 * tot / max(allsamp, 1):.1f} %)")
for k, v in agg.most_common(topn):
    print(f"{k[0]:>20}:{k[1]:<4} {100.0 * v / max(tot, 1):5.1f}%   {dict(ops[k].most_common(3))}")
