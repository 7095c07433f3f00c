#!/usr/bin/env python3
"""Aggregate an ncu report's per-SASS-instruction counters by CUDA source line.
Usage: tools/ncu_lines.py report.ncu-rep libfriq_b200.so <mangled-kernel-substring> [launch-index] [top-N]
The .so must be the build that was profiled (compiled with -lineinfo); nvdisasm supplies the address -> line map."""
import csv, io, os, re, subprocess, sys, tempfile, collections
rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
launch = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "engine" in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
line_of = {}
cur = None; inside = False
for l in dis.splitlines():
    if l.startswith(".text."):
        inside = kern in l
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the csv holds one table per profiled launch, each starting with a "Kernel Name"-less header row beginning with "Address"
tables = []; rows = list(csv.reader(io.StringIO(out))); name = ""
want = kern.split("ILi")[0].lstrip("_Z0123456789") + "<"          # demangled prefix, e.g. "k_run_steps<"
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
    elif r and r[0] == "Address":
        tables.append([r] if want in name else None)
    elif tables and tables[-1] is not None and r and re.match(r"^[0-9a-fx]+$", r[0]):
        tables[-1].append(r)
tables = [t for t in tables if t]
t = tables[launch]
hdr = t[0]; data = t[1:]
ia, isrc, ii, it, isamp = (hdr.index(x) for x in ["Address", "Source", "Instructions Executed", "Predicated-On Thread Instructions Executed", "# Samples"])
base = int(data[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for r in data:
    off = int(r[ia], 16) - base
    k = line_of.get(off, ("?", 0))
    n, th, sm = int(r[ii]), int(r[it]), int(r[isamp])
    a = agg[k]; a[0] += n; a[1] += th; a[2] += sm
    op = r[isrc].split(); op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
    if op[:2] in ("DA", "DM", "DF", "DS"): a[3] += n
    tot[0] += n; tot[1] += th; tot[2] += sm
print(f"launch {launch}: warp inst {tot[0]:.3e}, avg lanes {tot[1]/max(tot[0],1):.1f}, samples {tot[2]}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{k[0]:>18}:{k[1]:<4} inst {100*a[0]/tot[0]:5.1f}%  samples {100*a[2]/max(tot[2],1):5.1f}%  lanes {a[1]/max(a[0],1):4.1f}  fp64 {100*a[3]/max(a[0],1):3.0f}%")
