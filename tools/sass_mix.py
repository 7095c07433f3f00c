#!/usr/bin/env python3
"""Static SASS instruction mix of every kernel in libfriq_b200.so (cuobjdump -sass): what the claims about code size, FP64 share,
TMA (UBLKCP) and MUFU use in DESIGN.md / profiles/README.md can be checked against without rebuilding.
Usage: tools/sass_mix.py [lib.so] > profiles/r2_sass_mix.md"""
import collections, os, re, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "fri-reinforcementlearning-c_b200", "libfriq_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
CLASSES = [("fp64", r"^(DADD|DMUL|DFMA|DSETP|DMNMX)"), ("mufu", r"^MUFU"), ("ldg", r"^(LDG|LD\.)"), ("stg", r"^(STG|ST\.)"),
           ("lds/sts", r"^(LDS|STS|LDSM)"), ("ldc", r"^(LDC|ULDC)"), ("local", r"^(LDL|STL)"), ("shfl/vote", r"^(SHFL|VOTE|MATCH|REDUX)"),
           ("atom", r"^(ATOM|RED|ATOMS|ATOMG)"), ("branch", r"^(BRA|BRX|CALL|RET|EXIT|BSSY|BSYNC|WARPSYNC|BREAK|JMP)"),
           ("tma", r"^(UBLKCP|UTMA|SYNCS)"), ("bar", r"^(BAR|MEMBAR|FENCE|ERRBAR)")]
kern = None; mix = collections.OrderedDict(); arch = None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m: arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); mix[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1); c = mix[kern]; c["total"] += 1
        for name, rx in CLASSES:
            if re.match(rx, op):
                c[name] += 1; break
        else:
            c["int/other"] += 1
print(f"# SASS instruction mix per kernel ({os.path.basename(so)}, arch {arch}; static counts, 16 bytes per instruction)\n")
cols = ["total", "KB"] + [c for c, _ in CLASSES] + ["int/other"]
print("| kernel | " + " | ".join(cols) + " |"); print("|---|" + "---|" * len(cols))
for k, c in sorted(mix.items(), key=lambda kv: -kv[1]["total"]):
    if c["total"] < 40: continue
    name = demangle(k); name = re.sub(r"\(const friq_devprob.*$|\(friq_devstate.*$", "", name).replace("void ", "").replace("(int)", "").replace("(bool)", "")
    print(f"| `{name}` | {c['total']} | {c['total'] * 16 / 1024:.1f} | " + " | ".join(str(c[x]) for x in cols[2:]) + " |")
