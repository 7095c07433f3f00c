#!/usr/bin/env python3
"""Scheduling control of a kernel's SASS, instruction by instruction: stall count, write / read scoreboard, wait mask.

  tools/sass_scoreboards.py <object-or-library> <mangled-kernel-substring> [lo-address hi-address] [--loads]

sm_100a instructions are 128 bits; cuobjdump prints the upper 64-bit word on a line of its own.  In that word bits 41-44 are
the stall count, 46-48 the scoreboard a variable-latency instruction signals when its RESULT is ready ("wr"), 49-51 the one it
signals when its SOURCES have been read ("rd"), 52-57 the mask of scoreboards this instruction waits for (7 = none).  A
scoreboard is a counter shared by everything that signals it, so a wait on it waits for ALL of them: loads that are meant to
overlap (the two halves of a double buffer) must sit on different scoreboards.  ptxas puts every load of a loop body that
contains control flow on ONE scoreboard; this tool is how that was found (profiles/README.md, "the scoreboard finding").
--loads prints only memory instructions, instructions that wait, and branches."""
import re
import subprocess
import sys


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only = "--loads" in sys.argv
    obj, kern = args[0], args[1]
    lo = int(args[2], 16) if len(args) > 2 else 0
    hi = int(args[3], 16) if len(args) > 3 else 1 << 30
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
    pat = re.compile(r"^\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
    pat2 = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")
    inside = False
    n = stalls = 0
    i = 0
    while i < len(sass) - 1:
        line = sass[i]
        if "Function :" in line:
            inside = kern in line
        m = pat.match(line) if inside else None
        if m:
            m2 = pat2.match(sass[i + 1])
            if m2:
                addr = int(m.group(1), 16); w = int(m2.group(1), 16)
                stall, wr, rd, wait = (w >> 41) & 0xf, (w >> 46) & 7, (w >> 49) & 7, (w >> 52) & 0x3f
                text = m.group(2)
                if lo <= addr <= hi:
                    n += 1; stalls += stall
                    if not only or wait or re.search(r"\b(LD|ST|LDG|STG|LDS|STS|LDL|STL|BRA|ATOM|RED)\b|LDG|STG|LDS|BRA", text):
                        print("%05x st=%2d wr=%s rd=%s wait=%s  %s" % (addr, stall, "-" if wr == 7 else wr, "-" if rd == 7 else rd,
                                                                     format(wait, "06b"), text[:100]))
                i += 2
                continue
        i += 1
    print("# %d instructions, static stall counts sum to %d cycles" % (n, stalls))


if __name__ == "__main__":
    main()
