"""Per-kernel SASS opcode summary of boom_b200/libboomgpu.so (run here, no GPU): which kernels carry FP64 tensor instructions
(DMMA), TMA tensor loads (UTMALDG), bulk copies (UBLKCP), mbarrier traffic (SYNCS), setmaxnreg (USETMAXREG) ...
    python profiles/sass_summary.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "boom_b200", "libboomgpu.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ("DMMA", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "ATOM", "RED",
       "MEMBAR", "FENCE", "BAR", "IMAD", "LOP3", "FFMA", "NOP")
name, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        counts[name] = collections.Counter()
        order.append(name)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        counts[name][m.group(1)] += 1
        counts[name]["_total"] += 1
print("# SASS opcode counts per kernel (static), sm_100a, %s" % os.path.relpath(lib, ROOT))
print("# no UTC*MMA anywhere: tcgen05 has no .kind::f64, the FP64 tensor path on sm_100a is the warp-level DMMA (DESIGN.md 4)")
for n in sorted(order):
    c = counts[n]
    if c["_total"] < 40:
        continue
    print("%-90s total %6d  %s" % (n[:90], c["_total"], "  ".join("%s %d" % (k, c[k]) for k in KEY if c[k])))
