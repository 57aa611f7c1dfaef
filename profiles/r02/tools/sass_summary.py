"""SASS opcode summary of every kernel in the product library (no GPU needed).

    python profiles/r02/tools/sass_summary.py [lib.so] > profiles/r02/sass_opcodes.txt

Per kernel: instructions, registers / shared / stack (cuobjdump -res-usage) and the counts of the opcodes that identify
the hardware path: FP64 math (DFMA/DMUL/DADD), FP32 math (FFMA), tensor core (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit), bulk async copies (UBLKCP = cp.async.bulk), cp.async (LDGSTS), mbarrier ops (SYNCS), shuffles,
local-memory traffic (LDL/STL).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "curiousrl_b200", "libcuriousrl_b200.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
    elif name and "REG:" in line:
        usage[name] = line.strip()
        name = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
groups = [("FP64", ("DFMA", "DMUL", "DADD")), ("FP32", ("FFMA", "FMUL", "FADD")), ("tcgen05.mma", ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA")),
          ("tcgen05.ld", ("LDTM",)), ("tcgen05.commit", ("UTCBAR",)), ("cp.async.bulk", ("UBLKCP",)), ("cp.async", ("LDGSTS",)),
          ("mbarrier", ("SYNCS",)), ("SHFL", ("SHFL",)), ("LDS/STS", ("LDS", "STS")), ("LDG/STG", ("LDG", "STG", "LD", "ST")),
          ("local", ("LDL", "STL")), ("MUFU", ("MUFU",)), ("BAR", ("BAR",))]
cur, counts, total = None, None, 0
out = []


def flush():
    if cur is None:
        return
    demangled = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
    demangled = re.sub(r"\(.*", "", demangled)
    cells = ["%s %d" % (g, sum(counts[o] for o in ops)) for g, ops in groups if sum(counts[o] for o in ops)]
    out.append((demangled, total, usage.get(cur, ""), ", ".join(cells)))


for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        cur, counts, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[m.group(2)] += 1
        total += 1
flush()
print("SASS opcode summary of %s (sm_100a)\n" % os.path.basename(lib))
for name, n, use, cells in sorted(out, key=lambda r: -r[1]):
    print("%s\n    %d instructions; %s\n    %s\n" % (name, n, use, cells))
