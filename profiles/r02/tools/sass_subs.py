"""Per-subroutine opcode histogram of one kernel in a cuobjdump -sass listing (development tool).

    cuobjdump -sass lib.so > lib.sass;  python profiles/r02/tools/sass_subs.py lib.sass KERNEL_SUBSTRING [ADDR]

Lists the non-inlined device functions (CALL targets) of the first kernel whose mangled name contains KERNEL_SUBSTRING
with their instruction counts and most frequent opcodes; with ADDR (hex) prints that subroutine's instructions.
"""
import re
import sys

text = open(sys.argv[1]).read().split("\n")
start = next(i for i, l in enumerate(text) if "Function : " in l and sys.argv[2] in l)
end = next((i for i in range(start + 1, len(text)) if "Function : " in text[i]), len(text))
ins = []
for l in text[start:end]:
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = sorted(set(int(x, 16) for _, t in ins for x in re.findall(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)))
bounds = [0] + targets + [ins[-1][0] + 16]
for i in range(len(bounds) - 1):
    body = [(a, x) for a, x in ins if bounds[i] <= a < bounds[i + 1]]
    if len(sys.argv) > 3:
        if bounds[i] == int(sys.argv[3], 16):
            for a, x in body:
                print("%06x  %s" % (a, x))
        continue
    ops = {}
    for _, x in body:
        f = x.split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
    print("%#8x %6d  %s" % (bounds[i], len(body), " ".join("%s:%d" % kv for kv in top)))
