"""Instruction mix of the hottest loop of a kernel, from `cuobjdump -sass -fun <mangled> lib.so` output.
usage: python profiles/sass_loop_mix.py sass.txt  -> finds the backward branches, prints per-loop opcode histogram."""
import collections
import re
import sys

ins = []
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            loops.append((addr[tgt], i))
loops.sort(key=lambda p: p[0] - p[1])
print("total instructions", len(ins), "backward branches", len(loops))
for lo, hi in loops[:3]:
    body = ins[lo:hi + 1]
    hist = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        op = t.split()[0]
        hist[op.split(".")[0]] += 1
    print(f"\nloop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instructions")
    for op, n in hist.most_common(40):
        print(f"  {op:10s} {n:5d} {100.0 * n / len(body):5.1f}%")
