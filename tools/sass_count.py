#!/usr/bin/env python
"""Static instruction mix of a kernel's SASS (no GPU needed).

usage: tools/sass_count.py <object-or-so> <function-regex> [--loop]

Prints, for every matching function, the instruction count by mnemonic class for the whole function and for its
LARGEST backward-branch loop body (the k march of the tendency kernels), which is the offline proxy used to track
instructions per cell between ncu captures."""
import collections
import re
import subprocess
import sys


def classify(op):
    if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        return "FP64:" + op
    return op


def parse(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return funcs


def opcode(txt):
    t = txt.split()
    if t[0].startswith("@"):
        t = t[1:]
    return t[0].split(".")[0]


def report(name, ins):
    addrs = [a for a, _ in ins]
    loops = []
    for a, t in ins:
        if opcode(t) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < a:
                    loops.append((a - tgt, tgt, a))
    print("==", name, "instructions:", len(ins))
    def mix(sub, label):
        c = collections.Counter(classify(opcode(t)) for _, t in sub)
        fp64 = sum(v for k, v in c.items() if k.startswith("FP64:"))
        print("  %s: total %d, FP64 %d, other %d" % (label, len(sub), fp64, len(sub) - fp64))
        print("   ", ", ".join("%s %d" % kv for kv in c.most_common(24)))
    mix(ins, "whole")
    for span, tgt, a in sorted(loops, reverse=True)[:2]:
        sub = [(x, t) for x, t in ins if tgt <= x <= a]
        mix(sub, "loop 0x%x..0x%x" % (tgt, a))


def main():
    path, rx = sys.argv[1], re.compile(sys.argv[2])
    for name, ins in parse(path).items():
        if rx.search(name):
            report(name, ins)


if __name__ == "__main__":
    main()
