#!/usr/bin/env python
"""Instruction mix of the innermost hot loop of a kernel from `cuobjdump -sass` output.

usage: sass_loop_mix.py <lib.so> <mangled-name-substring> [min_len]
Finds backward branches, takes the tightest loop of at least min_len instructions that contains
the most FP64 instructions, and prints the opcode histogram.
"""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    on, rows = False, []
    for line in txt.splitlines():
        if "Function :" in line:
            on = pat in line
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: k for k, (a, _) in enumerate(rows)}
    loops = []
    for k, (a, ins) in enumerate(rows):
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", ins)
        if m:
            t = int(m.group(1), 16)
            if t in addr and addr[t] < k and k - addr[t] >= min_len:
                loops.append((addr[t], k))
    if not loops:
        print("no loop found; function has", len(rows), "instructions")
        return
    def n_fp64(lo, hi):
        return sum(1 for _, i in rows[lo:hi + 1] if re.search(r"\bD(FMA|ADD|MUL|SETP)\b", i))
    loops.sort(key=lambda lh: (lh[1] - lh[0]))
    lo, hi = loops[0]
    print("function instructions:", len(rows), " loops(>=%d):" % min_len, [(h - l + 1) for l, h in loops])
    body = rows[lo:hi + 1]
    hist = collections.Counter()
    for _, ins in body:
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        hist[ins.split()[0].split(".")[0]] += 1
    print("loop length", len(body), " fp64", n_fp64(lo, hi))
    for op, c in hist.most_common():
        print("  %-10s %d" % (op, c))


if __name__ == "__main__":
    main()
