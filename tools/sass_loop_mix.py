#!/usr/bin/env python3
"""Instruction mix of the innermost hot loop of one kernel in a cubin/.so (static SASS inspection, no GPU).
usage: sass_loop_mix.py <file> <kernel-name-substring>
The hot loop is the smallest backward-branch body holding at least 64 packed FP32 instructions."""
import re, subprocess, sys, collections
path, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
cur, funcs = None, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    addr = {a: k for k, (a, _) in enumerate(ins)}
    best = None
    for k, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                body = ins[addr[tgt]:k + 1]
                npk = sum(1 for _, x in body if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", x))
                if npk >= 64 and (best is None or len(body) < len(best[1])):
                    best = (npk, body)
    if not best:
        continue
    cnt = collections.Counter()
    for _, t in best[1]:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        cnt[t.split()[0]] += 1
    tot = sum(cnt.values())
    print(f"{name[-60:]}: hot loop {tot} instr, packed {best[0]}, total kernel {len(ins)}")
    print("   " + "  ".join(f"{k}:{v}" for k, v in cnt.most_common(14)))
