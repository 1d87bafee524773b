"""Warp-stall samples of a SASS range, aggregated per opcode.

    ncu -i report.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_opcode_stalls.py sass.csv <first sass index> <last sass index> [top]

The indices are the instruction numbers `tools/ncu_regions.py` prints (0 = first instruction of the kernel).  One line per opcode:
number of static instructions, share of the range's samples, executed warp-instructions, the four most frequent stall reasons.
"""
import collections
import csv
import sys

path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 12
rows = list(csv.reader(open(path)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], rows[start + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
for k, r in enumerate(data):
    if not lo <= k <= hi or len(r) < len(hdr):
        continue
    op = [t for t in r[ix["Source"]].split() if not t.startswith("@")][0].rstrip(";")
    a = agg[op]
    a[0] += 1
    a[1] += int(r[ix["# Samples"]])
    a[2] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        a[3][s] += int(r[ix[s]])
total = sum(a[1] for a in agg.values()) or 1
print(rows[0][1] if rows and len(rows[0]) > 1 else "")
print(f"sass {lo}-{hi}: {total} samples")
for op, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    why = ", ".join(f"{s[6:]}={100 * v / max(1, a[1]):.0f}%" for s, v in a[3].most_common(4))
    print(f"{op:26s} n={a[0]:4d} samples={100 * a[1] / total:5.1f}% warp-instr={a[2] / 1e6:8.1f}M  {why}")
