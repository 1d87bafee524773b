#!/usr/bin/env python3
"""Summarise an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv) per code region of k_aggregate:
regions are delimited by BAR.SYNC / mbarrier try-wait instructions (phase A | barrier | phase B | tail)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
marks = [k for k, r in enumerate(data) if 'BAR.SYNC' in r[ix['Source']] or 'SYNCS.PHASECHK' in r[ix['Source']]]
print(rows[0][1]); print("total samples", tot, "instructions", len(data))
prev = 0
keys = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for k in marks + [len(data) - 1]:
    seg = data[prev:k + 1]
    if not seg: continue
    s = sum(int(r[ix['# Samples']]) for r in seg)
    inst = sum(int(r[ix['Instructions Executed']]) for r in seg)
    st = {key: sum(int(r[ix[key]]) for r in seg) for key in keys}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print(f"sass {prev:5d}-{k:5d} samples {100*s/tot:5.1f}%  warp-instr {inst/1e6:8.1f}M  ends: {data[k][ix['Source']].strip()[:34]:34s}", ' '.join(f"{a[6:]}={100*b/max(s,1):.0f}%" for a, b in top))
    prev = k + 1
