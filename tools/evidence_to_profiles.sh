#!/bin/bash
# Turn gpurun_out/ (written by tools/gpu_evidence.sh on the GPU box) into the tracked summaries under profiles/.
# usage: tools/evidence_to_profiles.sh <tag>     e.g. r01b
set -eu
T=${1:-r01}
O=gpurun_out
P=profiles
for k in asw_tc gsw_ws; do
  [ -f $O/$k.ncu-rep ] || continue
  ncu -i $O/$k.ncu-rep --page raw --csv > $O/$k.raw.csv
  if [ $k = asw_tc ]; then python tools/ncu_summary.py $O/$k.raw.csv $P/${T}_ncu_k_aggregate_tc_asw.md $P/ncu_traffic.json > /dev/null
  else python tools/ncu_summary.py $O/$k.raw.csv $P/${T}_ncu_k_aggregate_ws_${k%_ws}.md > /dev/null; fi
  ncu -i $O/$k.ncu-rep --page source --csv > $O/$k.src.csv
  if [ $k = asw_tc ]; then python tools/ncu_regions.py $O/$k.src.csv > $P/${T}_ncu_k_aggregate_tc_asw_regions.txt
  else python tools/ncu_regions.py $O/$k.src.csv > $P/${T}_ncu_k_aggregate_ws_${k%_ws}_regions.txt; fi
done
cp $O/bench_n1.json $P/${T}_bench_n1.json
cp $O/bench_reference_arm.json $P/${T}_bench_reference_arm.json
cp $O/configs.txt $P/${T}_configs_timing.txt
cp $O/launches.csv $P/${T}_launches.csv
python - "$O/launches.csv" "$P/${T}_launches_summary.md" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows:
    if r is hdr or r[ik] == "Kernel Name":
        continue
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
    n = r[ik][:48]
    a = tot.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
s = sum(v for _, v in tot.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs", "",
       "kernel | launches | total ms | share", "---|---|---|---"]
for n, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{n} | {c} | {v:.3f} | {100 * v / s:.1f}%")
open(sys.argv[2], "w").write("\n".join(out) + "\n")
print("\n".join(out))
PY

for f in sanitizer_memcheck.txt sanitizer_racecheck.txt parity_report.json; do [ -f $O/$f ] && cp $O/$f $P/${T}_$f; done
