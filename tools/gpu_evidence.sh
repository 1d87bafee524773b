#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/gpu_evidence.sh): tests, bench lines, ncu launch list and full captures.
# Everything lands in gpurun_out/; tools/evidence_to_profiles.sh turns it into the tracked summaries under profiles/.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cut -c1-300 $O/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; cut -c1-200 $O/bench_reference_arm.json
python tools/time_configs.py c1 c2 c2c c3 c4 c5 > $O/configs.txt 2>&1; cat $O/configs.txt
# every launch of a short bench run with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > $O/launches_bench.log 2>&1
# the dominant kernels, full set, one launch each
ncu --set full --clock-control none --import-source on -k regex:k_aggregate_tc -s 1 -c 1 -o $O/asw_ws \
    python tools/time_configs.py c2 reps=1 > $O/ncu_asw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_aggregate_ws -s 1 -c 1 -o $O/gsw_ws \
    python tools/time_configs.py c3 reps=1 > $O/ncu_gsw.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.csv
ls -la $O
