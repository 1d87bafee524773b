#!/bin/bash
# Run on the GPU box (gpurun -- bash tools/gpu_evidence.sh): tests, bench lines, ncu launch list and full captures, sanitizer
# records.  Everything lands in gpurun_out/; tools/evidence_to_profiles.sh <tag> turns it into the tracked summaries under profiles/.
# Every step runs under its own timeout: a hung kernel must not eat the call.
set -u
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
export PYTHONPATH=$PWD
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 240 > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; cut -c1-300 $O/bench_n1.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; cut -c1-200 $O/bench_reference_arm.json
timeout 300 python tools/time_configs.py c1 c2 c2c c3 c4 c5 > $O/configs.txt 2>&1; cat $O/configs.txt
# every launch of a short bench run with its device time (cold cache, serialised: compare shares)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > $O/launches_bench.log 2>&1
# the dominant kernels, full set, one launch each
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_aggregate_tc -s 1 -c 1 -o $O/asw_tc \
    python tools/time_configs.py c2 reps=1 > $O/ncu_asw.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_aggregate_ws -s 1 -c 1 -o $O/gsw_ws \
    python tools/time_configs.py c3 reps=1 > $O/ncu_gsw.log 2>&1
# sanitizers on the small-shape tests (both aggregation kernels, fused L-R epilogue, tail split, multi-chunk)
SMALL="golden or randomised_small_shapes or multi_chunk or tail_wave or borders or isolated"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SMALL" > $O/sanitizer_memcheck.txt 2>&1; tail -3 $O/sanitizer_memcheck.txt
# racecheck takes 12 minutes under the tool: only with SS_RACECHECK=1
[ -n "${SS_RACECHECK:-}" ] && timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "asw_synth_d140 or asw_crop_consistent or gsw_synth or tail_wave" > $O/sanitizer_racecheck.txt 2>&1; tail -3 $O/sanitizer_racecheck.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.csv
ls -la $O
