#!/bin/bash
# round-2 GPU call 1: correctness of the fused L-R / multi-device / chunk-grid changes + where-does-the-time-go experiments
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.csv
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2c1_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2c1_parity_report.json 2>/dev/null
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2c1_smoke.log 2>&1
{
  timeout 300 python tools/time_configs.py c2 c2c c3 c1 reps=10
  for fr in 2 4 6 8 10; do SS_FREERUN=$fr timeout 120 python tools/time_configs.py c2 reps=10; done
  SS_TCDEN=0 timeout 120 python tools/time_configs.py c2 reps=10
  SS_TCDEN=0 SS_FREERUN=1 timeout 120 python tools/time_configs.py c2 reps=10
  timeout 600 python tools/time_configs.py c4 c5 reps=3
} > gpurun_out/r2c1_timing.txt 2>&1
timeout 120 ./tools/microbench6 > gpurun_out/r2c1_microbench6.txt 2>&1
timeout 300 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,sm__cycles_elapsed.max,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum --launch-skip 0 --launch-count 63 --csv --log-file gpurun_out/r2c1_ncu_microbench6.csv ./tools/microbench6 > /dev/null 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
echo done
