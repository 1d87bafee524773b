#!/bin/bash
# round-2 GPU call 8: conflict-free raw-cost tile (128-byte columns + 24 bytes per 8 columns) in k_aggregate_tc
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
{
  timeout 60 python tools/time_configs.py c2 reps=10
  timeout 60 python tools/time_configs.py c2c reps=10
  SS_FREERUN=2 timeout 60 python tools/time_configs.py c2 reps=10
  timeout 120 python tools/time_configs.py c5 reps=3
} > gpurun_out/r2c8_timing.txt 2>&1
timeout 800 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -40 > gpurun_out/r2c8_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2c8_parity_report.json 2>/dev/null
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2c8_smoke.log 2>&1
timeout 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum -k regex:k_aggregate_tc -c 1 --csv --log-file gpurun_out/r2c8_ncu_tc.csv python tools/time_configs.py c2 reps=1 > /dev/null 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c8_bench.json 2> gpurun_out/r2c8_bench.err
echo done
