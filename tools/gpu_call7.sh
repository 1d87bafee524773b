#!/bin/bash
# round-2 GPU call 7: k_aggregate_tc8 (128-column tiles, 8 x 8 lane tiles) against k_aggregate_tc, tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
{
  timeout 60 python tools/time_configs.py c2 reps=10
  SS_TC8=0 timeout 60 python tools/time_configs.py c2 reps=10
  timeout 60 python tools/time_configs.py c2c reps=10
  for fr in 2 4 8 10; do SS_FREERUN=$fr timeout 60 python tools/time_configs.py c2 reps=10; done
  timeout 120 python tools/time_configs.py c5 reps=3
} > gpurun_out/r2c7_timing.txt 2>&1
timeout 700 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -40 > gpurun_out/r2c7_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2c7_parity_report.json 2>/dev/null
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2c7_smoke.log 2>&1
timeout 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum -k regex:k_aggregate_tc -c 1 --csv --log-file gpurun_out/r2c7_ncu_tc8.csv python tools/time_configs.py c2 reps=1 > /dev/null 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c7_bench.json 2> gpurun_out/r2c7_bench.err
echo done
