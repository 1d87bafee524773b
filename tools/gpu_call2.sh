#!/bin/bash
# round-2 GPU call 2: utility-warp kernel, tail split, test suite with per-test timeouts, LDS pattern microbenchmark
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
{
  timeout 100 python tools/time_configs.py c2 reps=10
  timeout 100 python tools/time_configs.py c2c c1 c3 reps=10
  for fr in 2 4 10 12; do SS_FREERUN=$fr timeout 100 python tools/time_configs.py c2 reps=10; done
} > gpurun_out/r2c2_timing.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 240 2>&1 | tail -40 > gpurun_out/r2c2_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2c2_parity_report.json 2>/dev/null
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2c2_smoke.log 2>&1
timeout 100 ./tools/microbench7 > gpurun_out/r2c2_microbench7.txt 2>&1
timeout 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum --csv --log-file gpurun_out/r2c2_ncu_microbench7.csv ./tools/microbench7 > /dev/null 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err
{
  timeout 300 python tools/time_configs.py c4 c5 reps=3
} >> gpurun_out/r2c2_timing.txt 2>&1
echo done
