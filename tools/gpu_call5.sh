#!/bin/bash
# round-2 GPU call 4: A/B of the k_aggregate_tc variants (deferred MMA issue, producer unroll, register split), consumer-only
# wavefront counters on the real kernel, test suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
{
  timeout 60 python tools/time_configs.py c2 reps=10
  for v in u3 u1 r144 r128u1; do timeout 60 python tools/time_configs.py c2 reps=10 lib=$PWD/gpurun_variants/libss_$v.so; done
  timeout 60 python tools/time_configs.py c2 reps=10
  timeout 60 python tools/time_configs.py c2c c1 c3 reps=10
  for fr in 2 4 8 10; do SS_FREERUN=$fr timeout 60 python tools/time_configs.py c2 reps=10; done
} > gpurun_out/r2c5_timing.txt 2>&1
timeout 60 ./tools/microbench7 > gpurun_out/r2c5_microbench7.txt 2>&1
timeout 120 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum --csv --log-file gpurun_out/r2c5_ncu_microbench7.csv ./tools/microbench7 > /dev/null 2>&1
SS_FREERUN=10 timeout 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,smsp__inst_executed.sum,gpu__time_duration.sum -k regex:k_aggregate_tc -c 1 --csv --log-file gpurun_out/r2c5_ncu_consumers_only.csv python tools/time_configs.py c2 reps=1 > /dev/null 2>&1
timeout 700 python -m pytest tests -m gpu -q --timeout 200 2>&1 | tail -40 > gpurun_out/r2c5_pytest.log
cp gpurun_out/parity_report.json gpurun_out/r2c5_parity_report.json 2>/dev/null
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2c5_smoke.log 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err
echo done
