#!/bin/bash
# round-2 GPU call 9: byte -> float conversion of the raw costs off the XU pipe (SS_TC_CVT variants), timing only
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONPATH=$PWD
{
  timeout 60 python tools/time_configs.py c2 reps=10
  for v in 1 3 5 4; do timeout 60 python tools/time_configs.py c2 reps=10 lib=$PWD/gpurun_variants/cvt$v.so; done
  timeout 60 python tools/time_configs.py c2 reps=10
} > gpurun_out/r2c9_timing.txt 2>&1
echo done
