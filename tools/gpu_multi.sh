#!/bin/bash
# multi-GPU evidence (gpurun --gpus N -- bash tools/gpu_multi.sh N): in-process devices= path, both torchrun partitions, bench
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
export PYTHONPATH=$PWD
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/multi_n${N}_gpus.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "devices_kwarg or different_streams or shards" > $O/multi_n${N}_pytest.log 2>&1; tail -3 $O/multi_n${N}_pytest.log
for cfg in c2 c5; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/run_sharded.py $cfg >> $O/multi_n${N}_sharded.txt 2>> $O/multi_n${N}_sharded.err
done
cat $O/multi_n${N}_sharded.txt
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > $O/multi_n${N}_bench.json 2> $O/multi_n${N}_bench.err
cut -c1-400 $O/multi_n${N}_bench.json
SS_NO_TAIL_SPLIT=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --no-other-configs > $O/multi_n${N}_bench_nosplit.json 2>> $O/multi_n${N}_bench.err
cut -c1-200 $O/multi_n${N}_bench_nosplit.json
