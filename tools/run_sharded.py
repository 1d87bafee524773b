"""Multi-GPU check of both partitions of SURVEY.md 8(e) over NCCL (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/run_sharded.py [c2|c5]

Row stripes and disparity-range shards must both reproduce the unsharded map bit for bit on every rank; prints
device-timed ms per frame (max over ranks) for unsharded / rows / disparity.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import simplestereo_b200 as ss  # noqa: E402
from simplestereo_b200 import _cabi  # noqa: E402
from simplestereo_b200.sharding import ShardedStereoASW  # noqa: E402
from simplestereo_b200.synth import synth_pair  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
W, H, maxD, cons = {"c2": (1242, 375, 127, True), "c5": (3840, 2160, 511, False)}[cfg]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
_cabi.check(_cabi.lib().ss_init(local))
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
left, right, _ = synth_pair(W, H, maxD, 0)
dl, dr = torch.from_numpy(left).to(dev), torch.from_numpy(right).to(dev)
m = ss.passive.StereoASW(35, maxD, 0, 5.0, 17.5, consistent=cons)
st = torch.cuda.current_stream()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps):
        out = fn()
    e1.record(st)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


reps = 5 if cfg == "c2" else 2
one = ShardedStereoASW(m, mode="rows")
one.world, one.rank = 1, 0                                    # the unsharded call on every rank
want, t_one = timed(lambda: one.compute_device(dl, dr).clone(), reps)
res = {}
for mode in ("rows", "disparity"):
    sh = ShardedStereoASW(m, mode=mode)
    got, t = timed(lambda: sh.compute_device(dl, dr), reps)
    same = bool(torch.equal(got, want))
    flag = torch.tensor([int(same)], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res[mode] = (t, bool(flag.item()))
if rank == 0:
    D = maxD + 1
    print(f"{cfg} {W}x{H} D={D} consistent={cons} on {world} GPU(s): unsharded {t_one:.3f} ms | "
          + " | ".join(f"{k} {t:.3f} ms ({W * H * D / t / 1e3:.0f} Mpix*disp/s, x{t_one / t:.2f}, identical on all ranks: {ok})"
                       for k, (t, ok) in res.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if all(ok for _, ok in res.values()) else 1)
