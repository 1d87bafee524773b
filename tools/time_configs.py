"""Time the aggregation kernel and the whole device-resident call on the BASELINE.json configs (one GPU).

    python tools/time_configs.py [c2] [c3] [c4] [c5] [lib=<path to libsspassive.so>] [reps=N]

Prints one line per config: whole-call ms (CUDA events), aggregation-kernel ms, Mpix*disp/s.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplestereo_b200 import _cabi  # noqa: E402

args = sys.argv[1:]
reps = 5
for a in args:
    if a.startswith("lib="):
        _cabi.LIB_PATH = a[4:]
    if a.startswith("reps="):
        reps = int(a[5:])
which = [a for a in args if "=" not in a] or ["c2"]

import torch  # noqa: E402
from simplestereo_b200.synth import synth_pair  # noqa: E402

L = _cabi.lib()
_cabi.check(L.ss_init(0))
st = torch.cuda.current_stream()

CONFIGS = {
    #       W     H    win maxD  kind  consistent
    "c1": (384, 288, 35, 16, "asw", 0),
    "c2": (1242, 375, 35, 127, "asw", 0),
    "c2c": (1242, 375, 35, 127, "asw", 1),
    "c3": (1242, 375, 35, 127, "gsw", 1),
    "c4": (2880, 1988, 51, 255, "asw", 1),
    "c5": (3840, 2160, 35, 511, "asw", 0),
}

for name in which:
    W, H, win, maxD, kind, cons = CONFIGS[name]
    left, right, _ = synth_pair(W, H, maxD, 0)
    dl, dr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
    out = torch.empty((H, W), dtype=torch.int16, device="cuda")

    def step():
        if kind == "asw":
            _cabi.check(L.ss_asw_compute_device(dl.data_ptr(), dr.data_ptr(), W, H, win, maxD, 0, 5.0, 17.5, cons, 0, H,
                                                out.data_ptr(), st.cuda_stream))
        else:
            _cabi.check(L.ss_gsw_compute_device(dl.data_ptr(), dr.data_ptr(), W, H, win, maxD, 0, 10, 120.0, 3, 20, 0, H,
                                                out.data_ptr(), st.cuda_stream))

    n = reps if name in ("c1", "c2", "c2c", "c3") else max(1, reps // 3)
    for _ in range(2 if n > 1 else 1):
        step()
    torch.cuda.synchronize()
    L.ss_profile_reset()
    L.ss_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms, nl, tot = _cabi.profile_read()
    L.ss_profile_enable(0)
    call_ms = e0.elapsed_time(e1) / n
    D = maxD + 1
    print(f"{name} {kind} {W}x{H} D={D} win={win} cons={cons} lib={os.path.basename(os.path.dirname(_cabi.LIB_PATH))}/"
          f"{os.path.basename(_cabi.LIB_PATH)} env={ {k: v for k, v in os.environ.items() if k.startswith('SS_')} }: "
          f"call {call_ms:.3f} ms, aggregate {ms / n:.3f} ms ({nl // n} launches), "
          f"{W * H * D / call_ms / 1e3:.1f} Mpix*disp/s, checksum {int(out.to(torch.int64).sum())}", flush=True)
    del dl, dr, out
    torch.cuda.empty_cache()
