"""Randomised parity sweep of the CUDA path against the oracle (beyond the fixed seeds of tests/test_gpu_parity.py).

    python tools/fuzz_parity.py [n_cases] [seed]

Shapes 1..260 x 1..48 (a fifth of the cases up to 400 x 160: several waves of blocks, tail-wave split), windows 1..51, disparity ranges up to 300 (multi-chunk), minDisparity up to 9, both matchers,
consistent on/off.  Every case goes through parity.check_cost + parity.check_staged.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import simplestereo_b200 as ss  # noqa: E402
from simplestereo_b200.synth import synth_pair  # noqa: E402
from tests import parity  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
rng = np.random.default_rng(seed)
t0 = time.time()
flips = 0
for case in range(n):
    w, h = int(rng.integers(1, 261)), int(rng.integers(1, 49))
    mind = int(rng.integers(0, 10)) if rng.random() < 0.4 else 0
    maxd = mind + int(rng.choice([0, 3, 15, 31, 32, 63, 64, 100, 127, 128, 200, 300]))
    tall = rng.random() < 0.2                               # enough tiles for more than one wave of 148 blocks: the tail-wave split
    if tall:
        w, h = int(rng.integers(97, 400)), int(rng.integers(49, 161))
        maxd = mind + int(rng.choice([70, 100, 127, 128, 200]))
    kind = rng.random()
    pair_seed = int(rng.integers(0, 1 << 30))
    if kind < 0.75:
        l, r, _ = synth_pair(w, h, maxd, pair_seed)
    elif kind < 0.9:
        r2 = np.random.default_rng(pair_seed)
        l = r2.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = r2.integers(0, 256, (h, w, 3), dtype=np.uint8)
    else:
        l = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
        r = l.copy()
    stress = kind >= 0.75                                   # saturated / exact ties: only near-tie adjudication applies
    desc = None
    try:
        if rng.random() < 0.6:
            kw = dict(winSize=int(rng.choice([1, 3, 5, 9, 15, 21] if tall else [1, 3, 5, 9, 15, 21, 33, 35, 37, 41, 51])), maxDisparity=maxd, minDisparity=mind,
                      gammaC=float(rng.uniform(2, 25)), gammaP=float(rng.uniform(4, 40)), consistent=bool(rng.integers(0, 2)))
            desc = ("asw", w, h, kw)
            gpu = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True)
            ref = oracle.asw(l, r, stages=True, cost=True, **kw)
            parity.check_cost(gpu["cost"], ref["cost"])
            nl, nr = parity.check_staged(gpu, ref, ref["cost"], ref["cost"], mind, kw["consistent"],
                                         None if stress else 0.02)
        else:
            kw = dict(winSize=int(rng.choice([1, 3, 5, 7, 9, 11] if tall else [1, 3, 5, 7, 9, 11, 15, 21, 35, 51])), maxDisparity=maxd, minDisparity=mind,
                      gamma=int(rng.integers(2, 30)), fMax=float(rng.uniform(20, 300)), iterations=int(rng.integers(0, 4)), bins=20)
            desc = ("gsw", w, h, kw)
            gpu = ss.passive.StereoGSW(**kw).compute_staged(l, r, cost=True)
            ref = oracle.gsw(l, r, stages=True, cost=True, **kw)
            parity.check_cost(gpu["cost_left"], ref["cost_left"], "cost_left")
            parity.check_cost(gpu["cost_right"], ref["cost_right"], "cost_right")
            nl, nr = parity.check_staged(gpu, ref, ref["cost_left"], ref["cost_right"], mind, True,
                                         None if stress else 0.02, saturation=None)
        flips += nl + nr
    except AssertionError as ex:
        print("FAIL case", case, desc, "pair", "synth" if kind < 0.75 else ("noise" if kind < 0.9 else "const"), pair_seed, "->", str(ex)[:300], flush=True)
        sys.exit(1)
print(f"fuzz ok: {n} cases (seed {seed}), {flips} adjudicated near-tie flips in total, {time.time() - t0:.0f} s", flush=True)
