"""One tiny ASW + GSW call (both aggregation kernels) for compute-sanitizer --tool racecheck.
usage: compute-sanitizer --tool racecheck python tools/racecheck_small.py [lib=<alternative libsspassive.so>]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplestereo_b200 import _cabi  # noqa: E402
for a in sys.argv[1:]:
    if a.startswith("lib="):
        _cabi.LIB_PATH = a[4:]
import simplestereo_b200 as ss  # noqa: E402
from simplestereo_b200.synth import synth_pair  # noqa: E402
l, r, _ = synth_pair(120, 3, 70, 0)
a = ss.passive.StereoASW(9, 70, 0, 5.0, 17.5, True).compute(l, r)          # k_aggregate_tc
b = ss.passive.StereoASW(9, 20, 0, 5.0, 17.5, True).compute(l, r)          # k_aggregate_ws, 32-disparity chunks
g = ss.passive.StereoGSW(5, 20).compute(l, r)                              # k_aggregate_ws (GSW)
print("checksums", int(a.sum()), int(b.sum()), int(g.sum()))
