"""Run one small ASW call on the SS_DEBUG_ADDR build of libsspassive.so (it printf's the shared-memory addresses of every
consumer lane's three load streams for one block) -- evidence for the wavefront analysis in DESIGN.md.
usage: python tools/dump_addr.py lib=<path to the -DSS_DEBUG_ADDR build>"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from simplestereo_b200 import _cabi  # noqa: E402
for a in sys.argv[1:]:
    if a.startswith("lib="):
        _cabi.LIB_PATH = a[4:]
import simplestereo_b200 as ss  # noqa: E402
from simplestereo_b200.synth import synth_pair  # noqa: E402
l, r, _ = synth_pair(700, 1, 127, 0)
ss.passive.StereoASW(35, 127, 0, 5.0, 17.5, False).compute(l, r)
_cabi.lib().ss_shutdown()
