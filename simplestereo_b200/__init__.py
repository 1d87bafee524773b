"""
simplestereo_b200 -- B200 (sm_100a) replacement for the ASW / GSW hot path of decadenza/SimpleStereo.

    import simplestereo_b200 as ss
    disp = ss.passive.StereoASW(winSize=35, maxDisparity=127).compute(left_bgr, right_bgr)

Only ``ss.passive`` exists: everything else in SimpleStereo (rigs, calibration, rectification,
structured light, ...) is out of scope.  There is no CPU fallback.
"""
from . import passive  # noqa: F401
from . import synth  # noqa: F401

__version__ = "0.1.0"
