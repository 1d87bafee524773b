"""
simplestereo_b200 -- B200 (sm_100a) replacement for the ASW / GSW hot path of decadenza/SimpleStereo.

    import simplestereo_b200 as ss
    disp = ss.passive.StereoASW(winSize=35, maxDisparity=127).compute(left_bgr, right_bgr)

``ss.passive`` is the hot path.  ``ss.rectify`` (the remap that feeds it), ``ss.points`` (the 3-D reprojection
that follows it) and ``ss.display`` (the examples' min-max + colour-map post-filter) are the "next" rows of
SURVEY.md 8(f); everything else in SimpleStereo (calibration, rectification algebra, structured light, ...) is
out of scope.  There is no CPU fallback.
"""
from . import passive  # noqa: F401
from . import points  # noqa: F401
from . import rectify  # noqa: F401
from . import display  # noqa: F401
from . import synth  # noqa: F401

__version__ = "0.1.0"
