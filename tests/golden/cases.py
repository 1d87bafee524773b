"""
Case list shared by make_golden.py (which runs the UNMODIFIED reference, oracle/_ref) and the tests.

Every case is (name, kind, input-spec, kwargs).  Inputs are rebuilt deterministically:
  ("tsukuba", (y0, y1, x0, x1))  crop of tests/golden/tsukuba_{l,r}.png (the reference's own fixture,
                                 examples/res/tsukuba/, already rectified -- examples/010:19-21)
  ("synth", (W, H, maxD, seed))  simplestereo_b200.synth.synth_pair
  ("const", (W, H, value))       constant image pair (exact zero-cost ties)
  ("noise", (W, H, seed))        i.i.d. uniform noise pair (saturated-cost ties)
Edge cases follow SURVEY.md section 3.6.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

ASW_CASES = [
    # the reference's known-answer configuration (examples/res/tsukuba/disparityASW.png)
    ("asw_tsukuba_kat", ("tsukuba", None), dict(winSize=35, maxDisparity=16, minDisparity=0, gammaC=17.5, gammaP=17.5, consistent=False)),
    # class defaults (passive.py:59) and examples/010:30
    ("asw_tsukuba_defaults", ("tsukuba", None), dict(winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=False)),
    ("asw_tsukuba_ex010", ("tsukuba", None), dict(winSize=35, maxDisparity=14, minDisparity=4, gammaC=15, gammaP=17.5, consistent=True)),
    ("asw_crop_consistent", ("tsukuba", (90, 150, 120, 300)), dict(winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_crop_win51", ("tsukuba", (60, 130, 40, 200)), dict(winSize=51, maxDisparity=20, minDisparity=0, gammaC=7, gammaP=25.0, consistent=True)),
    ("asw_topleft_corner", ("tsukuba", (0, 40, 0, 90)), dict(winSize=21, maxDisparity=16, minDisparity=2, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_bottomright_corner", ("tsukuba", (248, 288, 294, 384)), dict(winSize=21, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_win_gt_image", ("tsukuba", (100, 120, 100, 130)), dict(winSize=35, maxDisparity=8, minDisparity=0, gammaC=5, gammaP=17.5, consistent=False)),
    ("asw_maxd_ge_width", ("tsukuba", (100, 124, 100, 140)), dict(winSize=9, maxDisparity=64, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_mind_gt_maxd", ("tsukuba", (100, 116, 100, 132)), dict(winSize=7, maxDisparity=3, minDisparity=5, gammaC=5, gammaP=17.5, consistent=False)),
    ("asw_mind_gt_maxd_consistent", ("tsukuba", (100, 116, 100, 132)), dict(winSize=7, maxDisparity=3, minDisparity=5, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_win1", ("tsukuba", (100, 130, 100, 180)), dict(winSize=1, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_const", ("const", (48, 20, 77)), dict(winSize=11, maxDisparity=9, minDisparity=2, gammaC=5, gammaP=17.5, consistent=False)),
    ("asw_const_consistent", ("const", (48, 20, 77)), dict(winSize=11, maxDisparity=9, minDisparity=2, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_synth_d48", ("synth", (160, 48, 47, 3)), dict(winSize=35, maxDisparity=47, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_synth_d140", ("synth", (224, 40, 139, 5)), dict(winSize=35, maxDisparity=139, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_one_row", ("tsukuba", (140, 141, 100, 260)), dict(winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
    ("asw_one_col", ("tsukuba", (100, 160, 140, 141)), dict(winSize=9, maxDisparity=4, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)),
]

GSW_CASES = [
    ("gsw_crop_defaults", ("tsukuba", (100, 150, 100, 240)), dict(winSize=11, maxDisparity=16, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)),
    ("gsw_crop_win35", ("tsukuba", (120, 144, 150, 200)), dict(winSize=35, maxDisparity=12, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)),
    ("gsw_right_border", ("tsukuba", (40, 70, 304, 384)), dict(winSize=9, maxDisparity=12, minDisparity=0, gamma=7, fMax=60.0, iterations=2, bins=20)),
    ("gsw_top_right", ("tsukuba", (0, 24, 310, 384)), dict(winSize=9, maxDisparity=12, minDisparity=2, gamma=7, fMax=60.0, iterations=2, bins=20)),
    ("gsw_mind", ("tsukuba", (100, 130, 100, 200)), dict(winSize=7, maxDisparity=14, minDisparity=4, gamma=12, fMax=90.5, iterations=1, bins=20)),
    ("gsw_iter0", ("tsukuba", (100, 120, 100, 160)), dict(winSize=5, maxDisparity=10, minDisparity=0, gamma=10, fMax=120, iterations=0, bins=20)),
    ("gsw_grey", ("grey", (100, 130, 100, 190)), dict(winSize=9, maxDisparity=16, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)),
    ("gsw_const", ("const", (40, 16, 50)), dict(winSize=7, maxDisparity=6, minDisparity=1, gamma=10, fMax=120, iterations=3, bins=20)),
    ("gsw_synth", ("synth", (128, 32, 40, 7)), dict(winSize=11, maxDisparity=40, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)),
    ("gsw_win_gt_image", ("tsukuba", (100, 112, 100, 124)), dict(winSize=15, maxDisparity=6, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)),
]


def load_inputs(spec):
    import cv2
    kind, arg = spec
    if kind in ("tsukuba", "grey"):
        l = cv2.imread(os.path.join(HERE, "tsukuba_l.png"))
        r = cv2.imread(os.path.join(HERE, "tsukuba_r.png"))
        if kind == "grey":
            l = np.repeat(cv2.cvtColor(l, cv2.COLOR_BGR2GRAY)[:, :, None], 3, 2)
            r = np.repeat(cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)[:, :, None], 3, 2)
        if arg is not None:
            y0, y1, x0, x1 = arg
            l, r = l[y0:y1, x0:x1], r[y0:y1, x0:x1]
        return np.ascontiguousarray(l), np.ascontiguousarray(r)
    if kind == "synth":
        from simplestereo_b200.synth import synth_pair
        w, h, maxd, seed = arg
        l, r, _ = synth_pair(w, h, maxd, seed)
        return l, r
    if kind == "const":
        w, h, v = arg
        a = np.full((h, w, 3), v, np.uint8)
        return a, a.copy()
    if kind == "noise":
        w, h, seed = arg
        rng = np.random.default_rng(seed)
        return (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    raise ValueError(kind)


def load_golden():
    """name -> int16 disparity map produced by the unmodified reference (oracle/_ref)."""
    z = np.load(os.path.join(HERE, "ref_outputs.npz"))
    return {k: z[k] for k in z.files}
