"""The numpy restatements in oracle/post_oracle.py against OpenCV's own outputs (golden fixture + live cv2)."""
import os

import numpy as np
import pytest

import oracle.post_oracle as po

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_outputs.npz"))
NAMES = sorted({k.split("__")[0] for k in GOLD.files if "__" in k})


def case(name):
    return {k.split("__")[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "__")}


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("reproject")])
def test_reproject_matches_cv2_golden(name):
    c = case(name)
    got = po.reproject(c["disp"], c["Q"])
    assert got.dtype == np.float32 and np.array_equal(got, c["points"], equal_nan=True)


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("colormap")])
def test_normalize_colormap_matches_cv2_golden(name):
    c = case(name)
    g, bgr = po.normalize_colormap(c["disp"], GOLD["jet_lut"])
    assert np.array_equal(g, c["gray"]) and np.array_equal(bgr, c["bgr"])


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("remap")])
def test_remap_matches_cv2_golden(name):
    c = case(name)
    assert np.array_equal(po.remap_linear(c["src"], c["mapx"], c["mapy"]), c["dst"])


def test_q_matrices_follow_the_reference_formulas():
    """points.py:147-174 with b=1, fx=fy=W, cx=W/2, cy=H/2, no shear."""
    Q = po.adimensional_q(384, 288)
    want = np.array([[1, 0, 0, -192.0], [0, 1, 0, -144.0], [0, 0, 0, -384.0], [0, 0, 1.0, 0]])
    assert np.array_equal(Q, want)


def test_live_cv2_random_inputs():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for _ in range(4):
        d = rng.integers(-5, 700, (int(rng.integers(3, 50)), int(rng.integers(3, 70)))).astype(np.int16)
        Q = rng.normal(size=(4, 4))
        assert np.array_equal(po.reproject(d, Q), cv2.reprojectImageTo3D(d, Q), equal_nan=True)
        g = cv2.normalize(d, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1)
        assert np.array_equal(po.normalize_minmax_u8(d), g)
        src = rng.integers(0, 256, (int(rng.integers(5, 60)), int(rng.integers(5, 60)), 3), dtype=np.uint8)
        mx = rng.uniform(-5, src.shape[1] + 5, d.shape).astype(np.float32)
        my = rng.uniform(-5, src.shape[0] + 5, d.shape).astype(np.float32)
        assert np.array_equal(po.remap_linear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))
