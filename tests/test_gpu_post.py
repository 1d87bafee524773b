"""
GPU parity tests of the steps either side of the hot path: rectification remap, 3-D reprojection,
display post-filter.  Everything goes through the public Python mirrors, i.e. through the C ABI of include/ss_post.h;
oracle/post_oracle.py (pinned against OpenCV in tests/test_post_oracle.py) and the cv2 golden fixture are the checkers.
Bar: bit-exact (integer / byte work; float32 points are compared bit for bit as well, NaN == NaN).
"""
import os

import numpy as np
import pytest

import oracle.post_oracle as po
import simplestereo_b200 as ss
from simplestereo_b200.synth import synth_pair

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "post_outputs.npz"))
NAMES = sorted({k.split("__")[0] for k in GOLD.files if "__" in k})


def case(name):
    return {k.split("__")[1]: GOLD[k] for k in GOLD.files if k.startswith(name + "__")}


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("reproject")])
def test_reproject_golden(name):
    c = case(name)
    got = ss.points.reprojectImageTo3D(c["disp"], c["Q"])
    assert got.dtype == np.float32 and got.shape == c["points"].shape
    assert np.array_equal(got, c["points"], equal_nan=True)


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("colormap")])
def test_colormap_golden(name):
    c = case(name)
    assert np.array_equal(ss.display.normalize(c["disp"]), c["gray"])
    assert np.array_equal(ss.display.applyColorMap(c["disp"]), c["bgr"])


@pytest.mark.parametrize("name", [n for n in NAMES if n.startswith("remap")])
def test_remap_golden(name):
    c = case(name)
    assert np.array_equal(ss.rectify.remap(c["src"], c["mapx"], c["mapy"]), c["dst"])


def test_randomised_shapes_against_oracle():
    rng = np.random.default_rng(99)
    for _ in range(8):
        h, w = int(rng.integers(1, 300)), int(rng.integers(1, 400))
        d = rng.integers(-3, 600, (h, w)).astype(np.int16)
        Q = rng.normal(size=(4, 4))
        assert np.array_equal(ss.points.reprojectImageTo3D(d, Q), po.reproject(d, Q), equal_nan=True)
        assert np.array_equal(ss.points.getAdimensional3DPoints(d), po.reproject(d, po.adimensional_q(w, h)), equal_nan=True)
        g, bgr = po.normalize_colormap(d, ss.display.COLORMAP_JET)
        assert np.array_equal(ss.display.normalize(d), g) and np.array_equal(ss.display.applyColorMap(d), bgr)
        sh, sw = int(rng.integers(1, 200)), int(rng.integers(1, 300))
        src = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        mx = rng.uniform(-4, sw + 4, (h, w)).astype(np.float32)
        my = rng.uniform(-4, sh + 4, (h, w)).astype(np.float32)
        assert np.array_equal(ss.rectify.remap(src, mx, my), po.remap_linear(src, mx, my))


def test_full_size_against_live_cv2():
    """KITTI-size frames through the whole chain, against OpenCV itself on the host."""
    cv2 = pytest.importorskip("cv2")
    l, r, _ = synth_pair(1242, 375, 127, 0)
    K = np.array([[721.5, 0.3, 609.6], [0, 721.5, 172.9], [0, 0, 1]])
    dist = np.array([-0.28, 0.07, 0.0002, -0.0003, 0.0])
    R, _ = cv2.Rodrigues(np.array([0.004, -0.006, 0.002]))
    mx, my = cv2.initUndistortRectifyMap(K, dist, R, K, (1242, 375), cv2.CV_32FC1)      # _rigs.py:540-541
    l2, r2 = ss.rectify.rectifyImages(l, r, mx, my, mx, my)
    assert np.array_equal(l2, cv2.remap(l, mx, my, cv2.INTER_LINEAR))
    assert np.array_equal(r2, cv2.remap(r, mx, my, cv2.INTER_LINEAR))
    m = ss.passive.StereoASW(winSize=9, maxDisparity=63, consistent=True)
    disp = m.compute(l, r)
    Q = ss.points.buildQ(b=0.54, fx=721.5, fy=721.5, cx1=609.6, cx2=609.6, a1=0.3, a2=0.3, cy=172.9)
    pts = ss.points.reprojectImageTo3D(disp, Q)
    assert np.array_equal(pts, cv2.reprojectImageTo3D(disp, Q), equal_nan=True)
    assert np.array_equal(ss.points.get3DPoints(disp, K, K, 0.54), pts, equal_nan=True)
    # fused call: disparity stays on the device between the matcher and the reprojection
    pts2, disp2 = ss.points.computePoints(m, l, r, Q, return_disparity=True)
    assert np.array_equal(disp2, disp) and np.array_equal(pts2, pts, equal_nan=True)
    img = cv2.applyColorMap(cv2.normalize(disp, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1), cv2.COLORMAP_JET)
    assert np.array_equal(ss.display.applyColorMap(disp), img)


def test_tsukuba_known_answer_image_end_to_end_on_device():
    """examples/010:30-45 end to end: matcher + min-max + JET on the GPU == the reference's golden PNG
    (same allowance as test_gpu_parity.test_tsukuba_known_answer_image: float32 near-ties only)."""
    cv2 = pytest.importorskip("cv2")
    from tests.golden import cases
    l, r = cases.load_inputs(("tsukuba", None))
    d = ss.passive.StereoASW(35, 16, 0, 17.5, 17.5, False).compute(l, r)
    img = ss.display.applyColorMap(d)
    kat = cv2.imread(os.path.join(cases.HERE, "disparityASW.png"))
    assert int((img != kat).any(axis=2).sum()) <= 0.001 * d.size
