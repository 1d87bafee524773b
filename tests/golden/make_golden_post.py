#!/usr/bin/env python3
"""
Generate tests/golden/post_outputs.npz: outputs of OpenCV itself (the library the reference calls for the steps either
side of the hot path -- points.py:176, _rigs.py:564-565,628, examples/010:44-45) on small deterministic inputs.

    python tests/golden/make_golden_post.py        # needs cv2 (present in the build container and on the GPU box)

Stored: the COLORMAP_JET look-up table, and for every case the inputs (rebuilt by post_cases()) and cv2's outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def post_cases():
    """Deterministic inputs: name -> dict."""
    import oracle.post_oracle as po
    c = {}
    rng = np.random.default_rng(2024)
    # -- reprojection: adimensional Q (w == 0 at d == 0 -> inf), a rig-like Q with shear, negative disparities (-1 = invalid) --
    d = rng.integers(0, 64, (40, 56)).astype(np.int16)
    d[0, :5] = 0
    c["reproject_adim"] = dict(disp=d, Q=po.adimensional_q(56, 40))
    d2 = rng.integers(-1, 300, (33, 47)).astype(np.int16)
    c["reproject_rig"] = dict(disp=d2, Q=po.rig_q(b=119.7, fx=1432.1, fy=1431.6, cx1=633.2, cx2=655.9, a1=0.31, a2=-0.12, cy=371.4))
    # large enough that "multiply by 1/w" and "divide by w" differ on dozens of values (OpenCV does the former)
    c["reproject_rig_shear_wide"] = dict(disp=rng.integers(0, 64, (48, 640)).astype(np.int16),
                                         Q=po.rig_q(b=0.54, fx=721.5, fy=721.5, cx1=609.6, cx2=609.6, a1=0.3, a2=0.3, cy=172.9))
    c["reproject_random_q"] = dict(disp=rng.integers(-16, 512, (21, 64)).astype(np.int16), Q=rng.normal(size=(4, 4)))
    # -- normalise + JET --
    c["colormap_range17"] = dict(disp=rng.integers(0, 17, (48, 64)).astype(np.int16))
    c["colormap_range300"] = dict(disp=rng.integers(-1, 300, (31, 77)).astype(np.int16))
    c["colormap_constant"] = dict(disp=np.full((9, 13), 7, np.int16))
    c["colormap_negative"] = dict(disp=rng.integers(-200, -3, (16, 16)).astype(np.int16))
    # -- remap: noisy warp with out-of-image and non-finite coordinates; a real undistort-rectify map --
    src = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    y, x = np.mgrid[0:52, 0:70].astype(np.float32)
    mx = (x * 83 / 70 + rng.normal(0, 3, (52, 70))).astype(np.float32) - 4
    my = (y * 61 / 52 + rng.normal(0, 3, (52, 70))).astype(np.float32) - 3
    mx[0, 0], my[1, 1], mx[2, 2], mx[3, 3] = np.nan, np.inf, 1e9, -1e9
    c["remap_noisy"] = dict(src=src, mapx=mx, mapy=my)
    import cv2
    K = np.array([[95.0, 0.4, 47.0], [0, 94.0, 31.0], [0, 0, 1]])
    dist = np.array([-0.31, 0.12, 0.001, -0.002, -0.02])
    R, _ = cv2.Rodrigues(np.array([0.02, -0.03, 0.01]))
    Knew = np.array([[90.0, 0, 50.0], [0, 90.0, 30.0], [0, 0, 1]])
    mx2, my2 = cv2.initUndistortRectifyMap(K, dist, R, Knew, (100, 64), cv2.CV_32FC1)
    c["remap_rectify"] = dict(src=rng.integers(0, 256, (64, 96, 3), dtype=np.uint8), mapx=mx2, mapy=my2)
    return c


def main():
    import cv2
    out = {"jet_lut": cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(1, 256), cv2.COLORMAP_JET)[0]}
    for name, c in post_cases().items():
        for k, v in c.items():
            out[f"{name}__{k}"] = v
        if name.startswith("reproject"):
            out[f"{name}__points"] = cv2.reprojectImageTo3D(c["disp"], c["Q"])
        elif name.startswith("colormap"):
            g = cv2.normalize(c["disp"], None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1)
            out[f"{name}__gray"] = g
            out[f"{name}__bgr"] = cv2.applyColorMap(g, cv2.COLORMAP_JET)
        else:
            out[f"{name}__dst"] = cv2.remap(c["src"], c["mapx"], c["mapy"], cv2.INTER_LINEAR)
    p = os.path.join(HERE, "post_outputs.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    main()
