#!/usr/bin/env python3
"""
Generate tests/golden/ply_*.ply with the reference's own exportPLY (simplestereo/points.py:10-80, loaded from
/root/reference by file path -- the package itself cannot be imported) on the inputs of ply_cases().

    python tests/golden/make_golden_ply.py        # build container only (needs /root/reference)
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def ply_cases():
    rng = np.random.default_rng(11)
    pts = (rng.normal(size=(7, 9, 3)) * np.array([100.0, 50.0, 1000.0])).astype(np.float32)
    pts[0, 0] = [np.inf, -np.inf, np.nan]                 # what reprojection yields at w == 0
    pts[0, 1] = [0.0, -0.0, 1e-7]
    pts[0, 2] = [123456.7890625, -0.0000005, 2.5]
    return {
        "plain_f32": dict(points3D=pts, referenceImage=None, precision=6),
        "bgr_f32_p3": dict(points3D=pts, referenceImage=rng.integers(0, 256, (7, 9, 3), dtype=np.uint8), precision=3),
        "plain_f64_p10": dict(points3D=rng.normal(size=(5, 3)) * 1e3, referenceImage=None, precision=10),
        "gray_int64": dict(points3D=pts, referenceImage=rng.integers(0, 4000, (7, 9)).astype(np.int64), precision=4),
        "gray_u8_as_float": dict(points3D=pts, referenceImage=rng.integers(0, 256, (7, 9), dtype=np.uint8), precision=6),
        "gray_float": dict(points3D=pts, referenceImage=rng.normal(size=(7, 9)) * 10, precision=8),
        "empty": dict(points3D=np.zeros((0, 3), np.float32), referenceImage=None, precision=6),
    }


def main():
    spec = importlib.util.spec_from_file_location("ref_points", "/root/reference/simplestereo/points.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    for name, kw in ply_cases().items():
        path = os.path.join(HERE, f"ply_{name}.ply")
        ref.exportPLY(kw["points3D"], path, kw["referenceImage"], kw["precision"])
        print(name, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
