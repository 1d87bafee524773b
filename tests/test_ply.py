"""points.exportPLY / importPLY: the native writer is byte-identical to the reference's Python loop
(golden files written by the reference's own exportPLY, tests/golden/make_golden_ply.py).  Host-only: no GPU needed."""
import os

import numpy as np
import pytest

import simplestereo_b200 as ss
from tests.golden.make_golden_ply import HERE, ply_cases

CASES = ply_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_export_matches_reference_bytes(name, tmp_path):
    kw = CASES[name]
    out = tmp_path / "out.ply"
    ss.points.exportPLY(kw["points3D"], str(out), kw["referenceImage"], kw["precision"])
    want = open(os.path.join(HERE, f"ply_{name}.ply"), "rb").read()
    assert out.read_bytes() == want


def test_import_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(40, 50, 3)).astype(np.float32) * 100
    p = tmp_path / "cloud.ply"
    ss.points.exportPLY(pts, str(p), rng.integers(0, 256, (40, 50, 3), dtype=np.uint8), precision=7)
    back = ss.points.importPLY(str(p))
    assert back.shape == (2000, 3) and np.allclose(back, pts.reshape(-1, 3), atol=1e-6)
    rgb = ss.points.importPLY(str(p), 3, 4, 5)
    assert rgb.shape == (2000, 3) and rgb.min() >= 0 and rgb.max() <= 255


def test_large_cloud_is_written_in_order(tmp_path):
    """More points than one work item per thread: chunks must land in order."""
    n = 300_000
    pts = np.stack([np.arange(n), np.zeros(n), -np.arange(n)], axis=1).astype(np.float32)
    p = tmp_path / "big.ply"
    ss.points.exportPLY(pts, str(p), precision=1)
    back = ss.points.importPLY(str(p), 0, 2)
    assert np.array_equal(back[:, 0], np.arange(n)) and np.array_equal(back[:, 1], -np.arange(n))


def test_unwritable_path_raises(tmp_path):
    with pytest.raises(ValueError):
        ss.points.exportPLY(np.zeros((2, 3), np.float32), str(tmp_path / "no_such_dir" / "x.ply"))
