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


def test_import_scans_whole_lines_like_the_reference(tmp_path):
    """points.py:108-111 compares each stripped, lower-cased LINE with "end_header": a comment that merely contains the
    token is not the end of the header; without any end_header line nothing is left to read (empty array, no exception)."""
    p = tmp_path / "c.ply"
    p.write_text("ply\nformat ascii 1.0\ncomment not the end_header\nelement vertex 2\nproperty double x\nEND_HEADER  \n1 2 3\n4 5 6\n")
    assert np.array_equal(ss.points.importPLY(str(p)), [[1, 2, 3], [4, 5, 6]])
    q = tmp_path / "n.ply"
    q.write_text("ply\nformat ascii 1.0\nelement vertex 1\n1 2 3\n")
    out = ss.points.importPLY(str(q))
    assert out.shape == (0,) and out.dtype == float


def test_export_integer_colour_images_and_extreme_precision(tmp_path):
    """The reference formats colour triples with "{:d}" (points.py:53-55): any integer dtype goes, floats raise ValueError.
    Coordinates are "{:.{p}f}" for any p: a 1e300 double at 300 decimals is 600 characters."""
    pts = np.array([[1.5, -2.25, 3.0], [1e300, -1e-300, 0.1]], np.float64)
    img = np.array([[300, -7, 65536], [1, 2, 3]], np.int32)
    p = tmp_path / "i.ply"
    ss.points.exportPLY(pts, str(p), img, precision=2)
    body = p.read_text().split("end_header\n")[1].splitlines()
    assert body[0] == "{:.2f} {:.2f} {:.2f} {:d} {:d} {:d}".format(1.5, -2.25, 3.0, 65536, -7, 300)
    assert body[1] == "{:.2f} {:.2f} {:.2f} {:d} {:d} {:d}".format(1e300, -1e-300, 0.1, 3, 2, 1)
    with pytest.raises(ValueError):
        ss.points.exportPLY(pts, str(p), img.astype(np.float32))
    ss.points.exportPLY(pts, str(p), precision=300)
    body = p.read_text().split("end_header\n")[1].splitlines()
    assert body[1] == "{:.300f} {:.300f} {:.300f}".format(1e300, -1e-300, 0.1)
    g = np.array([0.5, 1e30])                                        # "{:{p}f}": width p, 6 decimals
    ss.points.exportPLY(pts, str(p), g, precision=3)
    body = p.read_text().split("end_header\n")[1].splitlines()
    assert body[1].split(" ")[-1] == "{:3f}".format(1e30)
