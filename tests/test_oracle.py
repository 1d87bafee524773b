"""
Pins the CPU oracle (oracle/passive_oracle.c) BEFORE it is trusted as the parity checker:
  * against the reference's only known-answer fixture (examples/res/tsukuba/disparityASW.png),
  * against outputs of the unmodified reference (tests/golden/ref_outputs.npz, made by make_golden.py),
  * and, when oracle/_ref is present, against the live reference on fresh random crops.
Bit-exact: disparities are int16 and the restatement repeats the reference's double/float ops in order.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle
from tests.golden import cases

GOLD = cases.load_golden()


def _inputs(name, spec):
    l, r = cases.load_inputs(spec)
    md5 = np.frombuffer(hashlib.md5(l.tobytes() + r.tobytes()).digest(), np.uint8)
    if not (md5 == GOLD["md5_" + name]).all():
        pytest.skip("input builder drifted from the machine that made the golden file (cv2/numpy version)")
    return l, r


def test_tsukuba_known_answer_image():
    """examples/010:44-45 recipe applied to the oracle's map reproduces disparityASW.png exactly."""
    import cv2
    l, r = cases.load_inputs(("tsukuba", None))
    d = oracle.asw(l, r, 35, 16, 0, 17.5, 17.5, False)
    img = cv2.applyColorMap(cv2.normalize(d, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1), cv2.COLORMAP_JET)
    kat = cv2.imread(os.path.join(cases.HERE, "disparityASW.png"))
    assert img.shape == kat.shape
    assert (img == kat).all()
    # the colour map is injective over the 17 levels, so this pins every pixel's disparity
    assert len(np.unique(d)) == len(np.unique(kat.reshape(-1, 3), axis=0))


@pytest.mark.parametrize("name,spec,kw", cases.ASW_CASES, ids=[c[0] for c in cases.ASW_CASES])
def test_asw_matches_reference_golden(name, spec, kw):
    l, r = _inputs(name, spec)
    got = oracle.asw(l, r, **kw)
    assert got.dtype == np.int16 and got.shape == GOLD[name].shape
    assert np.array_equal(got, GOLD[name])


@pytest.mark.parametrize("name,spec,kw", cases.GSW_CASES, ids=[c[0] for c in cases.GSW_CASES])
def test_gsw_closed_form_matches_reference_golden(name, spec, kw):
    l, r = _inputs(name, spec)
    assert np.array_equal(oracle.gsw(l, r, **kw), GOLD[name])


@pytest.mark.parametrize("name,spec,kw", [c for c in cases.GSW_CASES if c[2]["winSize"] <= 11],
                         ids=[c[0] for c in cases.GSW_CASES if c[2]["winSize"] <= 11])
def test_gsw_literal_relaxation_matches_reference_golden(name, spec, kw):
    l, r = _inputs(name, spec)
    assert np.array_equal(oracle.gsw(l, r, literal=True, **kw), GOLD[name])


def test_asw_right_pass_identity():
    """C_R[xr,disp] == C_L[xr+disp,disp] bit-for-bit: literal right pass == diagonal re-read."""
    l, r = cases.load_inputs(("synth", (96, 30, 24, 11)))
    a = oracle.asw(l, r, 21, 24, 0, 5, 17.5, True, stages=True)
    b = oracle.asw(l, r, 21, 24, 0, 5, 17.5, True, stages=True, literal_right=True)
    for k in ("final", "left", "right", "invalid"):
        assert np.array_equal(a[k], b[k]), k


def test_stage_semantics():
    l, r = cases.load_inputs(("tsukuba", (100, 140, 100, 220)))
    s = oracle.asw(l, r, 21, 16, 0, 5, 17.5, True, stages=True, cost=True)
    plain = oracle.asw(l, r, 21, 16, 0, 5, 17.5, False)
    assert np.array_equal(s["left"], plain)                      # stage 1 == non-consistent output
    cost = s["cost"]
    H, W, D = cost.shape
    # left map is the smallest-disparity argmin of the volume; x < minD has no candidate -> x
    arg = np.argmin(np.where(np.isfinite(cost), cost, np.inf), axis=2)
    assert np.array_equal(arg.astype(np.int16), s["left"])
    assert (s["final"] >= 0).all()
    assert np.array_equal(s["final"][s["invalid"] == 0], s["left"][s["invalid"] == 0])


def test_rows_subset():
    l, r = cases.load_inputs(("tsukuba", (100, 150, 100, 200)))
    full = oracle.asw(l, r, 15, 16, 0, 5, 17.5, True)
    part = oracle.asw(l, r, 15, 16, 0, 5, 17.5, True, rows=(10, 23))
    assert np.array_equal(full[10:23], part[10:23])
    fullg = oracle.gsw(l, r, 7, 16, 0, 10, 120, 3, 20)
    partg = oracle.gsw(l, r, 7, 16, 0, 10, 120, 3, 20, rows=(40, 50))
    assert np.array_equal(fullg[40:50], partg[40:50])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_random_crops():
    rng = np.random.default_rng(123)
    l, r = cases.load_inputs(("tsukuba", None))
    for _ in range(3):
        h, w = int(rng.integers(8, 40)), int(rng.integers(16, 90))
        y0, x0 = int(rng.integers(0, 288 - h)), int(rng.integers(0, 384 - w))
        lc, rc = l[y0:y0 + h, x0:x0 + w].copy(), r[y0:y0 + h, x0:x0 + w].copy()
        win = int(rng.choice([3, 9, 15, 35]))
        mind = int(rng.integers(0, 4)); maxd = mind + int(rng.integers(0, 20))
        cons = bool(rng.integers(0, 2))
        assert np.array_equal(oracle.ref_asw(lc, rc, win, maxd, mind, 5.0, 17.5, cons),
                              oracle.asw(lc, rc, win, maxd, mind, 5.0, 17.5, cons))
        wing = int(rng.choice([3, 5, 9]))
        assert np.array_equal(oracle.ref_gsw(lc, rc, wing, maxd, mind, 10, 120.0, 3, 20),
                              oracle.gsw(lc, rc, wing, maxd, mind, 10, 120.0, 3, 20))
