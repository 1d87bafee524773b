"""
GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the public Python API,
i.e. through the C ABI of libsspassive.so; the oracle is only the checker.
"""
import os

import numpy as np
import pytest

import oracle
import simplestereo_b200 as ss
from simplestereo_b200.synth import synth_pair
from tests import parity
from tests.golden import cases

pytestmark = pytest.mark.gpu

GOLD = cases.load_golden()
TIE_STRESS = {"asw_const", "asw_const_consistent", "gsw_const"}


def test_library_is_the_cuda_one():
    from simplestereo_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH)
    assert _cabi.lib().ss_init(0) == 0


def _stripe_parity(name, l, r, kw, rows, expect_kernel=None, max_fraction=parity.MAX_MISMATCH_FRACTION):
    """Staged maps + cost volume of an image-row stripe against the oracle; records the statistics for profiles/."""
    from simplestereo_b200 import _cabi
    cons = bool(kw.get("consistent", False))
    gpu = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True, rows=rows)
    kern, chunk = _cabi.last_kernel()
    if expect_kernel is not None:
        assert kern == expect_kernel, f"{name}: served by {kern} (chunk {chunk}), expected {expect_kernel}"
    ref = oracle.asw(l, r, stages=True, cost=True, rows=rows, **kw)
    sl = slice(*rows)
    rr = {k: ref[k][sl] for k in ("left", "right", "invalid", "final")}
    rel = parity.check_cost(gpu["cost"], ref["cost"], name)
    nl, nr = parity.check_staged(gpu, rr, ref["cost"], ref["cost"], kw["minDisparity"], cons, max_fraction)
    parity.record(name, kernel=kern, chunk=chunk, pixels=int(gpu["left"].size), left_mismatch=nl, right_mismatch=nr,
                  max_rel_cost_err=rel, final_identical=bool(np.array_equal(gpu["final"], rr["final"])))
    return gpu, ref, nl, nr


def test_tsukuba_known_answer_image():
    """The reference's own golden image (examples/res/tsukuba/disparityASW.png), via examples/010:44-45.
    Exact, or -- float32 sums against the reference's float64 -- differing only in the pixels listed in
    tests/golden/kat_adjudicated.json, each of which is a near-tie of the reference's own float64 costs."""
    import json
    import cv2
    l, r = cases.load_inputs(("tsukuba", None))
    d = ss.passive.StereoASW(35, 16, 0, 17.5, 17.5, False).compute(l, r)
    assert d.dtype == np.int16 and d.shape == l.shape[:2]
    img = cv2.applyColorMap(cv2.normalize(d, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1), cv2.COLORMAP_JET)
    kat = cv2.imread(os.path.join(cases.HERE, "disparityASW.png"))
    nbad_img = int((img != kat).any(axis=2).sum())
    bad = np.argwhere(d != GOLD["asw_tsukuba_kat"])
    got = sorted([int(y), int(x), int(d[y, x]), int(GOLD["asw_tsukuba_kat"][y, x])] for y, x in bad)
    print(f"Tsukuba KAT: {len(got)} of {d.size} disparities differ from the reference, {nbad_img} pixels of the golden PNG: {got}")
    parity.record("tsukuba_kat", pixels=int(d.size), left_mismatch=len(got), png_pixels_differing=nbad_img, list=got)
    if got:
        ref = oracle.asw(l, r, 35, 16, 0, 17.5, 17.5, False, stages=True, cost=True)
        parity.adjudicate_left(d, GOLD["asw_tsukuba_kat"], ref["cost"], 0)
        allowed = json.load(open(os.path.join(cases.HERE, "kat_adjudicated.json")))["pixels"]
        assert all(g in allowed for g in got), f"pixels outside tests/golden/kat_adjudicated.json: {[g for g in got if g not in allowed]}"
    else:
        assert nbad_img == 0


@pytest.mark.parametrize("name,spec,kw", cases.ASW_CASES, ids=[c[0] for c in cases.ASW_CASES])
def test_asw_golden_cases(name, spec, kw):
    l, r = cases.load_inputs(spec)
    m = ss.passive.StereoASW(**kw)
    out = m.compute(l, r)
    ref = oracle.asw(l, r, stages=True, cost=True, **kw)
    assert np.array_equal(ref["final"], GOLD[name]) or name not in GOLD       # oracle == reference (pinned on CPU too)
    gpu = m.compute_staged(l, r, cost=True)
    assert np.array_equal(gpu["final"], out)
    parity.check_cost(gpu["cost"], ref["cost"].astype(np.float64))
    frac = None if name in TIE_STRESS else parity.MAX_MISMATCH_FRACTION
    nl, nr = parity.check_staged(gpu, ref, ref["cost"], ref["cost"], kw["minDisparity"], kw["consistent"], frac)
    if nl == 0 and nr == 0:
        assert np.array_equal(out, GOLD[name])


@pytest.mark.parametrize("name,spec,kw", cases.GSW_CASES, ids=[c[0] for c in cases.GSW_CASES])
def test_gsw_golden_cases(name, spec, kw):
    l, r = cases.load_inputs(spec)
    m = ss.passive.StereoGSW(**kw)
    out = m.compute(l, r)
    ref = oracle.gsw(l, r, stages=True, cost=True, **kw)
    gpu = m.compute_staged(l, r, cost=True)
    assert np.array_equal(gpu["final"], out)
    parity.check_cost(gpu["cost_left"], ref["cost_left"], "cost_left")
    parity.check_cost(gpu["cost_right"], ref["cost_right"], "cost_right")
    frac = None if name in TIE_STRESS else parity.MAX_MISMATCH_FRACTION
    nl, nr = parity.check_staged(gpu, ref, ref["cost_left"], ref["cost_right"], kw["minDisparity"], True, frac, saturation=None)
    if nl == 0 and nr == 0:
        assert np.array_equal(out, GOLD[name])


def test_saturated_noise_pair_only_flips_near_ties():
    """i.i.d. noise saturates the truncated AD: every argmin is a rounding tie."""
    l, r = cases.load_inputs(("noise", (96, 24, 5)))
    kw = dict(winSize=15, maxDisparity=20, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    gpu = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True)
    ref = oracle.asw(l, r, stages=True, cost=True, **kw)
    parity.check_cost(gpu["cost"], ref["cost"])
    parity.check_staged(gpu, ref, ref["cost"], ref["cost"], 0, True, max_fraction=None)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_randomised_small_shapes(seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        w, h = int(rng.integers(1, 150)), int(rng.integers(1, 40))
        mind = int(rng.integers(0, 5))
        maxd = mind + int(rng.integers(0, 70))
        win = int(rng.choice([1, 3, 7, 11, 21, 35]))
        cons = bool(rng.integers(0, 2))
        l, r, _ = synth_pair(w, h, maxd, int(rng.integers(0, 1 << 30)))
        kw = dict(winSize=win, maxDisparity=maxd, minDisparity=mind, gammaC=float(rng.uniform(3, 20)), gammaP=float(rng.uniform(5, 30)), consistent=cons)
        gpu = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True)
        ref = oracle.asw(l, r, stages=True, cost=True, **kw)
        parity.check_cost(gpu["cost"], ref["cost"])
        parity.check_staged(gpu, ref, ref["cost"], ref["cost"], mind, cons, max_fraction=0.01)
        wing = int(rng.choice([1, 3, 5, 9]))
        kwg = dict(winSize=wing, maxDisparity=maxd, minDisparity=mind, gamma=int(rng.integers(3, 20)), fMax=float(rng.uniform(30, 200)), iterations=int(rng.integers(0, 4)), bins=20)
        gpu = ss.passive.StereoGSW(**kwg).compute_staged(l, r, cost=True)
        ref = oracle.gsw(l, r, stages=True, cost=True, **kwg)
        parity.check_cost(gpu["cost_left"], ref["cost_left"], "cost_left")
        parity.check_cost(gpu["cost_right"], ref["cost_right"], "cost_right")
        parity.check_staged(gpu, ref, ref["cost_left"], ref["cost_right"], mind, True, max_fraction=0.01, saturation=None)


def test_multi_chunk_disparity_range():
    """D = 300 -> three 128-wide chunks merged through the atomicMin keys."""
    l, r, _ = synth_pair(400, 12, 299, 9)
    kw = dict(winSize=9, maxDisparity=299, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    gpu = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True)
    ref = oracle.asw(l, r, stages=True, cost=True, **kw)
    parity.check_cost(gpu["cost"], ref["cost"])
    parity.check_staged(gpu, ref, ref["cost"], ref["cost"], 0, True)


# ---- BASELINE.json full-size configurations: oracle on a stripe + size-independent properties -----

C2 = dict(winSize=35, maxDisparity=127, minDisparity=0, gammaC=5, gammaP=17.5)


@pytest.fixture(scope="module")
def c2_pair():
    return synth_pair(1242, 375, 127, 0)


@pytest.fixture(scope="module")
def c2_full(c2_pair):
    l, r, _ = c2_pair
    return ss.passive.StereoASW(consistent=True, **C2).compute_staged(l, r)


def test_c2_stripe_against_oracle(c2_pair, c2_full):
    l, r, _ = c2_pair
    for rows in ((0, 4), (180, 192), (371, 375)):
        gpu, ref, nl, nr = _stripe_parity(f"c2_rows{rows[0]}", l, r, dict(consistent=True, **C2), rows, expect_kernel="tc")
        for k in ("left", "right", "invalid", "final"):            # the stripe call is the full-frame call, bit for bit
            assert np.array_equal(gpu[k], c2_full[k][slice(*rows)])


def test_c2_row_stripes_equal_full_frame(c2_pair, c2_full):
    """Row-stripe sharding (the multi-GPU unit) is bit-identical to the full-frame call."""
    l, r, _ = c2_pair
    m = ss.passive.StereoASW(consistent=True, **C2)
    for r0, r1 in ((0, 47), (47, 94), (300, 375), (187, 188)):
        assert np.array_equal(m.compute(l, r, rows=(r0, r1)), c2_full["final"][r0:r1])


def test_c2_recovers_ground_truth(c2_pair, c2_full):
    _, _, gt = c2_pair
    left = c2_full["left"]
    assert (left == gt).mean() > 0.85          # reference: 91.8 % on this pair
    assert left.min() >= 0 and left.max() <= 127
    assert (c2_full["final"] >= 0).all()


def test_c2_disparity_shift_equivariance(c2_pair, c2_full):
    """Size-independent property at the full C2 size: with the right image shifted left by k columns and the search range
    moved up by k (minDisparity = k), every winner moves by exactly k wherever both calls see the same candidates and the same
    window pixels -- x >= maxD + k + pad (no candidate or window column falls off the left edge in either call) and
    x <= W - 1 - pad (none reaches the columns the shift had to invent).  Same pairs, same data, other feature / raw-cost
    offsets (minD > 0 path) at full size: the maps must agree bit for bit."""
    l, r, _ = c2_pair
    k, pad, W = 7, C2["winSize"] // 2, l.shape[1]
    rs = np.empty_like(r)
    rs[:, :-k] = r[:, k:]
    rs[:, -k:] = r[:, -1:]
    kw = dict(C2, minDisparity=k, maxDisparity=C2["maxDisparity"] + k)
    shifted = ss.passive.StereoASW(consistent=False, **kw).compute_staged(l, rs)["left"]
    x0, x1 = C2["maxDisparity"] + k + pad, W - pad
    assert np.array_equal(shifted[:, x0:x1], c2_full["left"][:, x0:x1] + k)
    parity.record("c2_shift_equivariance", pixels=int(shifted[:, x0:x1].size), shift=k, identical=True)


def test_identical_pair_gives_min_disparity():
    """cost(d=minD) is exactly 0 when left == right shifted by minD... here minD = 0: every pixel -> 0."""
    l, _, _ = synth_pair(1242, 64, 127, 4)
    d = ss.passive.StereoASW(consistent=True, **C2).compute(l, l.copy())
    assert (d == 0).all()
    g = ss.passive.StereoGSW(winSize=35, maxDisparity=127, minDisparity=0).compute(l, l.copy())
    assert (g == 0).all()


def test_c3_gsw_stripe_against_oracle(c2_pair):
    l, r, _ = c2_pair
    kw = dict(winSize=35, maxDisparity=127, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)
    rows = (100, 104)
    gpu = ss.passive.StereoGSW(**kw).compute_staged(l, r)
    ref = oracle.gsw(l, r, stages=True, cost=True, rows=rows, **kw)
    sl = slice(*rows)
    g = {k: gpu[k][sl] for k in ("left", "right", "invalid", "final")}
    rr = {k: ref[k][sl] for k in ("left", "right", "invalid", "final")}
    parity.check_staged(g, rr, ref["cost_left"], ref["cost_right"], 0, True, saturation=None)


def test_c4_shape_win51_d256_stripe():
    """Middlebury-full parameters (win 51, 256 disparities, L-R) on a full-width band."""
    l, r, _ = synth_pair(2880, 72, 255, 2)
    kw = dict(winSize=51, maxDisparity=255, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    _stripe_parity("c4_shape_band_win51_d256", l, r, kw, (34, 37), expect_kernel="ws")


# ---- the tensor-core kernel (k_aggregate_tc) on real images, at borders, and at its precision bound ----------------

@pytest.mark.parametrize("win,gc", [(35, 5.0), (35, 17.5), (41, 5.0)])
def test_tensor_core_kernel_on_tsukuba(win, gc):
    """Tsukuba with 128 candidates reaches k_aggregate_tc (its 17-candidate default runs 32-disparity chunks on
    k_aggregate_ws): real texture, real occlusions, sharp colour edges -- both operand-staging variants (win <= 39 / 41)."""
    l, r = cases.load_inputs(("tsukuba", None))
    kw = dict(winSize=win, maxDisparity=127, minDisparity=0, gammaC=gc, gammaP=17.5, consistent=True)
    for rows in ((0, 6), (136, 148), (282, 288)):                     # top border, interior, bottom border
        _stripe_parity(f"tsukuba_d128_win{win}_gc{gc}_rows{rows[0]}", l, r, kw, rows, expect_kernel="tc")


def test_tensor_core_kernel_at_image_borders():
    """Images barely wider than one tile / narrower than the window: every K-group pattern of the truncation
    compensation (window columns clipped left, right, both) and clipped window rows, full cost-volume check."""
    for (w, h, maxd, mind, win, seed) in ((150, 30, 100, 0, 35, 31), (97, 20, 90, 2, 41, 32), (40, 12, 70, 0, 35, 33), (300, 9, 200, 5, 21, 34)):
        l, r, _ = synth_pair(w, h, maxd, seed)
        kw = dict(winSize=win, maxDisparity=maxd, minDisparity=mind, gammaC=5.0, gammaP=17.5, consistent=True)
        _stripe_parity(f"border_{w}x{h}_d{maxd}_win{win}", l, r, kw, (0, h), expect_kernel="tc", max_fraction=0.01)


def test_tensor_core_denominator_bound_on_isolated_pixels():
    """Salt-and-pepper pairs: a pixel unlike all its neighbours has den = 1 + tiny, the MMAs add (almost) nothing and lose
    nothing, so the centred compensation is at its worst (+n * 2^-24); sparse texture on black the other extreme."""
    rng = np.random.default_rng(77)
    for name, p in (("salt", 0.5), ("sparse", 0.03)):
        l = (rng.random((24, 200, 1)) < p).astype(np.uint8).repeat(3, 2) * 255
        r = np.roll(l, -7, axis=1)
        kw = dict(winSize=35, maxDisparity=90, minDisparity=0, gammaC=5.0, gammaP=17.5, consistent=False)
        _stripe_parity(f"isolated_{name}", l, r, kw, (0, 24), expect_kernel="tc", max_fraction=None)


def test_large_windows_run_the_cuda_core_kernel():
    """win 51 / 63 / 79: beyond the window range whose tensor-core truncation bound fits the tolerance (TC_MAX_WIN = 41)."""
    l, r, _ = synth_pair(260, 16, 130, 41)
    for win in (51, 63, 79):
        kw = dict(winSize=win, maxDisparity=130, minDisparity=0, gammaC=7.0, gammaP=30.0, consistent=True)
        _stripe_parity(f"large_window_win{win}", l, r, kw, (4, 10), expect_kernel="ws", max_fraction=0.01)


def test_lab_conversion_stage_matches_the_reference():
    """ColorConversion::ImageFromBGR2Lab (headers/colorconversion.hpp:18-86) on its own: the device Lab image against the
    oracle's float64 restatement, over all 2^24 BGR triples.  The kernels keep Lab in float32: the bar is the float32
    rounding of the reference's double, except where the device's pow and glibc's powf round the cube root to different
    float32 neighbours (one ulp of f, i.e. at most 116 / 500 / 200 * 2^-23 in L / a / b); the exact fraction is recorded."""
    from simplestereo_b200 import _cabi
    v = np.arange(1 << 24, dtype=np.uint32)
    img = np.stack([v & 255, (v >> 8) & 255, v >> 16], axis=1).astype(np.uint8).reshape(4096, 4096, 3)
    gpu = _cabi.lab(img)
    ref = oracle.bgr2lab(img)
    exact = gpu == ref.astype(np.float32)
    err = np.abs(gpu.astype(np.float64) - ref)
    # L = 116 fy - 16, a = 500 (fx - fy), b = 200 (fy - fz) with fx, fy, fz the float32 results of powf (<= 1.09): where the device's
    # double pow lands on the other side of a float32 rounding boundary than glibc's powf, f moves by one float32 ulp (2^-23)
    tol = np.array([116.0, 500.0, 200.0]) * 2.0 ** -23 * 1.01 + np.abs(ref) * 2.0 ** -24
    assert (err <= tol).all(), f"max Lab error {err.max():.3e} (L, a, b: {err.reshape(-1, 3).max(axis=0)})"
    assert exact.mean() > 0.998, f"only {exact.mean():.6f} of the Lab components are the exact float32 rounding"
    parity.record("lab_all_colours", components=int(exact.size), exact_float32_rounding=int(exact.sum()), max_abs_err=float(err.max()))


# ---- API / validation behaviour ------------------------------------------------------

def test_error_behaviour_matches_reference():
    a = np.zeros((8, 9, 3), np.uint8)
    with pytest.raises(ValueError, match="winSize must be a positive odd number!"):
        ss.passive.StereoASW(winSize=4)
    with pytest.raises(TypeError, match="Wrong type input!"):
        ss.passive.StereoASW().compute(a.astype(np.float32), a)
    with pytest.raises(TypeError, match="Wrong type input!"):
        ss.passive.StereoASW().compute(a, a.astype(np.float32))
    with pytest.raises(ValueError, match="Wrong image dimensions!"):
        ss.passive.StereoASW().compute(a[:, :, 0], a[:, :, 0])
    with pytest.raises(ValueError, match="Wrong image dimensions!"):
        ss.passive.StereoASW().compute(a, a[:, :8])
    with pytest.raises(ValueError, match="Invalid input format!"):
        ss.passive.StereoASW().compute(a.tolist(), a)
    with pytest.raises(ValueError, match="Invalid input format!"):
        ss.passive.StereoGSW(gamma=10.5).compute(a, a)
    with pytest.raises(ValueError):
        ss.passive.StereoASW(gammaC=0).compute(a, a)
    with pytest.raises(ValueError):
        ss.passive.StereoASW(minDisparity=-1).compute(a, a)
    m = ss.passive.StereoASW(winSize=3, maxDisparity=2)
    m.winSize = 4                                       # bypass the constructor check, hit the C-side one
    with pytest.raises(ValueError, match="winSize must be a positive odd number!"):
        m.compute(a, a)


def test_non_contiguous_input_is_accepted():
    l, r, _ = synth_pair(64, 20, 8, 1)
    m = ss.passive.StereoASW(winSize=7, maxDisparity=8)
    want = m.compute(l, r)
    lf, rf = np.asfortranarray(l), np.asfortranarray(r)
    assert np.array_equal(m.compute(lf, rf), want)
    big_l = np.zeros((20, 128, 3), np.uint8); big_l[:, ::2] = l
    big_r = np.zeros((20, 128, 3), np.uint8); big_r[:, ::2] = r
    assert np.array_equal(m.compute(big_l[:, ::2], big_r[:, ::2]), want)


# ---- sharding entry points (single GPU emulating N shards; the N-rank path is tests/test_sharding.py + bench.py) ----

def test_disparity_range_shards_merge_equals_unsharded():
    """ss_asw_partial_device over 4 disparity shards + ss_merge_keys_device + ss_finalize_keys_device
    == the unsharded call (config C5's partition, on one GPU)."""
    import torch
    from simplestereo_b200 import _cabi
    from simplestereo_b200.sharding import disparity_shards
    l, r, _ = synth_pair(320, 40, 150, 6)
    h, w = l.shape[:2]
    kw = dict(winSize=21, maxDisparity=150, minDisparity=3, gammaC=5, gammaP=17.5, consistent=True)
    want = ss.passive.StereoASW(**kw).compute(l, r)
    L = _cabi.lib()
    dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    n = 4
    keys = torch.empty((n * 2, h * w), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for k, (d0, d1) in enumerate(disparity_shards(3, 150, n)):
        _cabi.check(L.ss_asw_partial_device(dl.data_ptr(), dr.data_ptr(), w, h, 21, 150, 3, 5.0, 17.5, 1, 0, h, d0, d1,
                                            keys[2 * k].data_ptr(), keys[2 * k + 1].data_ptr(), st))
    _cabi.check(L.ss_merge_keys_device(keys.data_ptr(), n, 2 * h * w, st))
    out = torch.empty((h, w), dtype=torch.int16, device="cuda")
    _cabi.check(L.ss_finalize_keys_device(keys[0].data_ptr(), keys[1].data_ptr(), w, h, 3, out.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)


def test_disparity_shards_of_a_near_tie_stress_pair_equal_unsharded():
    """Saturated i.i.d. noise: every argmin is a rounding tie, so any difference in arithmetic between the sharded and
    the unsharded call would show.  Shards that start inside a 128-disparity chunk run the same kernel on the same
    chunk grid as the whole call (include/ss_passive.h), hence bit-identical keys."""
    import torch
    from simplestereo_b200 import _cabi
    from simplestereo_b200.sharding import disparity_shards
    l, r = cases.load_inputs(("noise", (200, 24, 9)))
    h, w = l.shape[:2]
    for (mind, maxd, win) in ((0, 100, 15), (3, 200, 9)):
        want = ss.passive.StereoASW(win, maxd, mind, 5.0, 17.5, True).compute(l, r)
        assert _cabi.last_kernel()[0] == "tc"
        L = _cabi.lib()
        dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
        st = torch.cuda.current_stream().cuda_stream
        for n in (3, 8):
            keys = torch.empty((n * 2, h * w), dtype=torch.int64, device="cuda")
            for k, (d0, d1) in enumerate(disparity_shards(mind, maxd, n)):
                _cabi.check(L.ss_asw_partial_device(dl.data_ptr(), dr.data_ptr(), w, h, win, maxd, mind, 5.0, 17.5, 1, 0, h, d0, d1,
                                                    keys[2 * k].data_ptr(), keys[2 * k + 1].data_ptr(), st))
            _cabi.check(L.ss_merge_keys_device(keys.data_ptr(), n, 2 * h * w, st))
            out = torch.empty((h, w), dtype=torch.int16, device="cuda")
            _cabi.check(L.ss_finalize_keys_device(keys[0].data_ptr(), keys[1].data_ptr(), w, h, mind, out.data_ptr(), st))
            torch.cuda.synchronize()
            assert np.array_equal(out.cpu().numpy(), want), (mind, maxd, n)


def test_calls_on_different_streams_do_not_race_on_the_cached_scratch():
    """The device entry points share the per-device scratch; a call on another stream must wait for the previous one."""
    import torch
    from simplestereo_b200 import _cabi
    L = _cabi.lib()
    pairs = [synth_pair(400, 60, 100, 50 + k)[:2] for k in range(4)]
    m = ss.passive.StereoASW(21, 100, 0, 5.0, 17.5, True)
    want = [m.compute(l, r) for l, r in pairs]
    dev = [(torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()) for l, r in pairs]
    outs = [torch.empty((60, 400), dtype=torch.int16, device="cuda") for _ in pairs]
    streams = [torch.cuda.Stream() for _ in pairs]
    torch.cuda.synchronize()
    for rep in range(3):
        for (dl, dr), o, s in zip(dev, outs, streams):
            _cabi.check(L.ss_asw_compute_device(dl.data_ptr(), dr.data_ptr(), 400, 60, *m._args(), 0, 60, o.data_ptr(), s.cuda_stream))
        g = ss.passive.StereoGSW(9, 100).compute(*pairs[0])           # a host call on the library's own stream in between
        torch.cuda.synchronize()
        for o, w_ in zip(outs, want):
            assert np.array_equal(o.cpu().numpy(), w_)
        assert g.shape == (60, 400)


def test_devices_kwarg_shards_rows_over_gpus_in_process():
    """StereoASW(devices=[0, 1]).compute == one GPU, bit for bit (ss_init_devices: one host thread + context per GPU)."""
    import torch
    from simplestereo_b200 import _cabi
    n = torch.cuda.device_count()
    l, r, _ = synth_pair(500, 90, 100, 61)
    kw = dict(winSize=21, maxDisparity=100, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    _cabi.use_devices([0])
    want = ss.passive.StereoASW(**kw).compute(l, r)
    wantg = ss.passive.StereoGSW(9, 100).compute(l, r)
    try:
        devs = list(range(min(n, 4)))
        assert np.array_equal(ss.passive.StereoASW(devices=devs, **kw).compute(l, r), want)
        assert _cabi.lib().ss_device_count() == len(devs)
        assert np.array_equal(ss.passive.StereoGSW(9, 100, devices=devs).compute(l, r), wantg)
        assert np.array_equal(ss.passive.StereoASW(devices="all", **kw).compute(l, r, rows=(7, 80)), want[7:80])
        if n >= 2:
            # device-resident form: one grouped in-place ncclAllGather leaves the whole map on every device
            import ctypes
            devs = list(range(n))
            _cabi.init_devices(devs)
            S = -(-90 // n)
            dl = [torch.from_numpy(l).to(f"cuda:{k}") for k in devs]
            dr = [torch.from_numpy(r).to(f"cuda:{k}") for k in devs]
            do = [torch.zeros((n * S, 500), dtype=torch.int16, device=f"cuda:{k}") for k in devs]
            arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
            win, maxd, mind, gc, gp, cons = ss.passive.StereoASW(**kw)._args()
            _cabi.check(_cabi.lib().ss_asw_compute_multi_device(arr(dl), arr(dr), 500, 90, win, maxd, mind, gc, gp, cons, arr(do), None))
            for o in do:
                assert np.array_equal(o.cpu().numpy()[:90], want)
    finally:
        _cabi.use_devices([0])


def test_tail_wave_split_is_bit_identical():
    """A launch whose last wave would fill only a few SMs runs those tail tiles as 32-column sub-blocks (launch_agg in
    csrc/ss_passive.cu).  150 tiles on 148 SMs -> 2 tail tiles x 3 sub-blocks; the two half-frame stripe calls (75 tiles each,
    no split) must give the same bits, and the split call must pass the oracle check, cost volume included."""
    l, r, _ = synth_pair(200, 50, 100, 71)
    kw = dict(winSize=21, maxDisparity=100, minDisparity=0, gammaC=5.0, gammaP=17.5, consistent=True)
    m = ss.passive.StereoASW(**kw)
    full = m.compute_staged(l, r, cost=True)
    ref = oracle.asw(l, r, stages=True, cost=True, **kw)
    parity.check_cost(full["cost"], ref["cost"])
    parity.check_staged(full, ref, ref["cost"], ref["cost"], 0, True, max_fraction=0.01)
    halves = [m.compute_staged(l, r, cost=True, rows=rr) for rr in ((0, 25), (25, 50))]
    for k in ("left", "right", "invalid", "final", "cost"):
        assert np.array_equal(np.concatenate([h[k] for h in halves]), full[k]), k
    # GSW (64-column tiles, 2 sub-blocks): 4 x 38 = 152 tiles
    lg, rg, _ = synth_pair(200, 38, 100, 72)
    g = ss.passive.StereoGSW(9, 100)
    fullg = g.compute_staged(lg, rg, cost=True)
    halvesg = [g.compute_staged(lg, rg, cost=True, rows=rr) for rr in ((0, 19), (19, 38))]
    for k in ("left", "right", "invalid", "final", "cost_left", "cost_right"):
        assert np.array_equal(np.concatenate([h[k] for h in halvesg]), fullg[k]), k


def test_row_stripes_device_entry_point():
    import torch
    from simplestereo_b200 import _cabi
    from simplestereo_b200.sharding import row_stripes
    l, r, _ = synth_pair(200, 37, 30, 8)
    h, w = l.shape[:2]
    m = ss.passive.StereoGSW(winSize=9, maxDisparity=30)
    want = m.compute(l, r)
    L = _cabi.lib()
    dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    st = torch.cuda.current_stream().cuda_stream
    got = torch.zeros((h, w), dtype=torch.int16, device="cuda")
    for r0, r1 in row_stripes(h, 3):
        _cabi.check(L.ss_gsw_compute_device(dl.data_ptr(), dr.data_ptr(), w, h, *m._args(), r0, r1, got[r0:r1].data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(got.cpu().numpy(), want)


# ---- BASELINE.json configs C4 / C5 at FULL size: size-independent properties + oracle rows ------------------

def test_c4_full_size_middlebury_lr_consistency():
    """C4: 2880x1988, 256 disparities, win 51, L-R check.  Full frame on the GPU; three rows against the oracle
    (which reads the full-height window), row stripes bit-identical to the full-frame call, ground truth recovered."""
    l, r, gt = synth_pair(2880, 1988, 255, 2)
    kw = dict(winSize=51, maxDisparity=255, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    m = ss.passive.StereoASW(**kw)
    gpu = m.compute_staged(l, r)
    for rows in ((1000, 1002), (0, 1), (1987, 1988)):
        g, ref, nl, nr = _stripe_parity(f"c4_full_rows{rows[0]}", l, r, kw, rows, expect_kernel="ws")
        for k in ("left", "right", "invalid", "final"):
            assert np.array_equal(g[k], gpu[k][slice(*rows)])
    for r0, r1 in ((0, 3), (994, 1003), (1985, 1988)):
        assert np.array_equal(m.compute(l, r, rows=(r0, r1)), gpu["final"][r0:r1])
    assert (gpu["left"] == gt).mean() > 0.85
    assert gpu["final"].min() >= 0 and gpu["final"].max() <= 255


def test_c5_full_size_4k_disparity_shards_equal_unsharded():
    """C5: 3840x2160, 512 disparities.  Eight disparity-range shards (the 8-GPU partition, emulated on one GPU
    through ss_asw_partial_device / ss_merge_keys_device / ss_finalize_keys_device) == the unsharded call,
    bit for bit; two rows against the oracle; ground truth recovered."""
    import torch
    from simplestereo_b200 import _cabi
    from simplestereo_b200.sharding import disparity_shards
    l, r, gt = synth_pair(3840, 2160, 511, 3)
    h, w = l.shape[:2]
    kw = dict(winSize=35, maxDisparity=511, minDisparity=0, gammaC=5, gammaP=17.5, consistent=False)
    m = ss.passive.StereoASW(**kw)
    want = m.compute(l, r)
    L = _cabi.lib()
    dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
    n = 8
    keys = torch.empty((n, h * w), dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for k, (d0, d1) in enumerate(disparity_shards(0, 511, n)):
        _cabi.check(L.ss_asw_partial_device(dl.data_ptr(), dr.data_ptr(), w, h, 35, 511, 0, 5.0, 17.5, 0, 0, h, d0, d1,
                                            keys[k].data_ptr(), None, st))
    _cabi.check(L.ss_merge_keys_device(keys.data_ptr(), n, h * w, st))
    out = torch.empty((h, w), dtype=torch.int16, device="cuda")
    _cabi.check(L.ss_finalize_keys_device(keys[0].data_ptr(), None, w, h, 0, out.data_ptr(), st))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), want)
    del keys, out, dl, dr
    torch.cuda.empty_cache()
    for rows in ((1080, 1081), (2159, 2160)):
        g, ref, nl, nr = _stripe_parity(f"c5_full_rows{rows[0]}", l, r, kw, rows, expect_kernel="tc")
        assert np.array_equal(g["final"], want[slice(*rows)])
    assert (want == gt).mean() > 0.80
    assert want.min() >= 0 and want.max() <= 511


def test_gsw_large_window_runs_narrower_chunks():
    """win 51: the float raw-cost tiles of a 128-disparity chunk do not fit the double-buffered staging of k_aggregate_ws;
    the call runs 64-disparity chunks instead (merged through the atomicMin keys).  Same parity bar."""
    from simplestereo_b200 import _cabi
    l, r, _ = synth_pair(200, 20, 100, 12)
    kw = dict(winSize=51, maxDisparity=100, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20)
    gpu = ss.passive.StereoGSW(**kw).compute_staged(l, r, cost=True)
    assert _cabi.last_kernel() == ("ws", 64)
    ref = oracle.gsw(l, r, stages=True, cost=True, **kw)
    parity.check_cost(gpu["cost_left"], ref["cost_left"], "cost_left")
    parity.check_cost(gpu["cost_right"], ref["cost_right"], "cost_right")
    parity.check_staged(gpu, ref, ref["cost_left"], ref["cost_right"], 0, True, max_fraction=0.01, saturation=None)


def test_c_abi_from_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: a plain-C program (tests/c_abi_smoke.c, compiled here with gcc against
    include/*.h and libsspassive.so) must produce the same bytes as the Python mirror."""
    import shutil
    import subprocess
    from simplestereo_b200 import _cabi
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = tmp_path / "c_abi_smoke"
    libdir = os.path.dirname(_cabi.LIB_PATH)
    subprocess.run(["gcc", "-O1", "-o", str(exe), os.path.join(here, "c_abi_smoke.c"), "-L" + libdir, "-lsspassive",
                    "-Wl,-rpath," + libdir], check=True)
    l, r, _ = synth_pair(150, 40, 24, 21)
    l.tofile(tmp_path / "l.bgr")
    r.tofile(tmp_path / "r.bgr")
    out = subprocess.run([str(exe), str(tmp_path / "l.bgr"), str(tmp_path / "r.bgr"), "150", "40", str(tmp_path / "o")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    asw = np.fromfile(tmp_path / "o.asw.i16", np.int16).reshape(40, 150)
    gsw = np.fromfile(tmp_path / "o.gsw.i16", np.int16).reshape(40, 150)
    pts = np.fromfile(tmp_path / "o.pts.f32", np.float32).reshape(40, 150, 3)
    assert np.array_equal(asw, ss.passive.StereoASW(9, 24, 0, 5.0, 17.5, True).compute(l, r))
    assert np.array_equal(gsw, ss.passive.StereoGSW(7, 24, 0, 10, 120.0, 3, 20).compute(l, r))
    d = ss.passive.StereoASW(9, 24, 0, 5.0, 17.5, False).compute(l, r)
    assert np.array_equal(pts, ss.points.getAdimensional3DPoints(d), equal_nan=True)


def test_tensor_core_and_cuda_core_aggregation_agree(monkeypatch):
    """ASW with 128-disparity chunks runs k_aggregate_tc (denominators on tcgen05, 3xTF32 + centred truncation bound; win 41 is
    its single-staged operand variant); SS_TCDEN=0 forces the all-CUDA-core k_aggregate_ws.  Same parity bar for both, and
    they must agree with each other: costs to 5e-5, maps except at near-ties."""
    from simplestereo_b200 import _cabi
    l, r, _ = synth_pair(330, 44, 139, 17)
    for win in (35, 39, 41):
        kw = dict(winSize=win, maxDisparity=139, minDisparity=0, gammaC=9.0, gammaP=25.0, consistent=True)
        ref = oracle.asw(l, r, stages=True, cost=True, **kw)
        out = {}
        for name, env, want in (("tc", {}, "k_aggregate_tc"), ("ws", {"SS_TCDEN": "0"}, "k_aggregate_ws")):
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            out[name] = ss.passive.StereoASW(**kw).compute_staged(l, r, cost=True)
            assert _cabi.last_kernel_name() == want
            for k in env:
                monkeypatch.delenv(k)
            parity.check_cost(out[name]["cost"], ref["cost"])
            parity.check_staged(out[name], ref, ref["cost"], ref["cost"], 0, True, max_fraction=0.01)
        fin = np.isfinite(out["ws"]["cost"])
        assert np.array_equal(fin, np.isfinite(out["tc"]["cost"]))
        assert np.allclose(out["tc"]["cost"][fin], out["ws"]["cost"][fin], rtol=5e-5, atol=1e-5)
        # where the two maps differ, the two picks are a near-tie of the oracle's own costs
        parity.adjudicate_left(out["tc"]["left"], out["ws"]["left"], ref["cost"], 0, max_fraction=None)
