"""
oracle -- CPU parity oracle for the ASW / GSW hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``simplestereo_b200``) never imports it and has no CPU fallback.

Two checkers live here:

* ``oracle.asw`` / ``oracle.gsw``  -- ctypes front-end of ``passive_oracle.c`` (our plain-C
  restatement of /root/reference/simplestereo/_passive.cpp, staged outputs + cost volumes,
  OpenMP over rows).  kind = "port".
* ``oracle.ref_asw`` / ``oracle.ref_gsw`` -- the UNMODIFIED reference extension compiled by
  ``make -C oracle ref`` into ``oracle/_ref`` (git-ignored, travels to the GPU box), run in a
  subprocess with a timeout because the reference leaks every buffer and can hang at the
  tail of its job queue (_passive.cpp:29-32 + headers/safequeue.hpp:106-115).  kind = "reference".

Parity status: PINNED -- see tests/test_oracle.py (Tsukuba known-answer image + fixtures
generated from oracle/_ref by tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes
import glob
import os
import pickle
import subprocess
import sys
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REF_DIR = os.path.join(_HERE, "_ref")
_lib = None


def build(ref: bool = True) -> None:
    """Compile the C restatement (and, if /root/reference is present, oracle/_ref)."""
    targets = ["all"]
    if ref and os.path.exists("/root/reference/simplestereo/_passive.cpp"):
        targets.append("ref")
    subprocess.run(["make", "-C", _HERE] + targets, check=True, capture_output=True)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build(ref=False)
    lib = ctypes.CDLL(_LIB_PATH)
    u8p = ctypes.POINTER(ctypes.c_uint8)
    i16p = ctypes.POINTER(ctypes.c_int16)
    f64p = ctypes.POINTER(ctypes.c_double)
    f32p = ctypes.POINTER(ctypes.c_float)
    ci = ctypes.c_int
    lib.orc_asw.argtypes = [u8p, u8p, ci, ci, ci, ci, ci, ctypes.c_double, ctypes.c_double,
                            ci, ci, ci, ci, ci, i16p, i16p, i16p, u8p, f64p]
    lib.orc_asw.restype = ci
    lib.orc_gsw.argtypes = [u8p, u8p, ci, ci, ci, ci, ci, ci, ctypes.c_float, ci, ci,
                            ci, ci, ci, ci, i16p, i16p, i16p, u8p, f32p, f32p]
    lib.orc_gsw.restype = ci
    lib.orc_bgr2lab.argtypes = [u8p, f64p, ci, ci]
    lib.orc_bgr2lab.restype = None
    lib.orc_num_threads.restype = ci
    _lib = lib
    return lib


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def _check_pair(img1, img2):
    img1 = np.ascontiguousarray(img1)
    img2 = np.ascontiguousarray(img2)
    assert img1.dtype == np.uint8 and img2.dtype == np.uint8
    assert img1.ndim == 3 and img1.shape == img2.shape and img1.shape[2] == 3
    return img1, img2


def num_threads() -> int:
    return int(_load().orc_num_threads())


def bgr2lab(img):
    img = np.ascontiguousarray(img)
    h, w, _ = img.shape
    out = np.empty((h, w, 3), np.float64)
    _load().orc_bgr2lab(_ptr(img, ctypes.c_uint8), _ptr(out, ctypes.c_double), w, h)
    return out


def asw(img1, img2, winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5,
        consistent=False, *, stages=False, cost=False, rows=None, literal_right=False, nthreads=0):
    """C restatement of _passive.computeASW (_passive.cpp:293-400).

    Returns the int16 disparity map, or with ``stages=True`` a dict with keys
    final/left/right/invalid (+ ``cost``: float64 [rows, W, D] when ``cost=True``).
    ``rows=(r0, r1)`` restricts the computed rows (other rows of the maps are left at 0).
    """
    img1, img2 = _check_pair(img1, img2)
    h, w, _ = img1.shape
    r0, r1 = (0, h) if rows is None else rows
    D = max(maxDisparity - minDisparity + 1, 0)
    final = np.zeros((h, w), np.int16)
    left = np.zeros((h, w), np.int16) if stages else None
    right = np.zeros((h, w), np.int16) if stages else None
    invalid = np.zeros((h, w), np.uint8) if stages else None
    vol = np.empty((r1 - r0, w, D), np.float64) if cost else None
    rc = _load().orc_asw(_ptr(img1, ctypes.c_uint8), _ptr(img2, ctypes.c_uint8), w, h,
                         int(winSize), int(maxDisparity), int(minDisparity), float(gammaC), float(gammaP),
                         int(bool(consistent)), int(bool(literal_right)), int(r0), int(r1), int(nthreads),
                         _ptr(final, ctypes.c_int16), _ptr(left, ctypes.c_int16), _ptr(right, ctypes.c_int16),
                         _ptr(invalid, ctypes.c_uint8), _ptr(vol, ctypes.c_double))
    if rc != 0:
        raise ValueError(f"orc_asw failed with code {rc}")
    if not stages and not cost:
        return final
    out = {"final": final, "left": left, "right": right, "invalid": invalid}
    if cost:
        out["cost"] = vol
    return out


def gsw(img1, img2, winSize=11, maxDisparity=16, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20,
        *, stages=False, cost=False, rows=None, literal=False, nthreads=0):
    """C restatement of _passive.computeGSW (_passive.cpp:703-774); ``literal=True`` runs the
    O(win^4) relaxation of the reference instead of its closed form."""
    img1, img2 = _check_pair(img1, img2)
    h, w, _ = img1.shape
    r0, r1 = (0, h) if rows is None else rows
    D = max(maxDisparity - minDisparity + 1, 0)
    final = np.zeros((h, w), np.int16)
    left = np.zeros((h, w), np.int16) if stages else None
    right = np.zeros((h, w), np.int16) if stages else None
    invalid = np.zeros((h, w), np.uint8) if stages else None
    vl = np.empty((r1 - r0, w, D), np.float32) if cost else None
    vr = np.empty((r1 - r0, w, D), np.float32) if cost else None
    rc = _load().orc_gsw(_ptr(img1, ctypes.c_uint8), _ptr(img2, ctypes.c_uint8), w, h,
                         int(winSize), int(maxDisparity), int(minDisparity), int(gamma), float(fMax),
                         int(iterations), int(bins), int(bool(literal)), int(r0), int(r1), int(nthreads),
                         _ptr(final, ctypes.c_int16), _ptr(left, ctypes.c_int16), _ptr(right, ctypes.c_int16),
                         _ptr(invalid, ctypes.c_uint8), _ptr(vl, ctypes.c_float), _ptr(vr, ctypes.c_float))
    if rc != 0:
        raise ValueError(f"orc_gsw failed with code {rc}")
    if not stages and not cost:
        return final
    out = {"final": final, "left": left, "right": right, "invalid": invalid}
    if cost:
        out["cost_left"] = vl
        out["cost_right"] = vr
    return out


# ----------------------------------------------------------------------------------------------
# the unmodified reference, compiled into oracle/_ref
# ----------------------------------------------------------------------------------------------

def ref_available() -> bool:
    return bool(glob.glob(os.path.join(_REF_DIR, "_passive*.so")))


_REF_DRIVER = r"""
import sys, pickle, time
sys.path.insert(0, {ref_dir!r})
import _passive
with open(sys.argv[1], "rb") as f:
    fn, args = pickle.load(f)
t0 = time.perf_counter()
out = getattr(_passive, fn)(*args)
dt = time.perf_counter() - t0
with open(sys.argv[2], "wb") as f:
    pickle.dump((out, dt), f)
"""


def _ref_call(fn, args, timeout, taskset=None):
    if not ref_available():
        raise RuntimeError("oracle/_ref is not built (make -C oracle ref needs /root/reference)")
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.pkl"), os.path.join(td, "out.pkl")
        with open(fin, "wb") as f:
            pickle.dump((fn, args), f)
        cmd = [sys.executable, "-c", _REF_DRIVER.format(ref_dir=_REF_DIR), fin, fout]
        if taskset:
            cmd = ["taskset", "-c", taskset] + cmd
        subprocess.run(cmd, check=True, timeout=timeout)
        with open(fout, "rb") as f:
            return pickle.load(f)


def ref_asw(img1, img2, winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=False,
            *, timeout=900, with_time=False, taskset=None):
    """_passive.computeASW of the compiled reference, in a subprocess (GIL-holding, leaky, may hang)."""
    img1, img2 = _check_pair(img1, img2)
    out, dt = _ref_call("computeASW", (img1, img2, int(winSize), int(maxDisparity), int(minDisparity),
                                       float(gammaC), float(gammaP), bool(consistent)), timeout, taskset)
    return (out, dt) if with_time else out


def ref_gsw(img1, img2, winSize=11, maxDisparity=16, minDisparity=0, gamma=10, fMax=120, iterations=3, bins=20,
            *, timeout=900, with_time=False, taskset=None):
    img1, img2 = _check_pair(img1, img2)
    out, dt = _ref_call("computeGSW", (img1, img2, int(winSize), int(maxDisparity), int(minDisparity),
                                       int(gamma), float(fMax), int(iterations), int(bins)), timeout, taskset)
    return (out, dt) if with_time else out
