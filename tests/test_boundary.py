"""CPU-side tests of the drop-in boundary: the C-ABI library loads and exports every declared symbol,
the Python mirror validates like the reference, and nothing silently falls back to the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import simplestereo_b200 as ss
from simplestereo_b200 import _cabi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "ss_passive.h")).read() + open(os.path.join(ROOT, "include", "ss_post.h")).read()
    declared = set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), f"{s} not exported by libsspassive.so"
    assert _cabi.lib().ss_abi_version() == 2


def test_library_is_sm100a_and_uses_tma_and_packed_fp32():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run(["cuobjdump", "-lelf", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass          # cp.async.bulk (TMA) staging of window rows
    assert "FFMA2" in sass           # packed fp32 accumulation
    assert "UTCHMMA" in sass         # tcgen05.mma: the ASW denominators (k_aggregate_tc)
    assert "STTM" in sass and "LDTM" in sass   # tcgen05.st / tcgen05.ld: right weights into, denominators out of TMEM


def test_python_side_validation_without_gpu():
    a = np.zeros((8, 9, 3), np.uint8)
    with pytest.raises(ValueError, match="winSize must be a positive odd number!"):
        ss.passive.StereoASW(winSize=4)
    with pytest.raises(ValueError, match="winSize must be a positive odd number!"):
        ss.passive.StereoGSW(winSize=0)
    with pytest.raises(TypeError, match="Wrong type input!"):
        ss.passive.StereoASW().compute(a.astype(np.float32), a)
    with pytest.raises(ValueError, match="Wrong image dimensions!"):
        ss.passive.StereoGSW().compute(a, a[:4])
    with pytest.raises(ValueError, match="Invalid input format!"):
        ss.passive.StereoASW().compute([1, 2], a)
    m = ss.passive.StereoASW()
    assert (m.winSize, m.maxDisparity, m.minDisparity, m.gammaC, m.gammaP, m.consistent) == (35, 16, 0, 5, 17.5, False)
    g = ss.passive.StereoGSW()
    assert (g.winSize, g.maxDisparity, g.minDisparity, g.gamma, g.fMax, g.iterations, g.bins) == (11, 16, 0, 10, 120, 3, 20)


@pytest.mark.skipif(_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    a = np.zeros((8, 9, 3), np.uint8)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        ss.passive.StereoASW(winSize=3, maxDisparity=2).compute(a, a)
    d = np.zeros((8, 9), np.int16)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        ss.points.getAdimensional3DPoints(d)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        ss.display.applyColorMap(d)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        ss.rectify.remap(a, np.zeros((4, 5), np.float32), np.zeros((4, 5), np.float32))


def test_post_path_validation_without_gpu():
    d = np.zeros((8, 9), np.int16)
    with pytest.raises(TypeError, match="Wrong type input!"):
        ss.points.getAdimensional3DPoints(d.astype(np.float32))
    with pytest.raises(ValueError):
        ss.points.reprojectImageTo3D(d, np.eye(3))
    with pytest.raises(ValueError, match="Wrong image dimensions!"):
        ss.display.applyColorMap(np.zeros((2, 3, 3), np.int16))
    with pytest.raises(TypeError, match="Wrong type input!"):
        ss.rectify.remap(np.zeros((8, 9, 3), np.float32), np.zeros((4, 5), np.float32), np.zeros((4, 5), np.float32))
    # the Q of points.getAdimensional3DPoints (points.py:147-174)
    Q = ss.points.buildQ(b=1, fx=384, fy=384, cx1=192, cx2=192, a1=0, a2=0, cy=144)
    assert np.array_equal(Q, np.array([[1, 0, 0, -192.0], [0, 1, 0, -144.0], [0, 0, 0, -384.0], [0, 0, 1.0, 0]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "simplestereo_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_headers_are_plain_c():
    """The boundary is a C ABI: both headers must compile as C (no C++ constructs, no torch / numpy types) and the
    plain-C caller used on the GPU box must at least parse and link against the declared prototypes."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    for h in ("ss_passive.h", "ss_post.h"):
        src = f'#include "{os.path.join(ROOT, "include", h)}"\nint main(void) {{ return SS_OK; }}\n' if h == "ss_passive.h" else \
              f'#include "{os.path.join(ROOT, "include", h)}"\nint main(void) {{ return 0; }}\n'
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", "-"], input=src,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-fsyntax-only", os.path.join(ROOT, "tests", "c_abi_smoke.c")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
