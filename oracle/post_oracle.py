"""
post_oracle -- CPU restatements (numpy) of the steps either side of the ASW / GSW hot path, the rows
SURVEY.md section 8(f) ranks "next".  TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/: only
tests/, __graft_entry__.smoke() and bench legs may import it; the product never does).

The reference implements none of this arithmetic itself -- it calls OpenCV (not vendored in /root/reference;
the image has opencv-python 4.13, the reference's setup.py pins no version):

  reproject            cv2.reprojectImageTo3D(disparityMap, Q)       simplestereo/points.py:176, _rigs.py:628
  adimensional_q       the Q matrix of points.getAdimensional3DPoints   simplestereo/points.py:147-174
  rig_q                the Q matrix of RectifiedStereoRig.get3DPoints    simplestereo/_rigs.py:604-625
  normalize_colormap   cv2.normalize(.., 0, 255, NORM_MINMAX, CV_8UC1) + cv2.applyColorMap(.., COLORMAP_JET)
                                                                     examples/010 StereoMatchingTsukuba.py:44-45
  remap_linear         cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR)  simplestereo/_rigs.py:564-565

Each function restates OpenCV's published algorithm for exactly the argument types the reference passes
(int16 disparity, float64 Q, uint8 BGR images, float32 maps, BORDER_CONSTANT 0).

Parity status: PINNED -- tests/test_post_oracle.py checks every function bit-exactly against outputs of cv2
itself (tests/golden/post_outputs.npz, generated in the build container by tests/golden/make_golden_post.py,
plus live cv2 when it is importable).
"""
import numpy as np


def adimensional_q(width, height):
    """Q of points.getAdimensional3DPoints (points.py:147-174): b=1, fx=fy=width, cx1=cx2=width/2, cy=height/2."""
    return rig_q(b=1, fx=width, fy=width, cx1=width / 2, cx2=width / 2, a1=0, a2=0, cy=height / 2)


def rig_q(b, fx, fy, cx1, cx2, a1, a2, cy):
    """Q of RectifiedStereoRig.get3DPoints (_rigs.py:604-625)."""
    Q = np.eye(4, dtype="float64")
    Q[0, 1] = -a1 / fy
    Q[0, 3] = a1 * cy / fy - cx1
    Q[1, 1] = fx / fy
    Q[1, 3] = -cy * fx / fy
    Q[2, 2] = 0
    Q[2, 3] = -fx
    Q[3, 1] = (a2 - a1) / (fy * b)
    Q[3, 2] = 1 / b
    Q[3, 3] = ((a1 - a2) * cy + (cx2 - cx1) * fy) / (fy * b)
    return Q


def reproject(disparity, Q):
    """cv2.reprojectImageTo3D(int16 disparity, Q) -> float32 [H, W, 3] (handleMissingValues=False).

    OpenCV 4.x: homogeneous point = Q * (x, y, d, 1) in double (row sums left to right, unfused), the first
    three components are rounded to float32, then each is multiplied by the double reciprocal 1/w (Vec3f /= w is
    implemented as *= 1/w) and rounded again.  w == 0 gives +-inf / nan exactly as IEEE arithmetic does."""
    disparity = np.asarray(disparity)
    Q = np.asarray(Q, dtype=np.float64)
    H, W = disparity.shape
    y, x = np.mgrid[0:H, 0:W].astype(np.float64)
    d = disparity.astype(np.float64)
    out = np.empty((H, W, 3), np.float32)
    w = ((Q[3, 0] * x + Q[3, 1] * y) + Q[3, 2] * d) + Q[3, 3] * 1.0
    with np.errstate(all="ignore"):
        for k in range(3):
            X = (((Q[k, 0] * x + Q[k, 1] * y) + Q[k, 2] * d) + Q[k, 3] * 1.0).astype(np.float32)
            out[..., k] = (X.astype(np.float64) * (1.0 / w)).astype(np.float32)
    return out


def normalize_minmax_u8(disparity):
    """cv2.normalize(d, None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8UC1) for an int16 map.

    scale = 255 * (1 / (max - min)) (0 when max == min), shift = -min * scale in double; convertTo evaluates
    float(src) * float(scale) + float(shift) with ONE rounding (fused multiply-add on every FMA-capable host,
    which is what the golden vectors were generated on), rounds half to even and saturates."""
    d = np.asarray(disparity)
    smin, smax = float(d.min()), float(d.max())
    scale = 255.0 * (1.0 / (smax - smin) if smax - smin > np.finfo(np.float64).eps else 0.0)
    shift = 0.0 - smin * scale
    a, b = np.float64(np.float32(scale)), np.float64(np.float32(shift))
    f = (d.astype(np.float64) * a + b).astype(np.float32)          # exact product (int16 x float32) + one rounding
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def normalize_colormap(disparity, lut_bgr):
    """examples/010:44-45: min-max normalise to uint8, then a 256-entry BGR look-up table (COLORMAP_JET)."""
    g = normalize_minmax_u8(disparity)
    return g, np.asarray(lut_bgr, np.uint8)[g]


def remap_linear(src, mapx, mapy):
    """cv2.remap(src uint8 [h, w, 3], mapx, mapy float32 [H, W], cv2.INTER_LINEAR), BORDER_CONSTANT value 0.

    OpenCV converts the float maps to fixed point with 5 fractional bits (cvRound(v * 32), round half to even,
    non-finite / out-of-int-range -> INT_MIN as cvtss2si does), clamps the integer part to int16, and blends
    the four neighbours with 15-bit weights (32-fx)(32-fy)*32 ... that sum to 2^15; result = (acc + 2^14) >> 15."""
    src = np.asarray(src)
    sh, sw = src.shape[:2]

    def cvround(v):
        f = v.astype(np.float32) * np.float32(32)
        r = np.rint(f.astype(np.float64))
        bad = ~np.isfinite(r) | (np.abs(r) >= 2.0 ** 31)
        return np.where(bad, -2.0 ** 31, r).astype(np.int64)

    sx, sy = cvround(np.asarray(mapx)), cvround(np.asarray(mapy))
    ax, ay = sx & 31, sy & 31
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    w00, w01 = (32 - ax) * (32 - ay) * 32, ax * (32 - ay) * 32
    w10, w11 = (32 - ax) * ay * 32, ax * ay * 32

    def fetch(yy, xx):
        ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return np.where(ok[..., None], v, 0)

    acc = (fetch(iy, ix) * w00[..., None] + fetch(iy, ix + 1) * w01[..., None]
           + fetch(iy + 1, ix) * w10[..., None] + fetch(iy + 1, ix + 1) * w11[..., None])
    return np.clip((acc + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
