"""
points
======
B200-native counterpart of the 3-D reprojection step that follows the stereo matcher in SimpleStereo
(reference: simplestereo/points.py:124-176 ``getAdimensional3DPoints`` and simplestereo/_rigs.py:569-628
``RectifiedStereoRig.get3DPoints``).  Both build a 4x4 ``Q`` and call ``cv2.reprojectImageTo3D(disparityMap, Q)``;
here the same arithmetic (bit for bit, oracle/post_oracle.py) runs as a CUDA kernel behind ``include/ss_post.h``.
numpy in / numpy out, no CPU fallback.
"""
import numpy as np

from . import _cabi


def buildQ(b, fx, fy, cx1, cx2, a1, a2, cy):
    """The Q matrix the reference hands to OpenCV (_rigs.py:604-625; points.py:147-174 uses the same formulas)."""
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


def _check_disparity(disparityMap):
    if not isinstance(disparityMap, np.ndarray) or disparityMap.ndim != 2:
        raise ValueError("Wrong image dimensions!")
    if disparityMap.dtype != np.int16:
        # the matchers return int16 (passive.py:91); other dtypes take a different OpenCV code path
        raise TypeError("Wrong type input!")
    return np.ascontiguousarray(disparityMap)


def reprojectImageTo3D(disparityMap, Q):
    """``cv2.reprojectImageTo3D(disparityMap, Q)`` for an int16 disparity map: float32 array (height, width, 3)."""
    d = _check_disparity(disparityMap)
    Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64))
    if Q.shape != (4, 4):
        raise ValueError("Q must be a 4x4 matrix")
    h, w = d.shape
    out = np.empty((h, w, 3), np.float32)
    _cabi.check(_cabi.lib().ss_reproject(_cabi.ptr(d), w, h, _cabi.ptr(Q), _cabi.ptr(out)))
    return out


def getAdimensional3DPoints(disparityMap):
    """
    Get adimensional 3D points from the disparity map (reference: simplestereo/points.py:124-176).

    Returns numpy.ndarray of shape *(height,width,3)*, where at each y,x coordinates a *(x,y,z)* point is associated.
    """
    height, width = disparityMap.shape[:2]
    Q = buildQ(b=1, fx=width, fy=width, cx1=width / 2, cx2=width / 2, a1=0, a2=0, cy=height / 2)
    return reprojectImageTo3D(disparityMap, Q)


def get3DPoints(disparityMap, K1, K2, baseline):
    """
    ``RectifiedStereoRig.get3DPoints`` (reference: simplestereo/_rigs.py:569-628) for a rig described by its two
    final camera matrices ``K1``, ``K2`` (3x3, after rectification and fitting) and its baseline.
    """
    K1, K2 = np.asarray(K1, dtype=np.float64), np.asarray(K2, dtype=np.float64)
    Q = buildQ(b=baseline, fx=K1[0, 0], fy=K2[1, 1], cx1=K1[0, 2], cx2=K2[0, 2], a1=K1[0, 1], a2=K2[0, 1], cy=K1[1, 2])
    return reprojectImageTo3D(disparityMap, Q)


def computePoints(matcher, img1, img2, Q, return_disparity=False):
    """StereoASW.compute followed by the reprojection in ONE library call: the disparity map stays in HBM between
    the winner-take-all tail and the 3-D transform."""
    from .passive import StereoASW, _check_images
    if not isinstance(matcher, StereoASW):
        raise TypeError("computePoints needs a StereoASW matcher")
    img1, img2 = _check_images(img1, img2)
    Q = np.ascontiguousarray(np.asarray(Q, dtype=np.float64))
    if Q.shape != (4, 4):
        raise ValueError("Q must be a 4x4 matrix")
    h, w, _ = img1.shape
    pts = np.empty((h, w, 3), np.float32)
    disp = np.empty((h, w), np.int16) if return_disparity else None
    _cabi.check(_cabi.lib().ss_asw_compute_points(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *matcher._args(), _cabi.ptr(Q),
                                                  _cabi.ptr(disp), _cabi.ptr(pts)))
    return (pts, disp) if return_disparity else pts


def exportPLY(points3D, filepath, referenceImage=None, precision=6):
    """
    Export raw point cloud to PLY file (ASCII) -- reference: simplestereo/points.py:10-80, byte-identical output.

    The reference formats every point in a Python loop; here the text is produced by the native writer
    ``ss_export_ply`` on all host threads.

    Parameters
    ----------
    points3D : numpy.ndarray
        Array of 3D points. The last dimension must contain ordered x,y,z coordinates.
    filepath : str
        File path for the PLY file (absolute or relative).
    referenceImage : numpy.ndarray, optional
        Reference image to extract color from: same number of points as `points3D`, last dimension either
        1 (grayscale) or 3 (BGR, uint8).
    precision : int
        Decimal places to save coordinates with. Default to 6.
    """
    import ctypes
    import os
    points3D = np.asarray(points3D)
    shape = np.asarray(points3D.shape, dtype=np.int64)
    pts = points3D.reshape(-1, 3)
    if pts.dtype == np.float32:
        is_double = 0
    else:
        pts = pts.astype(np.float64, copy=False)
        is_double = 1
    pts = np.ascontiguousarray(pts)
    n = pts.shape[0]
    bgr = inten = None
    kind = 0
    if referenceImage is not None:
        referenceImage = np.asarray(referenceImage)
        if referenceImage.size == points3D.size:
            if referenceImage.dtype == np.uint8:
                bgr = np.ascontiguousarray(referenceImage.reshape(-1, 3))
            elif np.issubdtype(referenceImage.dtype, np.integer):     # "{:d}" accepts any integer image (points.py:53-55)
                inten, kind = np.ascontiguousarray(referenceImage.reshape(-1, 3), dtype=np.int64), 3
            else:                                                     # "{:d}" of a float raises ValueError upstream
                raise ValueError(f"Unknown format code 'd' for object of type '{referenceImage.dtype}'")
        else:
            inten = np.ravel(referenceImage)
            if inten.size != n:
                raise ValueError("Wrong image dimensions!")
            if np.issubdtype(inten.dtype, np.int64):                 # points.py:62
                inten, kind = np.ascontiguousarray(inten, dtype=np.int64), 1
            else:
                inten, kind = np.ascontiguousarray(inten, dtype=np.float64), 2
    _cabi.check(_cabi.lib().ss_export_ply(_cabi.ptr(pts), is_double, n, _cabi.ptr(shape), int(shape.size), _cabi.ptr(bgr),
                                          _cabi.ptr(inten), kind, ctypes.c_char_p(os.fsencode(filepath)), int(precision)))


def importPLY(filename, *properties):
    """
    Read the vertex table of an ASCII PLY file written by ``exportPLY`` (reference: simplestereo/points.py:82-121).

    ``properties`` selects columns of the vertex lines by position (default 0, 1, 2 = x, y, z); the result is a float
    array with one row per vertex and one column per requested property, in the requested order.
    """
    cols = properties if properties else (0, 1, 2)
    with open(filename, "r") as f:                # the reference's own scan: whole stripped lines, any capitalisation
        for line in f:
            if line.rstrip().lower() == "end_header":
                break
        rows = [ln.split(" ") for ln in f]        # no end_header: nothing is left to read, as upstream
    if not rows:
        return np.asarray([], dtype=float)
    out = np.empty((len(rows), len(cols)), dtype=float)
    for i, fields in enumerate(rows):
        for j, c in enumerate(cols):
            out[i, j] = float(fields[c])
    return out
