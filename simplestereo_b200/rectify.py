"""
rectify
=======
B200-native counterpart of ``RectifiedStereoRig.rectifyImages`` (reference: simplestereo/_rigs.py:543-567), the step
that feeds the stereo matcher: ``cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR)`` for both images of the pair, with the
float32 maps produced by ``cv2.initUndistortRectifyMap(..., cv2.CV_32FC1)`` (_rigs.py:540-541).  The CUDA kernel
reproduces OpenCV's fixed-point bilinear arithmetic bit for bit (oracle/post_oracle.py).  No CPU fallback.

Computing the maps themselves (calibration, rectifying homographies, fitting) stays out of scope.
"""
import numpy as np

from . import _cabi


def remap(img, mapx, mapy):
    """``cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR)`` for a uint8 BGR image and float32 maps (border: constant 0)."""
    if not isinstance(img, np.ndarray) or img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("Wrong image dimensions!")
    if img.dtype != np.uint8:
        raise TypeError("Wrong type input!")
    mapx = np.ascontiguousarray(mapx, dtype=np.float32)
    mapy = np.ascontiguousarray(mapy, dtype=np.float32)
    if mapx.ndim != 2 or mapx.shape != mapy.shape:
        raise ValueError("Wrong image dimensions!")
    img = np.ascontiguousarray(img)
    sh, sw, _ = img.shape
    dh, dw = mapx.shape
    out = np.empty((dh, dw, 3), np.uint8)
    _cabi.check(_cabi.lib().ss_remap_linear(_cabi.ptr(img), sw, sh, _cabi.ptr(mapx), _cabi.ptr(mapy), dw, dh, _cabi.ptr(out)))
    return out


def rectifyImages(img1, img2, mapx1, mapy1, mapx2, mapy2):
    """
    Undistort, rectify and apply the fitting transformation to a couple of images coming from the stereo rig
    (reference: RectifiedStereoRig.rectifyImages, _rigs.py:543-567, with the rig's four maps passed explicitly).

    Returns img1_rect, img2_rect.
    """
    return remap(img1, mapx1, mapy1), remap(img2, mapx2, mapy2)
