"""
passive
=======
B200-native drop-in for ``simplestereo.passive`` (reference: simplestereo/passive.py:16-158).

Same classes, constructor arguments, ``compute(img1, img2)`` signature and numpy-in / numpy-out
contract; the body forwards to the CUDA kernels in ``libsspassive.so`` through the C ABI of
``include/ss_passive.h`` instead of the CPython extension ``simplestereo._passive``.

Differences from the reference, all strict supersets of its behaviour:
  * inputs are made C-contiguous (the reference silently mis-reads Fortran-ordered arrays);
  * ``img2`` dtype is checked too (the reference tests ``img1`` twice, _passive.cpp:309);
  * ``gammaC``/``gammaP``/``gamma`` <= 0 and ``minDisparity`` < 0 raise ``ValueError`` instead of
    producing NaN costs / out-of-row reads;
  * three optional extras: ``rows=(r0, r1)`` computes an image-row stripe; ``devices=[0, 1, ...]`` (constructor keyword,
    ``"all"`` for every visible GPU) shards the image rows of ``compute`` over several GPUs inside this process --
    ``ss_init_devices``: one host thread and one context per GPU, no torch, no collective, bit-identical to one GPU; and
    ``StereoASW.compute_staged`` / ``StereoGSW.compute_staged`` expose the intermediate maps and cost volumes used by
    the parity tests.
"""
import numpy as np

from . import _cabi


def _check_images(img1, img2):
    # "O!O!..." parse failure -> ValueError("Invalid input format!")  (_passive.cpp:301-306)
    if not isinstance(img1, np.ndarray) or not isinstance(img2, np.ndarray):
        raise ValueError("Invalid input format!")
    if img1.dtype != np.uint8 or img2.dtype != np.uint8:
        raise TypeError("Wrong type input!")                        # _passive.cpp:309-314
    if (img1.ndim != 3 or img2.ndim != 3 or img1.shape[2] != 3 or img2.shape[2] != 3
            or img1.shape[0] != img2.shape[0] or img1.shape[1] != img2.shape[1]):
        raise ValueError("Wrong image dimensions!")                 # _passive.cpp:315-321
    return np.ascontiguousarray(img1), np.ascontiguousarray(img2)


def _as_int(v):
    # the "i" format of PyArg_ParseTuple rejects floats (e.g. GSW gamma, _passive.cpp:709)
    if isinstance(v, (bool, np.bool_)):
        return int(v)
    if isinstance(v, (int, np.integer)):
        return int(v)
    raise ValueError("Invalid input format!")


def _as_float(v):
    if isinstance(v, (int, float, np.integer, np.floating)):
        return float(v)
    raise ValueError("Invalid input format!")


class StereoASW():
    """
    Adaptive Support-Weight stereo matching (K. Yoon, I. Kweon, 2006) on B200.

    Parameters (identical to simplestereo.passive.StereoASW, passive.py:59)
    ----------
    winSize : int
        Side of the square window. Must be an odd positive number. Default is 35.
    maxDisparity : int
        Maximum accepted disparity (inclusive). Default is 16.
    minDisparity : int
        Minimum valid disparity (inclusive), usually zero. Default is 0.
    gammaC : float
        Color parameter. Default is 5.
    gammaP : float
        Proximity parameter. Default is 17.5.
    consistent : bool
        If True the right-reference pass, left-right check and occlusion filling of the reference
        (_passive.cpp:191-285) are applied.  On the GPU the right-reference WTA is fused into the
        aggregation launch (one shared-memory pass over the block's costs), not a second aggregation.
    devices : None, "all" or a sequence of CUDA device indices (keyword only, not in the reference)
        Shard the image rows of ``compute`` over these GPUs inside this process (``ss_init_devices``).
    """
    def __init__(self, winSize=35, maxDisparity=16, minDisparity=0, gammaC=5, gammaP=17.5, consistent=False, *, devices=None):
        if not (winSize > 0 and winSize % 2 == 1):
            raise ValueError("winSize must be a positive odd number!")
        self.devices = devices
        self.winSize = winSize
        self.maxDisparity = maxDisparity
        self.minDisparity = minDisparity
        self.gammaC = gammaC
        self.gammaP = gammaP
        self.consistent = consistent

    def _args(self):
        return (_as_int(self.winSize), _as_int(self.maxDisparity), _as_int(self.minDisparity),
                _as_float(self.gammaC), _as_float(self.gammaP), int(bool(self.consistent)))

    def compute(self, img1, img2, rows=None):
        """
        Compute the disparity map for a rectified BGR pair (left, right).

        Returns numpy.ndarray (np.int16) of the image height and width -- or of the requested
        ``rows=(r0, r1)`` stripe.
        """
        img1, img2 = _check_images(img1, img2)
        h, w, _ = img1.shape
        L = _cabi.lib()
        _cabi.use_devices(self.devices)
        if rows is None:
            out = np.empty((h, w), np.int16)
            _cabi.check(L.ss_asw_compute(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *self._args(), _cabi.ptr(out)))
        else:
            r0, r1 = int(rows[0]), int(rows[1])
            out = np.empty((max(r1 - r0, 0), w), np.int16)
            _cabi.check(L.ss_asw_compute_rows(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *self._args(), r0, r1, _cabi.ptr(out)))
        return out

    def compute_staged(self, img1, img2, cost=False, rows=None):
        """dict(left, right, invalid, final[, cost]) -- the staged outputs of SURVEY.md 8(c), for the whole frame or for the
        image-row stripe ``rows=(r0, r1)`` (every array then holds only the stripe's rows)."""
        img1, img2 = _check_images(img1, img2)
        h, w, _ = img1.shape
        win, maxd, mind, gc, gp, cons = self._args()
        r0, r1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
        n = max(r1 - r0, 0)
        D = max(maxd - mind + 1, 0)
        o = {"left": np.zeros((n, w), np.int16), "right": np.zeros((n, w), np.int16),
             "invalid": np.zeros((n, w), np.uint8), "final": np.zeros((n, w), np.int16)}
        vol = np.full((n, w, D), np.inf, np.float32) if cost else None
        _cabi.check(_cabi.lib().ss_asw_stages_rows(_cabi.ptr(img1), _cabi.ptr(img2), w, h, win, maxd, mind, gc, gp, cons, r0, r1,
                                                   _cabi.ptr(o["left"]), _cabi.ptr(o["right"]), _cabi.ptr(o["invalid"]),
                                                   _cabi.ptr(o["final"]), _cabi.ptr(vol) if (cost and D > 0) else None))
        if cost:
            o["cost"] = vol
        return o


class StereoGSW():
    """
    Geodesic Support-Weight stereo matching (Hosni et al., 2009) as implemented -- incompletely -- by
    the reference (passive.py:99-158, _passive.cpp:408-700), on B200.

    Parameters (identical to simplestereo.passive.StereoGSW, passive.py:133-134)
    ----------
    winSize : int, optional (default 11)
    maxDisparity, minDisparity : int, optional (defaults 16, 0; inclusive range)
    gamma : int, optional (default 10; must be an int, as upstream's "i" parse format)
    fMax : int or float, optional (default 120)
    iterations : int, optional (default 3).  The reference's relaxation converges in its first
        forward pass, so any value >= 1 gives the same map; <= 0 keeps only the
        centre weight.
    bins : int, optional (unused, as upstream)
    """
    def __init__(self, winSize=11, maxDisparity=16, minDisparity=0, gamma=10,
                 fMax=120, iterations=3, bins=20, *, devices=None):
        if not (winSize > 0 and winSize % 2 == 1):
            raise ValueError("winSize must be a positive odd number!")
        self.devices = devices
        self.winSize = winSize
        self.gamma = gamma
        self.maxDisparity = maxDisparity
        self.minDisparity = minDisparity
        self.fMax = fMax
        self.iterations = iterations
        self.bins = bins

    def _args(self):
        return (_as_int(self.winSize), _as_int(self.maxDisparity), _as_int(self.minDisparity),
                _as_int(self.gamma), _as_float(self.fMax), _as_int(self.iterations), _as_int(self.bins))

    def compute(self, img1, img2, rows=None):
        """Compute the disparity map for 3-channel images (left, right)."""
        img1, img2 = _check_images(img1, img2)
        h, w, _ = img1.shape
        L = _cabi.lib()
        _cabi.use_devices(self.devices)
        if rows is None:
            out = np.empty((h, w), np.int16)
            _cabi.check(L.ss_gsw_compute(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *self._args(), _cabi.ptr(out)))
        else:
            r0, r1 = int(rows[0]), int(rows[1])
            out = np.empty((max(r1 - r0, 0), w), np.int16)
            _cabi.check(L.ss_gsw_compute_rows(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *self._args(), r0, r1, _cabi.ptr(out)))
        return out

    def compute_staged(self, img1, img2, cost=False, rows=None):
        img1, img2 = _check_images(img1, img2)
        h, w, _ = img1.shape
        args = self._args()
        r0, r1 = (0, h) if rows is None else (int(rows[0]), int(rows[1]))
        n = max(r1 - r0, 0)
        D = max(args[1] - args[2] + 1, 0)
        o = {"left": np.zeros((n, w), np.int16), "right": np.zeros((n, w), np.int16),
             "invalid": np.zeros((n, w), np.uint8), "final": np.zeros((n, w), np.int16)}
        vl = np.full((n, w, D), np.inf, np.float32) if cost else None
        vr = np.full((n, w, D), np.inf, np.float32) if cost else None
        on = cost and D > 0
        _cabi.check(_cabi.lib().ss_gsw_stages_rows(_cabi.ptr(img1), _cabi.ptr(img2), w, h, *args, r0, r1,
                                                   _cabi.ptr(o["left"]), _cabi.ptr(o["right"]), _cabi.ptr(o["invalid"]),
                                                   _cabi.ptr(o["final"]), _cabi.ptr(vl) if on else None, _cabi.ptr(vr) if on else None))
        if cost:
            o["cost_left"], o["cost_right"] = vl, vr
        return o
