"""
Seeded synthetic rectified stereo pairs -- the measurement input defined in SURVEY.md section 8(d)/9.2.

Textured (blurred noise, min-max stretched) right image, piecewise-constant ground-truth disparity
(8 vertical bands x 2 halves), left = right warped by the disparity plus integer noise in [-2, 2].
i.i.d. noise or constant pairs are deliberately NOT used for throughput: they saturate the truncated
absolute difference everywhere and turn every argmin into a rounding tie.
"""
import numpy as np


def synth_pair(width, height, max_disp, seed=0):
    """Return (left_bgr_u8, right_bgr_u8, gt_disparity_int32), all C-contiguous."""
    import cv2
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (height, width + max_disp + 1, 3), dtype=np.uint8)
    base = cv2.GaussianBlur(base, (0, 0), 1.5)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    disp = np.zeros((height, width), np.int32)
    for b in range(8):
        x0, x1 = b * width // 8, (b + 1) * width // 8
        disp[:height // 2, x0:x1] = int(rng.integers(0, max_disp + 1))
        disp[height // 2:, x0:x1] = int(rng.integers(0, max_disp + 1))
    right = base[:, :width].copy()
    X = np.arange(width)[None, :].repeat(height, 0)
    Y = np.arange(height)[:, None].repeat(width, 1)
    left = right[Y, np.clip(X - disp, 0, width - 1)]
    left = np.clip(left.astype(int) + rng.integers(-2, 3, left.shape), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left), np.ascontiguousarray(right), disp
