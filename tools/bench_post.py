"""Device-timed throughput of the pre/post kernels (include/ss_post.h) against the HBM roofline.

    python tools/bench_post.py            # 4K frames (BASELINE.json config C5's size), CUDA events, L2 flushed

Algorithmic bytes per pixel: reproject 2 (int16 in) + 12 (3 x float32 out); normalise + colour map 2 x 2 (min/max pass
+ map pass) + 3 out; remap 8 (two float32 maps) + 3 in (each source byte once) + 3 out.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import simplestereo_b200 as ss  # noqa: E402
from simplestereo_b200 import _cabi  # noqa: E402

W, H = 3840, 2160
L = _cabi.lib()
_cabi.check(L.ss_init(0))
st = torch.cuda.current_stream()
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
rng = np.random.default_rng(0)
disp = torch.from_numpy(rng.integers(0, 512, (H, W)).astype(np.int16)).cuda()
pts = torch.empty((H, W, 3), dtype=torch.float32, device="cuda")
Q = np.ascontiguousarray(ss.points.buildQ(b=0.54, fx=2100.0, fy=2100.0, cx1=1920.0, cx2=1925.0, a1=0.1, a2=0.2, cy=1080.0))
lut = torch.from_numpy(ss.display.COLORMAP_JET.copy()).cuda()
bgr = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
src = torch.from_numpy(rng.integers(0, 256, (H, W, 3), dtype=np.uint8)).cuda()
yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
mx = torch.from_numpy((xx * 0.998 + 3.3 + 2e-6 * (yy - H / 2) ** 2).astype(np.float32)).cuda()
my = torch.from_numpy((yy * 1.001 - 1.7 + 1e-6 * (xx - W / 2) ** 2).astype(np.float32)).cuda()
dst = torch.empty((H, W, 3), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

KERNELS = {
    "reproject (k_reproject)": (lambda: L.ss_reproject_device(disp.data_ptr(), W, H, Q.ctypes.data, pts.data_ptr(), st.cuda_stream), 14),
    "normalize+colormap (k_minmax_i16 + k_normalize_colormap)": (lambda: L.ss_normalize_colormap_device(disp.data_ptr(), W, H, lut.data_ptr(), None, bgr.data_ptr(), st.cuda_stream), 7),
    "remap (k_remap_linear)": (lambda: L.ss_remap_linear_device(src.data_ptr(), W, H, mx.data_ptr(), my.data_ptr(), W, H, dst.data_ptr(), st.cuda_stream), 14),
}
for name, (fn, bpp) in KERNELS.items():
    for _ in range(3):
        _cabi.check(fn())
    ms = []
    for _ in range(10):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        _cabi.check(fn())
        e1.record(st)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms)) * 1e-3
    gbs = bpp * W * H / t / 1e9
    print(f"{name}: {t * 1e6:.1f} us per {W}x{H} frame, {W * H / t / 1e9:.2f} Gpix/s, {gbs:.0f} GB/s algorithmic = {gbs / hbm:.2f} of the measured HBM peak ({hbm:.0f} GB/s)")
