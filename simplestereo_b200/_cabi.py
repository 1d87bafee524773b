"""
ctypes binding of libsspassive.so (include/ss_passive.h).

This is the only bridge between the Python API and the CUDA kernels.  There is NO fallback: if the
shared library is missing or no B200 is usable, importing/calling raises -- by design (north_star:
"no Triton, no multi-backend dispatch, no CPU fallback").
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsspassive.so")

SS_OK = 0
SS_ERR_FORMAT, SS_ERR_TYPE, SS_ERR_DIMS, SS_ERR_WINSIZE, SS_ERR_PARAM, SS_ERR_CUDA, SS_ERR_NOMEM = -1, -2, -3, -4, -5, -6, -7

# every symbol include/ss_passive.h declares (tests check the library exports each one)
SYMBOLS = (
    "ss_init", "ss_init_devices", "ss_device_count", "ss_shutdown", "ss_last_error", "ss_abi_version",
    "ss_asw_compute", "ss_gsw_compute", "ss_asw_compute_rows", "ss_gsw_compute_rows",
    "ss_asw_compute_device", "ss_gsw_compute_device", "ss_asw_compute_multi_device", "ss_gsw_compute_multi_device",
    "ss_asw_partial_device", "ss_merge_keys_device", "ss_finalize_keys_device",
    "ss_asw_stages", "ss_gsw_stages", "ss_asw_stages_rows", "ss_gsw_stages_rows", "ss_debug_lab", "ss_debug_last_kernel",
    "ss_profile_enable", "ss_profile_read", "ss_profile_reset", "ss_measure_fp32_peak",
    # include/ss_post.h
    "ss_reproject", "ss_reproject_device", "ss_asw_compute_points",
    "ss_normalize_colormap", "ss_normalize_colormap_device", "ss_remap_linear", "ss_remap_linear_device",
    "ss_export_ply",
)

_lib = None


def build():
    """Compile libsspassive.so in-tree (nvcc, sm_100a)."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(simplestereo_b200 has no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_dbl, c_flt, c_vp, c_ll = ctypes.c_int, ctypes.c_double, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong
    L.ss_init.argtypes = [c_int]
    L.ss_init_devices.argtypes = [ctypes.POINTER(c_int), c_int]
    L.ss_device_count.argtypes = []
    L.ss_shutdown.argtypes = []
    L.ss_last_error.argtypes = []
    L.ss_last_error.restype = ctypes.c_char_p
    L.ss_abi_version.argtypes = []
    asw_tail = [c_int, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_int]            # W H win maxD minD gC gP consistent
    gsw_tail = [c_int, c_int, c_int, c_int, c_int, c_int, c_flt, c_int, c_int]     # W H win maxD minD gamma fMax iters bins
    L.ss_asw_compute.argtypes = [c_vp, c_vp] + asw_tail + [c_vp]
    L.ss_gsw_compute.argtypes = [c_vp, c_vp] + gsw_tail + [c_vp]
    L.ss_asw_compute_rows.argtypes = [c_vp, c_vp] + asw_tail + [c_int, c_int, c_vp]
    L.ss_gsw_compute_rows.argtypes = [c_vp, c_vp] + gsw_tail + [c_int, c_int, c_vp]
    L.ss_asw_compute_device.argtypes = [c_vp, c_vp] + asw_tail + [c_int, c_int, c_vp, c_vp]
    L.ss_gsw_compute_device.argtypes = [c_vp, c_vp] + gsw_tail + [c_int, c_int, c_vp, c_vp]
    L.ss_asw_compute_multi_device.argtypes = [c_vp, c_vp] + asw_tail + [c_vp, c_vp]
    L.ss_gsw_compute_multi_device.argtypes = [c_vp, c_vp] + gsw_tail + [c_vp, c_vp]
    L.ss_debug_lab.argtypes = [c_vp, c_int, c_int, c_vp]
    L.ss_debug_last_kernel.argtypes = [c_int, ctypes.POINTER(c_int)]
    L.ss_asw_partial_device.argtypes = [c_vp, c_vp] + asw_tail + [c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]
    L.ss_merge_keys_device.argtypes = [c_vp, c_int, c_ll, c_vp]
    L.ss_finalize_keys_device.argtypes = [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp]
    L.ss_asw_stages.argtypes = [c_vp, c_vp] + asw_tail + [c_vp, c_vp, c_vp, c_vp, c_vp]
    L.ss_gsw_stages.argtypes = [c_vp, c_vp] + gsw_tail + [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    L.ss_asw_stages_rows.argtypes = [c_vp, c_vp] + asw_tail + [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]
    L.ss_gsw_stages_rows.argtypes = [c_vp, c_vp] + gsw_tail + [c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]
    L.ss_profile_enable.argtypes = [c_int]
    L.ss_profile_read.argtypes = [ctypes.POINTER(c_dbl), ctypes.POINTER(c_ll), ctypes.POINTER(c_ll)]
    L.ss_profile_reset.argtypes = []
    L.ss_measure_fp32_peak.argtypes = [ctypes.POINTER(c_dbl), c_vp]
    # include/ss_post.h
    L.ss_reproject.argtypes = [c_vp, c_int, c_int, c_vp, c_vp]
    L.ss_reproject_device.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_vp]
    L.ss_asw_compute_points.argtypes = [c_vp, c_vp] + asw_tail + [c_vp, c_vp, c_vp]
    L.ss_normalize_colormap.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_vp]
    L.ss_normalize_colormap_device.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp]
    L.ss_remap_linear.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp]
    L.ss_remap_linear_device.argtypes = [c_vp, c_int, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp]
    L.ss_export_ply.argtypes = [c_vp, c_int, c_ll, c_vp, c_int, c_vp, c_vp, c_int, ctypes.c_char_p, c_int]
    for s in SYMBOLS:
        if s != "ss_last_error":
            getattr(L, s).restype = c_int
    _lib = L
    return L


_MESSAGES = {
    SS_ERR_FORMAT: (ValueError, "Invalid input format!"),                       # _passive.cpp:304, :712
    SS_ERR_TYPE: (TypeError, "Wrong type input!"),                              # :312, :720
    SS_ERR_DIMS: (ValueError, "Wrong image dimensions!"),                       # :319, :727
    SS_ERR_WINSIZE: (ValueError, "winSize must be a positive odd number!"),     # :323, :731
}


def check(rc):
    """Map a C return code to the exception the reference extension would raise."""
    if rc == SS_OK:
        return
    msg = lib().ss_last_error().decode("utf-8", "replace")
    if rc in _MESSAGES:
        exc, text = _MESSAGES[rc]
        raise exc(msg if rc == SS_ERR_DIMS and msg != text else text)
    if rc == SS_ERR_PARAM:
        raise ValueError(msg)
    if rc == SS_ERR_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(f"libsspassive: {msg} (code {rc})")


def init_devices(devices=None):
    """ss_init_devices: the devices the host entry points shard image rows over (None: every visible device)."""
    if devices is None:
        check(lib().ss_init_devices(None, 0))
    else:
        devs = [int(d) for d in devices]
        arr = (ctypes.c_int * len(devs))(*devs)
        check(lib().ss_init_devices(arr, len(devs)))
    return lib().ss_device_count()


_applied_devices = None


def use_devices(devices):
    """Apply a matcher's ``devices=`` choice (None: leave the process-wide setting alone)."""
    global _applied_devices
    if devices is None:
        return
    key = "all" if isinstance(devices, str) else tuple(int(d) for d in devices)
    if isinstance(devices, str) and devices != "all":
        raise ValueError("devices must be None, 'all' or a sequence of CUDA device indices")
    if key != _applied_devices:
        init_devices(None if key == "all" else key)
        _applied_devices = key


def ptr(a):
    """Host pointer of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def last_kernel(device=-1):
    """("tc" | "ws" | None, disparity chunk) of the last aggregation launch on ``device``."""
    dc = ctypes.c_int()
    k = lib().ss_debug_last_kernel(device, ctypes.byref(dc))
    return {1: "tc", 2: "ws"}.get(k), dc.value


def last_kernel_name(device=-1):
    """"k_aggregate_tc" | "k_aggregate_ws" | None of the last aggregation launch on ``device``."""
    k = lib().ss_debug_last_kernel(device, None)
    return {1: "k_aggregate_tc", 2: "k_aggregate_ws"}.get(k)


def lab(img):
    """BGR uint8 [H,W,3] -> float32 CIELab [H,W,3] as the kernels compute it (ss_debug_lab)."""
    img = np.ascontiguousarray(img)
    out = np.empty(img.shape, np.float32)
    check(lib().ss_debug_lab(ptr(img), img.shape[1], img.shape[0], ptr(out)))
    return out


def measure_fp32_peak(stream=None):
    tf = ctypes.c_double()
    check(lib().ss_measure_fp32_peak(ctypes.byref(tf), stream))
    return tf.value


def profile_read():
    ms, n, tot = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_longlong()
    check(lib().ss_profile_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(tot)))
    return ms.value, n.value, tot.value
