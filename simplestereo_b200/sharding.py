"""
Multi-GPU sharding of one stereo pair: one process per GPU (torch.distributed, NCCL over NVLink), the
compute goes through the device-resident C-ABI entry points, and ONE all-gather reassembles the result.

Two partitions, both named by BASELINE.json / SURVEY.md 8(e):

* ``rows``       image-row stripes.  Rows are independent jobs in the reference (a row index is the unit its
                 thread pool pops, _passive.cpp:372-374, and the L-R check + fill are row-local, :251-285), so
                 a stripe needs no halo exchange: every rank holds the (tiny) input pair and computes rows
                 [r0, r1).  Collective: all_gather of int16 stripes.
* ``disparity``  disparity-range shards (config C5).  Every rank evaluates a contiguous sub-range for all
                 pixels and returns packed (cost, disparity) winners; collective: all_gather of the uint64 key
                 planes, then an element-wise unsigned min (which is also the smallest-disparity tie-break),
                 then the row-local invalidate + fill on every rank.

torch is plumbing only (device buffers, streams, process group); tensors cross into the library as raw pointers.
"""
import numpy as np

from . import _cabi


def row_stripes(height, n):
    """Contiguous stripes of ceil(height/n) rows; trailing stripes may be short or empty."""
    s = -(-height // n) if n > 0 else height
    return [(min(k * s, height), min((k + 1) * s, height)) for k in range(n)]


def disparity_shards(min_disp, max_disp, n):
    """Inclusive sub-ranges [(d0, d1), ...] covering [min_disp, max_disp]; empty shards have d1 < d0."""
    D = max(max_disp - min_disp + 1, 0)
    s = -(-D // n) if n > 0 and D > 0 else 0
    out = []
    for k in range(n):
        d0 = min_disp + k * s
        d1 = min(d0 + s - 1, max_disp)
        out.append((d0, d1) if s > 0 and d0 <= max_disp else (max_disp + 1, max_disp))
    return out


def gather_rows(stripe, height, group=None):
    """all_gather equal-sized row stripes ([S, W], S = ceil(H/n)) into the [H, W] map (on every rank)."""
    import torch
    import torch.distributed as dist
    n = dist.get_world_size(group)
    s, w = stripe.shape
    full = torch.empty((n * s, w), dtype=stripe.dtype, device=stripe.device)
    # int16 is not a collective dtype (neither NCCL nor gloo): gather the raw bytes
    dist.all_gather_into_tensor(full.view(torch.uint8), stripe.contiguous().view(torch.uint8), group=group)
    return full[:height]


class ShardedStereoASW:
    """StereoASW over the ranks of a torch.distributed process group.

    ``matcher`` is a ``simplestereo_b200.passive.StereoASW``; ``mode`` is "rows" or "disparity".
    ``compute_device`` takes/returns torch CUDA tensors (uint8 [H,W,3] in, int16 [H,W] out, identical on
    every rank); ``compute`` is the numpy-in / numpy-out convenience wrapper.
    """

    def __init__(self, matcher, group=None, mode="rows"):
        import torch.distributed as dist
        if mode not in ("rows", "disparity"):
            raise ValueError("mode must be 'rows' or 'disparity'")
        self.m = matcher
        self.group = group
        self.mode = mode
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._bufs = {}

    def _buf(self, key, shape, dtype, device):
        import torch
        b = self._bufs.get(key)
        if b is None or tuple(b.shape) != tuple(shape) or b.dtype != dtype or b.device != device:
            b = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = b
        return b

    def compute_device(self, d_img1, d_img2):
        import torch
        import torch.distributed as dist
        h, w, _ = d_img1.shape
        win, maxd, mind, gc, gp, cons = self.m._args()
        L = _cabi.lib()
        st = torch.cuda.current_stream().cuda_stream
        dev = d_img1.device
        if self.mode == "rows":
            s = -(-h // self.world)
            r0, r1 = row_stripes(h, self.world)[self.rank]
            stripe = self._buf("stripe", (s, w), torch.int16, dev)
            _cabi.check(L.ss_asw_compute_device(d_img1.data_ptr(), d_img2.data_ptr(), w, h, win, maxd, mind, gc, gp, cons,
                                                r0, r1, stripe.data_ptr(), st))
            if self.world == 1:
                return stripe[:h]
            full = self._buf("full", (self.world * s, w), torch.int16, dev)
            # int16 is not an NCCL dtype: gather the raw bytes
            dist.all_gather_into_tensor(full.view(torch.uint8), stripe.view(torch.uint8), group=self.group)
            return full[:h]
        # disparity-range shards: keys planes [left | right] per rank
        d0, d1 = disparity_shards(mind, maxd, self.world)[self.rank]
        npx = h * w
        planes = 2 if cons else 1
        mine = self._buf("keys", (planes, npx), torch.int64, dev)
        _cabi.check(L.ss_asw_partial_device(d_img1.data_ptr(), d_img2.data_ptr(), w, h, win, maxd, mind, gc, gp, cons,
                                            0, h, d0, d1, mine[0].data_ptr(), mine[1].data_ptr() if cons else None, st))
        if self.world > 1:
            # rank-major concatenation along dim 0 (the layout every backend's all_gather_into_tensor accepts)
            allk = self._buf("allkeys", (self.world * planes, npx), torch.int64, dev)
            dist.all_gather_into_tensor(allk, mine, group=self.group)
            _cabi.check(L.ss_merge_keys_device(allk.data_ptr(), self.world, planes * npx, st))
            merged = allk[:planes]
        else:
            merged = mine
        out = self._buf("out", (h, w), torch.int16, dev)
        _cabi.check(L.ss_finalize_keys_device(merged[0].data_ptr(), merged[1].data_ptr() if cons else None, w, h, mind,
                                              out.data_ptr(), st))
        return out

    def compute(self, img1, img2):
        import torch
        from .passive import _check_images
        img1, img2 = _check_images(img1, img2)
        d1 = torch.from_numpy(img1).cuda(non_blocking=True)
        d2 = torch.from_numpy(img2).cuda(non_blocking=True)
        return self.compute_device(d1, d2).cpu().numpy()
