"""
Multi-GPU sharding of one stereo pair: one process per GPU (torch.distributed, NCCL over NVLink), the
compute goes through the device-resident C-ABI entry points, and ONE all-gather reassembles the result.

Two partitions, both named by BASELINE.json / SURVEY.md 8(e):

* ``rows``       image-row stripes.  Rows are independent jobs in the reference (a row index is the unit its
                 thread pool pops, _passive.cpp:372-374, and the L-R check + fill are row-local, :251-285), so
                 a stripe needs no halo exchange: every rank holds the (tiny) input pair and computes rows
                 [r0, r1).  Collective: all_gather of int16 stripes.
* ``disparity``  disparity-range shards (config C5).  Every rank evaluates a contiguous sub-range and returns packed
                 (cost, disparity) winners; collective: all_gather of the uint64 key planes, then an element-wise
                 unsigned min (which is also the smallest-disparity tie-break), then the row-local invalidate + fill on
                 every rank.  The library evaluates disparities in chunks anchored at minDisparity (128 wide when the
                 call spans more than 64 candidates, ``chunk_size``) and a shard that is a whole number of chunks wastes
                 no work, so the range is cut at chunk boundaries; when there are more ranks than chunks (C5: 512
                 disparities = 4 chunks on 8 GPUs) each disparity group is split further into image-row stripes
                 (``disparity_row_grid``).  Either way every cost is computed by the same kernel on the same chunk grid as
                 the unsharded call: the merged map is bit-identical to it.

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


def chunk_size(min_disp, max_disp):
    """Disparity chunk of the aggregation kernels for a call over [min_disp, max_disp] (csrc/ss_passive.cu, make_plan)."""
    D = max_disp - min_disp + 1
    return 32 if D <= 32 else (64 if D <= 64 else 128)


def disparity_row_grid(min_disp, max_disp, height, n):
    """Partition for ``mode="disparity"`` over n ranks: n_d chunk-aligned disparity groups x n_r row stripes, n_d * n_r = n,
    n_d the largest divisor of n not above the number of chunks.  Returns (n_d, n_r, [(d0, d1, r0, r1) per rank]) with rank
    = kd * n_r + kr; d1 < d0 marks an empty disparity group, r0 == r1 an empty stripe."""
    D = max(max_disp - min_disp + 1, 0)
    dc = chunk_size(min_disp, max_disp)
    nch = max(-(-D // dc), 1)
    n_d = max(k for k in range(1, n + 1) if n % k == 0 and k <= nch)
    n_r = n // n_d
    cpg = -(-nch // n_d)                      # chunks per disparity group
    S = -(-height // n_r)
    parts = []
    for rank in range(n):
        kd, kr = divmod(rank, n_r)
        d0 = min_disp + kd * cpg * dc
        d1 = min(max_disp, d0 + cpg * dc - 1)
        if D == 0 or d0 > max_disp:
            d0, d1 = max_disp + 1, max_disp
        parts.append((d0, d1, min(kr * S, height), min((kr + 1) * S, height)))
    return n_d, n_r, parts


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
        # disparity-range shards (x row stripes when ranks outnumber chunks): key planes [left | right] per rank
        n_d, n_r, parts = disparity_row_grid(mind, maxd, h, self.world)
        d0, d1, r0, r1 = parts[self.rank]
        S = -(-h // n_r)
        planes = 2 if cons else 1
        key = ("keys", planes, S * w)
        fresh = key not in self._bufs
        mine = self._buf(key, (planes, S * w), torch.int64, dev)
        if fresh:
            mine.fill_(-1)                      # KEY_NONE: rows past the image in the last stripe are never written
        _cabi.check(L.ss_asw_partial_device(d_img1.data_ptr(), d_img2.data_ptr(), w, h, win, maxd, mind, gc, gp, cons,
                                            r0, r1, d0, d1, mine[0].data_ptr(), mine[1].data_ptr() if cons else None, st))
        if self.world > 1:
            # rank-major concatenation along dim 0 (the layout every backend's all_gather_into_tensor accepts):
            # [kd][kr][plane][S*W] -> element-wise min over kd leaves the merged planes of every stripe in the first n_r slots
            allk = self._buf("allkeys", (self.world * planes, S * w), torch.int64, dev)
            dist.all_gather_into_tensor(allk, mine, group=self.group)
            _cabi.check(L.ss_merge_keys_device(allk.data_ptr(), n_d, n_r * planes * S * w, st))
            merged = allk[:n_r * planes].view(n_r, planes, S * w)
        else:
            merged = mine.view(1, planes, S * w)
        out = self._buf("out", (n_r * S, w), torch.int16, dev)
        for kr in range(n_r):
            rows = min((kr + 1) * S, h) - min(kr * S, h)
            if rows <= 0:
                continue
            _cabi.check(L.ss_finalize_keys_device(merged[kr, 0].data_ptr(), merged[kr, 1].data_ptr() if cons else None, w, rows,
                                                  mind, out[kr * S:].data_ptr(), st))
        return out[:h]

    def compute(self, img1, img2):
        import torch
        from .passive import _check_images
        img1, img2 = _check_images(img1, img2)
        d1 = torch.from_numpy(img1).cuda(non_blocking=True)
        d2 = torch.from_numpy(img2).cuda(non_blocking=True)
        return self.compute_device(d1, d2).cpu().numpy()
