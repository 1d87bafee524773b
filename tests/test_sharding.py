"""Host-side logic of the multi-GPU path, exercised on CPU with the gloo backend (world_size 2):
row-stripe / disparity-range partitioning and the all-gather reassembly.  The per-stripe compute is the
oracle here (there is no GPU in this tier); on the GPU box the same code path runs the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from simplestereo_b200.sharding import chunk_size, disparity_row_grid, disparity_shards, gather_rows, row_stripes


def test_row_stripes_cover_all_rows():
    for h in (1, 7, 47, 375, 2160):
        for n in (1, 2, 3, 4, 8):
            st = row_stripes(h, n)
            assert len(st) == n
            assert st[0][0] == 0 and st[-1][1] == h
            assert all(a[1] == b[0] for a, b in zip(st, st[1:]))
            s = -(-h // n)
            assert all(0 <= r1 - r0 <= s for r0, r1 in st)
            assert sum(r1 - r0 for r0, r1 in st) == h


def test_disparity_shards_cover_range():
    for (lo, hi) in ((0, 127), (4, 14), (0, 0), (5, 3), (0, 511)):
        for n in (1, 2, 4, 8):
            sh = disparity_shards(lo, hi, n)
            ds = [d for d0, d1 in sh for d in range(d0, d1 + 1)]
            assert ds == list(range(lo, hi + 1))


def test_disparity_row_grid_covers_every_pair_once_on_chunk_boundaries():
    for (lo, hi) in ((0, 127), (4, 14), (0, 0), (5, 3), (0, 511), (3, 150), (0, 255), (7, 300)):
        for h in (1, 21, 375):
            for n in (1, 2, 3, 4, 8):
                n_d, n_r, parts = disparity_row_grid(lo, hi, h, n)
                assert n_d * n_r == n and len(parts) == n
                dc = chunk_size(lo, hi)
                seen = np.zeros((h, max(hi - lo + 1, 0)), np.int32)
                for rank, (d0, d1, r0, r1) in enumerate(parts):
                    assert rank == (rank // n_r) * n_r + rank % n_r
                    if d1 >= d0:
                        assert (d0 - lo) % dc == 0                      # shards start on the library's chunk grid
                        assert d1 == hi or (d1 + 1 - lo) % dc == 0
                        seen[r0:r1, d0 - lo:d1 - lo + 1] += 1
                assert (seen == 1).all()
    # C5: 512 disparities on 8 GPUs -> 4 chunk groups x 2 row halves
    n_d, n_r, parts = disparity_row_grid(0, 511, 2160, 8)
    assert (n_d, n_r) == (4, 2) and parts[0] == (0, 127, 0, 1080) and parts[7] == (384, 511, 1080, 2160)


def test_disparity_row_grid_key_layout_merges_to_the_unsharded_winners():
    """The layout ShardedStereoASW(mode="disparity") relies on: all_gather concatenates the ranks' key planes as
    [kd][kr][plane][S*W]; an element-wise min over kd (ss_merge_keys_device with n = n_r*planes*S*W) leaves the merged
    planes of stripe kr at slot kr."""
    rng = np.random.default_rng(3)
    h, w, lo, hi, world, planes = 11, 17, 0, 299, 8, 2
    cost = rng.random((planes, h, w, hi - lo + 1)).astype(np.float32)
    keys_of = lambda c, d: (c.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.uint64(d)
    want = np.full((planes, h, w), np.iinfo(np.uint64).max, np.uint64)
    for d in range(lo, hi + 1):
        want = np.minimum(want, keys_of(cost[..., d - lo], d))
    n_d, n_r, parts = disparity_row_grid(lo, hi, h, world)
    S = -(-h // n_r)
    allk = np.full((world, planes, S * w), np.iinfo(np.uint64).max, np.uint64)
    for rank, (d0, d1, r0, r1) in enumerate(parts):
        for d in range(d0, d1 + 1):
            k = keys_of(cost[:, r0:r1, :, d - lo], d).reshape(planes, -1)
            allk[rank, :, :k.shape[1]] = np.minimum(allk[rank, :, :k.shape[1]], k)
    merged = allk.reshape(n_d, n_r * planes * S * w).min(axis=0).reshape(n_r, planes, S, w)
    got = np.concatenate([merged[kr] for kr in range(n_r)], axis=1)[:, :h]
    assert np.array_equal(got, want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from simplestereo_b200.synth import synth_pair
        left, right, _ = synth_pair(96, 21, 12, seed=5)
        kw = dict(winSize=7, maxDisparity=12, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
        h, w = left.shape[:2]
        # --- row stripes: each rank computes its stripe, one all_gather reassembles the map
        s = -(-h // world)
        r0, r1 = row_stripes(h, world)[rank]
        stripe = torch.zeros((s, w), dtype=torch.int16)
        stripe[: r1 - r0] = torch.from_numpy(oracle.asw(left, right, rows=(r0, r1), **kw)[r0:r1])
        full = gather_rows(stripe, h)
        # --- disparity shards: packed (cost, disparity) keys, all_gather + unsigned min
        d0, d1 = disparity_shards(0, 12, world)[rank]
        st = oracle.asw(left, right, stages=True, cost=True, **kw)
        cost = st["cost"].astype(np.float32)
        keys = np.full((h, w), np.iinfo(np.uint64).max, np.uint64)
        for d in range(d0, d1 + 1):
            c = cost[:, :, d]
            k = (c.view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.uint64(d)
            k[~np.isfinite(c)] = np.iinfo(np.uint64).max
            keys = np.minimum(keys, k)
        mine = torch.from_numpy(keys.view(np.int64).reshape(1, -1).copy())
        allk = torch.empty((world * 1, h * w), dtype=torch.int64)
        dist.all_gather_into_tensor(allk, mine)
        merged = allk.numpy().view(np.uint64).min(axis=0).reshape(h, w)
        left_from_keys = (merged & np.uint64(0xffffffff)).astype(np.int16)
        q.put((rank, full.numpy().copy(), left_from_keys))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_row_stripes_and_disparity_shards():
    import oracle
    from simplestereo_b200.synth import synth_pair
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    left, right, _ = synth_pair(96, 21, 12, seed=5)
    kw = dict(winSize=7, maxDisparity=12, minDisparity=0, gammaC=5, gammaP=17.5, consistent=True)
    want = oracle.asw(left, right, stages=True, **kw)
    for rank, full, left_from_keys in results:
        assert np.array_equal(full, want["final"]), f"rank {rank}: gathered stripes differ from the full-frame map"
        # float32 keys can only flip exact float64 near-ties; on this pair they do not
        assert (left_from_keys == want["left"]).mean() > 0.999
