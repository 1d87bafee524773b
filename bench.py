#!/usr/bin/env python3
"""
bench.py -- headline benchmark of the ASW hot path (BASELINE.json metric: Mpix*disparities/s, ASW 35x35).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[1] -- KITTI-size 1242x375 synthetic rectified pair
(simplestereo_b200.synth.synth_pair, seed 0), 128 disparities (0..127), StereoASW winSize=35 gammaC=5 gammaP=17.5.
A "step" is one full disparity map of that pair.

  value   Mpix*disp/s = W*H*D / t, inputs already resident in HBM, device-timed (CUDA events per step on the
          launching stream, L2 flushed between steps, max over ranks).
  e2e     the same metric through the public API StereoASW.compute(left, right): pinned HOST numpy arrays in,
          host int16 map out, H2D + kernels + D2H inside the timed region.
  roofline  the aggregation kernel against the MEASURED FP32 FFMA peak of this GPU (the path is FP32-issue bound,
          SURVEY.md 8d), with achieved HBM GB/s against MEASURED_PEAKS.json beside it.
  cpu_baseline  the unmodified reference (oracle/_ref) timed on this box's host cores on a bounded full-width
          crop and extrapolated by exact window-element counts (N=1, rank 0 only).
  other_configs  device-resident ms of the other BASELINE.json configurations on one GPU (C2 + L-R, C3 GSW, C4, C5; N=1),
          and at N > 1 config C5 (4K, 512 disparities) under BOTH partitions SURVEY.md 8(e) names -- image-row stripes and
          disparity-range shards -- with the flag that all three maps (unsharded, rows, disparity) are bit-identical.
          --no-other-configs skips them.

At N > 1 the frame is sharded into N image-row stripes (one rank per GPU) and reassembled with one NCCL
all-gather (strong scaling: total work fixed); the e2e leg then runs the class API itself,
StereoASW(devices=[0..N-1]).compute(left, right), in rank 0's process: ss_init_devices shards the rows over the N GPUs
with one host thread per GPU (the other ranks wait at the barrier).

--impl reference times the reference's own CPU implementation (oracle/_ref, else the C port) on the same config: every
step is a bounded full-width crop (ms_per_step is its MEASURED duration, value the full-frame-equivalent throughput), and once
per invocation the WHOLE C2 frame is timed (cpu_baseline.full_frame_s) to pin the extrapolation.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, MAXD, MIND, WIN, GC, GP = 1242, 375, 127, 0, 35, 5.0, 17.5
D = MAXD - MIND + 1
METRIC = "Mpix*disparities/s (ASW 35x35)"
UNIT = "Mpix*disp/s"
WORKLOAD = "KITTI-size 1242x375 synthetic rectified pair, 128 disp (0..127), StereoASW winSize=35 gammaC=5 gammaP=17.5"


def mpixdisp(seconds):
    return W * H * D / seconds / 1e6


# ----------------------------------------------------------------------------------------------------
# CPU reference arm
# ----------------------------------------------------------------------------------------------------

def window_elements(width, height, win, min_d, max_d):
    """Exact number of (x, d, i, j) window elements the reference evaluates in its left pass
    (_passive.cpp:56-84): rows ii in-image, right column jj >= 0, left column kk < W."""
    p = win // 2
    y = np.arange(height)
    nrows = np.minimum(y, p) + 1 + np.minimum(height - 1 - y, p)          # in-image window rows per output row
    total_cols = 0
    x = np.arange(width)
    for d in range(min_d, max_d + 1):
        xs = x[x >= d]
        if xs.size == 0:
            break
        lo = np.maximum(0, p - (xs - d))            # first j with jj = x-d-p+j >= 0
        hi = np.minimum(win - 1, width - 1 - xs + p)  # last j with kk = x-p+j <= W-1
        total_cols += int(np.maximum(hi - lo + 1, 0).sum())
    return int(nrows.sum()) * total_cols


def cpu_reference_sample(target_s, crop_rows=None):
    """Time the reference on a full-width crop taken from the middle of the C2 pair and extrapolate to the
    full frame by exact element counts.  Returns (Mpix*disp/s full-frame equivalent, info dict)."""
    import oracle
    from simplestereo_b200.synth import synth_pair
    left, right, _ = synth_pair(W, H, MAXD, 0)
    cores = len(os.sched_getaffinity(0))
    use_ref = oracle.ref_available()
    if crop_rows is None:
        # survey anchor: 18.5 ns per window element per core for the unmodified reference
        ns_per_el = 18.5 if use_ref else 1.0
        per_row = window_elements(W, 1, WIN, MIND, MAXD) * ns_per_el * 1e-9 / max(cores, 1)   # one window row
        crop_rows = cores
        while crop_rows + cores <= H and window_rows(crop_rows + cores) * per_row <= target_s:
            crop_rows += cores
    y0 = (H - crop_rows) // 2
    lc = np.ascontiguousarray(left[y0:y0 + crop_rows])
    rc = np.ascontiguousarray(right[y0:y0 + crop_rows])
    if use_ref:
        _, dt = oracle.ref_asw(lc, rc, WIN, MAXD, MIND, GC, GP, False, with_time=True, timeout=1800)
        kind = "reference"
    else:
        oracle.build(ref=False)
        t0 = time.perf_counter()
        oracle.asw(lc, rc, WIN, MAXD, MIND, GC, GP, False)
        dt = time.perf_counter() - t0
        kind = "port"
    scale = window_elements(W, H, WIN, MIND, MAXD) / window_elements(W, crop_rows, WIN, MIND, MAXD)
    t_full = dt * scale
    info = {"kind": kind, "cores": cores, "crop_rows": crop_rows, "sample_s": dt,
            "sample": f"full-width {W}x{crop_rows} crop of the workload pair ({dt:.2f} s measured), extrapolated x{scale:.2f} "
                      f"to {W}x{H} by exact window-element counts"}
    return mpixdisp(t_full), info, t_full


def window_rows(h):
    p = WIN // 2
    y = np.arange(h)
    return int((np.minimum(y, p) + 1 + np.minimum(h - 1 - y, p)).sum())


def cpu_reference_full_frame():
    """One call of the reference on the WHOLE C2 frame ("C1, C2: time the full image").  Seconds."""
    import oracle
    from simplestereo_b200.synth import synth_pair
    left, right, _ = synth_pair(W, H, MAXD, 0)
    if oracle.ref_available():
        _, dt = oracle.ref_asw(left, right, WIN, MAXD, MIND, GC, GP, False, with_time=True, timeout=3600)
        return dt
    oracle.build(ref=False)
    t0 = time.perf_counter()
    oracle.asw(left, right, WIN, MAXD, MIND, GC, GP, False)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = max(args.steps + args.warmup, 1)
    target = min(30.0, max(2.0, 150.0 / total))
    vals, times, samples = [], [], []
    info = None
    rows = None
    for k in range(args.warmup + args.steps):
        v, info, t_full = cpu_reference_sample(target, rows)
        rows = info["crop_rows"]
        if k >= args.warmup:
            vals.append(v)
            times.append(t_full)
            samples.append(info["sample_s"])
    value = statistics.median(vals)
    t_equiv = statistics.median(times)
    full = None
    if not args.no_full_frame:
        full_s = cpu_reference_full_frame()
        full = {"full_frame_s": full_s, "full_frame_value": mpixdisp(full_s),
                "extrapolation_error": (t_equiv - full_s) / full_s}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": statistics.median(samples) * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "a step is one bounded sample (full-width crop) and ms_per_step its measured duration; value is the "
                           "full-frame-equivalent throughput (exact window-element counts), pinned by cpu_baseline.full_frame_s",
                   "full_frame_ms_equivalent": t_equiv * 1e3},
        "cpu_baseline": {"value": value, "unit": UNIT, **info, **(full or {})},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------

class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every 10 ms
    from a thread (nvidia_ml_py), falling back to `nvidia-smi -lms` when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, uuid=None):
        self.index = index
        self.uuid = uuid
        self.rows = []          # (sm_mhz, reasons bitmask or list of names)
        self.max_mhz = None
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        if self.uuid:
            for cand in (f"GPU-{self.uuid}", str(self.uuid)):
                try:
                    return pynvml, pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if hasattr(cand, "encode") else cand)
                except Exception:
                    continue
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                     ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                     ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                        self.rows.append((mhz, [n for n, bit in names if mask & bit]))
                    except Exception:
                        pass
                    self.stop_flag.wait(0.01)

            self.source = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9 or not c[1].replace(".", "").isdigit():
                continue
            if c[2].replace(".", "").isdigit():
                self.max_mhz = max(self.max_mhz or 0.0, float(c[2]))
            reasons = [n for n, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8))
                       if c[col].lower().startswith("active")]
            self.rows.append((float(c[1]), reasons))

    def mark(self):
        """Index of the next sample: brackets the timed region inside a longer sampling run."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        rows = self.rows[first:last] or self.rows
        sm = [r[0] for r in rows]
        reasons = sorted({n for r in rows for n in r[1]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": self.source}


def c5_partitions(ss, _cabi, L, dev, stream, world, rank, barrier):
    """BASELINE.json config 5 (3840x2160, 512 disparities, win 35) over `world` GPUs: image-row stripes and disparity-range
    shards (chunk-aligned groups x row stripes when ranks outnumber the four 128-disparity chunks), device-timed, max over
    ranks; plus the unsharded call on rank 0 and whether the three maps are bit-identical."""
    import torch
    import torch.distributed as dist
    from simplestereo_b200.sharding import ShardedStereoASW, disparity_row_grid
    from simplestereo_b200.synth import synth_pair
    w, h, win, maxd = 3840, 2160, 35, 511
    l2, r2, _ = synth_pair(w, h, maxd, 0)
    a, b = torch.from_numpy(l2).to(dev), torch.from_numpy(r2).to(dev)
    m = ss.passive.StereoASW(win, maxd, 0, GC, GP, consistent=False)
    out = {}
    maps = {}
    for mode in ("rows", "disparity"):
        sh = ShardedStereoASW(m, mode=mode)
        maps[mode] = sh.compute_device(a, b).clone()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sh.compute_device(a, b)
        sh.compute_device(a, b)
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / 2], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[mode] = {"ms": round(float(t.item()), 3), "Mpix_disp_per_s": round(w * h * (maxd + 1) / float(t.item()) / 1e3, 1)}
        del sh
    n_d, n_r, _ = disparity_row_grid(0, maxd, h, world)
    out["disparity"]["grid"] = f"{n_d} disparity groups (chunk-aligned) x {n_r} row stripes"
    same = torch.tensor([1 if torch.equal(maps["rows"], maps["disparity"]) else 0], device=dev)
    if rank == 0:
        full = torch.empty((h, w), dtype=torch.int16, device=dev)

        def run():
            _cabi.check(L.ss_asw_compute_device(a.data_ptr(), b.data_ptr(), w, h, win, maxd, 0, GC, GP, 0, 0, h, full.data_ptr(),
                                                stream.cuda_stream))
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run()
        e1.record(stream)
        torch.cuda.synchronize()
        out["one_gpu"] = {"ms": round(e0.elapsed_time(e1), 3)}
        same *= 1 if torch.equal(full, maps["rows"]) else 0
        del full
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    barrier()
    out["maps_bit_identical"] = bool(int(same.item()))
    if "one_gpu" in out:
        for mode in ("rows", "disparity"):
            out[mode]["speedup_vs_one_gpu"] = round(out["one_gpu"]["ms"] / out[mode]["ms"], 2)
    del a, b, maps
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist
    import simplestereo_b200 as ss
    from simplestereo_b200 import _cabi
    from simplestereo_b200.sharding import ShardedStereoASW
    from simplestereo_b200.synth import synth_pair

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (simplestereo_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = _cabi.lib()
    _cabi.check(L.ss_init(local))
    cpu_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        # host-side barrier for the e2e leg: while rank 0 drives all N GPUs from its own process the other ranks must not sit in
        # an NCCL barrier kernel on those GPUs (two processes time-slice a GPU: measured 11.8 ms instead of 4.6 ms at N=2)
        cpu_group = dist.new_group(backend="gloo")

    left, right, _ = synth_pair(W, H, MAXD, 0)
    matcher = ss.passive.StereoASW(WIN, MAXD, MIND, GC, GP, consistent=False)
    sharded = ShardedStereoASW(matcher, mode="rows") if world > 1 else None

    # pinned host staging for the end-to-end leg
    h_left = torch.from_numpy(left).pin_memory()
    h_right = torch.from_numpy(right).pin_memory()
    d_left = h_left.to(dev)
    d_right = h_right.to(dev)
    d_out = torch.empty((H, W), dtype=torch.int16, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        if world > 1:
            return sharded.compute_device(d_left, d_right)
        _cabi.check(L.ss_asw_compute_device(d_left.data_ptr(), d_right.data_ptr(), W, H, WIN, MAXD, MIND, GC, GP, 0, 0, H,
                                            d_out.data_ptr(), stream.cuda_stream))
        return d_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step_device()
    barrier()

    # ---- device-timed leg -------------------------------------------------------------------------
    L.ss_profile_reset()
    L.ss_profile_enable(1)
    try:
        uuid = str(torch.cuda.get_device_properties(local).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local, uuid)
    if rank == 0:
        sampler.start()
    evs = []
    barrier()
    s_first = sampler.mark()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(0)                        # L2 flush between timed iterations (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        result = step_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop(s_first, sampler.mark()) if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    agg_ms, agg_launches, total_launches = _cabi.profile_read()
    L.ss_profile_enable(0)
    ms_per_step = total_ms / args.steps
    value = mpixdisp(ms_per_step * 1e-3)

    # ---- end-to-end leg: the class API a user calls, pinned host buffers in, host map out -----------------
    # N = 1: StereoASW.compute on this rank's GPU.  N > 1: rank 0 alone runs StereoASW(devices=[0..N-1]).compute, which
    # shards the image rows over the N GPUs inside its process (ss_init_devices: one host thread and context per GPU, H2D of
    # each stripe's rows, kernels, D2H of each stripe into the caller's array); the other ranks wait at the barrier.
    np_left, np_right = h_left.numpy(), h_right.numpy()
    e2e_matcher = matcher if world == 1 else ss.passive.StereoASW(WIN, MAXD, MIND, GC, GP, consistent=False, devices=list(range(world)))
    out_host = None

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    barrier()
    host_barrier()
    if rank == 0:
        for _ in range(2):
            e2e_matcher.compute(np_left, np_right)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out_host = e2e_matcher.compute(np_left, np_right)
        e2e_s = (time.perf_counter() - t0) / args.steps
        if world > 1:
            _cabi.use_devices([local])
    else:
        e2e_s = 0.0
    host_barrier()
    barrier()

    # sanity: device leg and e2e leg produce the same map
    same = bool(np.array_equal(np.asarray(out_host), result.cpu().numpy()[:H])) if rank == 0 else None

    # ---- config C5 (4K, 512 disparities) under both multi-GPU partitions (N > 1; every rank takes part) --------------
    c5 = None
    if world > 1 and not args.no_other_configs:
        c5 = c5_partitions(ss, _cabi, L, dev, stream, world, rank, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the aggregation kernel -------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    fp32_peak = _cabi.measure_fp32_peak(stream.cuda_stream)
    rows_per_rank = -(-H // world)
    flops_per_launch = 4.0 * W * rows_per_rank * D * WIN * WIN          # 4 flop per nominal window element
    bytes_per_launch = 8.0 * W * rows_per_rank                          # compulsory HBM bytes: 2*3 in + 2 out per pixel
    agg_avg_s = (agg_ms / max(agg_launches, 1)) * 1e-3
    achieved_tf = flops_per_launch / agg_avg_s / 1e12
    traffic = None
    traffic_note = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        traffic = prof.get("k_aggregate_dram_bytes_per_launch")
        traffic_note = prof.get("source")
        if traffic is not None and world > 1:
            # the capture is of the one-GPU launch (all 375 rows); a rank's launch covers rows_per_rank rows plus the halo
            traffic = traffic * (rows_per_rank + WIN - 1) / (H + WIN - 1)
            traffic_note = f"{traffic_note}; scaled to this rank's {rows_per_rank} rows (+{WIN - 1} halo rows)"
    except Exception:
        pass
    pipe_flops = 3.0 * W * rows_per_rank * D * WIN * WIN               # mul + fma: what the CUDA cores execute per element
    roofline = {
        "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
        "traffic": traffic, "traffic_source": traffic_note,
        "kernel": "k_aggregate_tc<ASW, DC=128> (warp-specialised support-weight aggregation + WTA for both references; numerators on packed FP32, denominators on tcgen05 3xTF32)", "kernel_ms": agg_avg_s * 1e3, "kernel_share_of_step": agg_ms / total_ms if world == 1 else None,
        "peak_source": "FFMA-only microbenchmark run live on this GPU (ss_measure_fp32_peak); nominal 74.4 TFLOP/s",
        "algorithmic_flops_per_launch": flops_per_launch,
        "note": "frac counts the 4 ALGORITHMIC flop per window element (mul, fma, add; SURVEY.md 8d) against the FFMA peak; "
                "the add (denominator) runs on the tensor cores (tcgen05 kind::tf32, 3xTF32), so the FP32 pipe itself executes "
                "3 flop per element: fp32_pipe is that utilisation",
        "fp32_pipe": {"flops_per_element": 3, "achieved": pipe_flops / agg_avg_s / 1e12, "frac": pipe_flops / agg_avg_s / 1e12 / fp32_peak},
        "hbm": {"achieved": bytes_per_launch / agg_avg_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_per_launch / agg_avg_s / 1e9 / hbm_peak, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": bytes_per_launch},
    }

    # the other BASELINE.json configurations, device-resident, for context (parity for them lives in tests/)
    others = None
    if world == 1 and not args.no_other_configs:
        others = {}
        specs = {"C2+LR 1242x375 D128 win35 consistent": (1242, 375, 35, 127, "asw", 1),
                 "C3 GSW 1242x375 D128 win35": (1242, 375, 35, 127, "gsw", 1),
                 "C4 2880x1988 D256 win51 consistent": (2880, 1988, 51, 255, "asw", 1),
                 "C5 3840x2160 D512 win35 (one GPU)": (3840, 2160, 35, 511, "asw", 0)}
        for name, (w, h, win, maxd, kind, cons) in specs.items():
            try:
                l2, r2, _ = synth_pair(w, h, maxd, 0)
                a, b = torch.from_numpy(l2).to(dev), torch.from_numpy(r2).to(dev)
                o = torch.empty((h, w), dtype=torch.int16, device=dev)

                def run():
                    if kind == "asw":
                        _cabi.check(L.ss_asw_compute_device(a.data_ptr(), b.data_ptr(), w, h, win, maxd, 0, GC, GP, cons, 0, h,
                                                            o.data_ptr(), stream.cuda_stream))
                    else:
                        _cabi.check(L.ss_gsw_compute_device(a.data_ptr(), b.data_ptr(), w, h, win, maxd, 0, 10, 120.0, 3, 20, 0, h,
                                                            o.data_ptr(), stream.cuda_stream))
                run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                run()
                run()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 2
                others[name] = {"ms": round(ms, 3), "Mpix_disp_per_s": round(w * h * (maxd + 1) / ms / 1e3, 1)}
                del a, b, o
            except Exception as ex:
                others[name] = {"error": repr(ex)[:120]}
        torch.cuda.empty_cache()
    if c5 is not None:
        others = {"C5 3840x2160 D512 win35 sharded over %d GPUs" % world: c5}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, info, _ = cpu_reference_sample(15.0)
            cpu = {"value": v, "unit": UNIT, **info}
        except Exception as ex:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": UNIT, "error": repr(ex)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sharding": "none" if world == 1 else f"{world} image-row stripes + one NCCL all_gather",
                   "e2e_path": "StereoASW.compute (one GPU)" if world == 1 else f"StereoASW(devices=[0..{world - 1}]).compute in one process: rows sharded over {world} GPUs, one host thread each, no collective",
                   "l2": "flushed between timed steps (256 MiB fill)", "timing": "CUDA events per step on the launching stream, max over ranks",
                   "maps_equal_across_legs": same},
        "clocks": clocks,
        "e2e": {"value": mpixdisp(e2e_s), "unit": UNIT, "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": 2 * W * H * 3, "d2h_bytes_per_step": W * H * 2},
        "gpu_launches": int(total_launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "other_configs": others,
        "wall_s_timed_region": t_wall,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-full-frame", action="store_true", help="reference arm: skip the one full-frame call (~45 s on 16 cores)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
