#!/usr/bin/env python
"""bench.py -- KLT feature-tracks/s of the B200-native pyramid Gauss-Newton KLT path.

Metric (BASELINE.json): KLT feature-tracks/sec, 4-level pyramid, at 1/2/4/8 B200, with the reference's CPU path timed
beside it, and pyramid GB/s.

Headline workload (config C3, the batched multi-GPU one the metric is quoted on): per GPU, B=256 independent synthetic
KITTI-shaped stereo pairs 1241x376 x 2000 features, 4 levels, the reference's 7x7 patch (src/algorithm.cpp:40:
half_patch_size=3 -- SURVEY.md F1; "8x8" in the metric text is not what the reference computes), forward mode, initial
guess kp2 = kp1.  A step = pyramids of all 512 images + the fused 4-level solver over 512,000 features.  Weak scaling:
every rank owns its own 256 pairs, no data-path collective, final gather of (x,y,flag) outside the timed region.

Beside the headline the line carries side records for the other BASELINE configs (C1 single call, C2 sequence, C4
1080p / 5 levels / both modes, C5 feature-count and patch sweep), sub-pixel source keypoints, a >= 1 s sustained run,
the host-to-device copy ceiling of this host and this run's own parity report against the CPU arm.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS, LEVELS = 376, 1241, 4
PATCH_LO, PATCH_HI = -3, 3
FLOP_PER_PIXEL_ITER = 123          # SURVEY.md 8d, reference formulation, forward mode (and inverse mode, first pass)
FLOP_PER_PIXEL_ITER_INV = 43       # inverse mode, passes after the first of a level
PYR_BYTES_PER_IMAGE = 619_601      # level 0 read once + levels 1..3 written once
METRIC = "KLT feature-tracks/sec (4-level pyramid, 1241x376, 2000 features/pair, batched pairs)"
DTYPE = "f32 sampling + f64 normal equations"


# ---------------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------------
def _make_pair(args):
    seed, n, rows, cols = args
    from lego_slam_b200 import synth
    L, R, kp1, kp2, truth = synth.stereo_case(rows, cols, n, seed=seed)
    return L, R, kp1, kp2, truth


def make_workload(n_feat: int, distinct: int, seed0: int, rows: int = ROWS, cols: int = COLS):
    """`distinct` generated pairs (seeds seed0..); fill_batch tiles them to the batch size."""
    jobs = [(seed0 + i, n_feat, rows, cols) for i in range(max(1, distinct))]
    workers = min(len(jobs), max(1, (os.cpu_count() or 2) // 2), 32)
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            return pool.map(_make_pair, jobs)
    return [_make_pair(j) for j in jobs]


def fill_batch(base, n_pairs, n_feat, alloc, rows: int = ROWS, cols: int = COLS):
    """Tiles the generated pairs cyclically; odd repeats are vertically flipped so that repeated slots are not
    byte-identical.  Every pair takes the first n_feat features of its base pair."""
    imgs1 = alloc((n_pairs, rows, cols), np.uint8)
    imgs2 = alloc((n_pairs, rows, cols), np.uint8)
    kp1 = alloc((n_pairs, n_feat, 2), np.float32)
    kp2 = alloc((n_pairs, n_feat, 2), np.float32)
    for b in range(n_pairs):
        L, R, a, g = base[b % len(base)][:4]
        a, g = a[:n_feat], g[:n_feat]
        if (b // len(base)) % 2 == 1:  # vertical flip: still a valid stereo pair
            imgs1[b], imgs2[b] = L[::-1], R[::-1]
            f = a.copy()
            f[:, 1] = (rows - 1) - f[:, 1]
            kp1[b] = f
            kp2[b] = f
        else:
            imgs1[b], imgs2[b], kp1[b], kp2[b] = L, R, a, g
    return imgs1, imgs2, kp1, kp2


def fill_truth(base, n_pairs, n_feat, rows: int = ROWS):
    """Ground-truth right-image positions of fill_batch's features (the generator's analytic disparity)."""
    out = np.empty((n_pairs, n_feat, 2), np.float32)
    for b in range(n_pairs):
        t = base[b % len(base)][4][:n_feat]
        if (b // len(base)) % 2 == 1:
            t = t.copy()
            t[:, 1] = (rows - 1) - t[:, 1]
        out[b] = t
    return out


def accuracy_vs_truth(kp, succ, truth):
    """Distance of the tracked positions from the generator's ground truth, over the features flagged successful."""
    ok = succ.astype(bool)
    if not ok.any():
        return None
    d = np.linalg.norm(kp.astype(np.float64) - truth, axis=2)[ok]
    return {"median_px": float(np.median(d)), "frac_below_0.5_px": float((d < 0.5).mean()), "features": int(ok.sum())}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING a timed region through NVML (the same counters
    `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints), polled every ~1 ms in a thread."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, cuda_index: int):
        self.samples, self.mask, self.h, self.max_mhz, self.err = [], 0, None, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.001)

    def start(self):
        if self.h is None:
            return self
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()
        return self

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self._stop.set()
        self.t.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# the CPU arm: the reference's own translation unit (oracle/_ref) when it was built, else the oracle port
# ---------------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def quiet_stdout_fd():
    """The reference prints "Update is NaN or INF." to std::cout (src/algorithm.cpp:97): keep file descriptor 1 clean
    for the JSON line while its code runs."""
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


class CpuArm:
    """LKOpticalFlow4Layer on the host cores, one pair per thread at a time (the reference's own cv::parallel_for_
    splits the features of ONE call; with hundreds of independent pairs the pair-parallel form keeps every core busy
    the same way and has no fork/join per level)."""

    def __init__(self, levels=LEVELS, half_patch=3, inverse=False):
        from oracle import ref_binding as rb
        self.levels, self.half_patch, self.inverse = levels, half_patch, inverse
        self.use_ref = rb.available(half_patch, levels)
        if self.use_ref:
            try:
                rb.lib(half_patch, levels)
            except Exception:
                self.use_ref = False
        self.kind = "reference" if self.use_ref else "port"
        verb = "" if (half_patch, levels) == (3, 4) else " with the patch / level literals substituted"
        self.what = (f"oracle/_ref: the reference's own src/algorithm.cpp + algorithm.h{verb}, -std=c++11 -O3, on stand-in "
                     "cv/Eigen headers" if self.use_ref else "oracle port (oracle/_ref not built on this box), -std=c++11 -O3")

    def track_pairs(self, imgs1, imgs2, kp1, kp2, threads):
        if self.use_ref:
            from oracle import ref_binding as rb
            with quiet_stdout_fd():
                return rb.track_pairs(np.ascontiguousarray(imgs1), np.ascontiguousarray(imgs2), kp1, kp2, threads,
                                      inverse=self.inverse, half_patch=self.half_patch, pyramids=self.levels)
        from concurrent.futures import ThreadPoolExecutor
        from oracle import binding as ob
        p = ob.make_params(levels=self.levels, patch_lo=-self.half_patch, patch_hi=self.half_patch, inverse=self.inverse)
        B = imgs1.shape[0]
        out = np.empty((B,) + kp1.shape[1:], np.float32)
        ok = np.empty((B, kp1.shape[1]), np.uint8)

        def one(b):
            out[b], ok[b], _ = ob.track(np.ascontiguousarray(imgs1[b]), np.ascontiguousarray(imgs2[b]), kp1[b], kp2[b], p)

        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(one, range(B)))
        return out, ok

    def throughput(self, imgs1, imgs2, kp1, kp2, budget_s: float, threads: int):
        """tracks/s on a bounded sample: blocks of 2*threads pairs of the workload until the budget is spent."""
        B, n = kp1.shape[0], kp1.shape[1]
        blk = min(B, 2 * threads)
        t0 = time.perf_counter()
        self.track_pairs(imgs1[:blk], imgs2[:blk], kp1[:blk], kp2[:blk], threads)
        t_blk = time.perf_counter() - t0
        reps = int(max(1, min(400, budget_s / max(t_blk, 1e-3))))
        t0 = time.perf_counter()
        for r in range(reps):
            lo = (r * blk) % max(B - blk + 1, 1)
            self.track_pairs(imgs1[lo:lo + blk], imgs2[lo:lo + blk], kp1[lo:lo + blk], kp2[lo:lo + blk], threads)
        dt = time.perf_counter() - t0
        return reps * blk * n / dt, reps * blk, dt


def workload_config(args):
    return {"workload": f"C3: {args.pairs} independent stereo pairs {COLS}x{ROWS} u8 per GPU x {args.features} "
                        f"features, {LEVELS}-level pyramid, 7x7 patch (reference half_patch_size=3), forward, kp2=kp1",
            "pairs_per_gpu": args.pairs, "features_per_pair": args.features, "levels": LEVELS,
            "patch": [PATCH_LO, PATCH_HI], "distinct_pairs": min(args.distinct, args.pairs),
            "l2_policy": "inputs larger than L2 (level-0 images of one step: %.0f MB > 126 MB)"
                         % (2 * args.pairs * ROWS * COLS / 1e6),
            "sharding": "block partition of pairs, one process per GPU, no data-path collective",
            "batches_in_flight": max(1, args.streams), "subpixel_keypoints": bool(args.subpixel)}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on all host threads -- oracle/_ref, i.e.
    the reference's own translation unit compiled on stand-in headers (oracle/build_ref.py), when it travelled to this
    box, else the oracle port.  Each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    arm = CpuArm()
    base = make_workload(args.features, min(args.distinct, 16), 1000)
    blk = 2 * threads
    imgs1, imgs2, kp1, kp2 = fill_batch(base, blk, args.features, lambda s, d: np.empty(s, d))
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, sample = [], None
    for s in range(args.warmup + args.steps):
        v, n_pairs, dt = arm.throughput(imgs1, imgs2, kp1, kp2, per_step_budget, threads)
        if s >= args.warmup:
            vals.append((v, dt))
        sample = f"{n_pairs} pairs x {args.features} features per step ({dt:.1f} s), full pyramids + {LEVELS} levels; {arm.what}"
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": threads, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(cuda_index: int):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch puts them on that node).  Returns a short description (or the reason it did not)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:       # "00000000:1b:00.0" -> sysfs uses a 4-digit domain
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return "no local cpus in the affinity mask"
        os.sched_setaffinity(0, cpus)
        node = open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip()
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({e!r})"


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def flops_of(iters, n_levels_features, patch_w, inverse):
    """Algorithmic flop on the REFERENCE formulation: passes x pixels x flop per pixel-pass."""
    P = patch_w * patch_w
    total = int(sum(iters))
    if not inverse:
        return total * P * FLOP_PER_PIXEL_ITER
    first = min(total, int(n_levels_features))   # one first pass per (feature, level)
    return first * P * FLOP_PER_PIXEL_ITER + (total - first) * P * FLOP_PER_PIXEL_ITER_INV


class Timer:
    """CUDA-event timing of a region on one stream (torch events see only torch streams: the library runs on the
    stream handed to lego_klt_set_stream)."""

    def __init__(self, torch, stream):
        self.torch, self.stream = torch, stream
        self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        self.e0.record(self.stream)
        return self

    def __exit__(self, *a):
        self.e1.record(self.stream)
        self.torch.cuda.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import lego_slam_b200 as klt
    from lego_slam_b200 import build, sharding, synth
    build.build()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the KLT path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()

    def reduce_max(values):
        t = torch.tensor(list(values), dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather_all(value):
        t = torch.tensor([float(value)], dtype=torch.float64, device=f"cuda:{local}")
        if world == 1:
            return [float(value)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o[0]) for o in out]

    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    B, n = args.pairs, args.features
    base = make_workload(n, min(args.distinct, B), 1000 + rank * B)
    imgs1, imgs2, kp1, kp2 = fill_batch(base, B, n, klt.pinned_empty)
    truth = fill_truth(base, B, n)
    if args.subpixel:
        kp1 += np.random.default_rng(77 + rank).uniform(-0.5, 0.5, kp1.shape).astype(np.float32)
        np.copyto(kp2, kp1)
    kp2_io = klt.pinned_empty((B, n, 2), np.float32)
    succ = klt.pinned_empty((B, n), np.uint8)
    params = klt.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI, kernel=args.kernel)
    n_tracks = B * n
    # `--streams S` batches in flight: S device-resident batch objects (same inputs), each on its own stream, stepped
    # round-robin -- the small kernels of one batch (pyramid, templates) fill the issue slots the tail of the other
    # batch's persistent solver kernel leaves idle.
    S = max(1, args.streams)
    trks = [klt.Tracker(local) for _ in range(S)]
    streams = [torch.cuda.Stream(device=local) for _ in range(S)]
    batches = []
    for t, st_ in zip(trks, streams):
        t.set_stream(st_.cuda_stream)
        batches.append(t.batch(B, ROWS, COLS, n, levels=LEVELS))
    trk, stream, batch = trks[0], streams[0], batches[0]

    def resident_steps(count, p=params):
        """`count` device-resident steps, round-robin over the batches in flight, timed on the device."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(e0)
        for i in range(count):
            batches[i % S].run(p)
        for st_ in streams[1:]:
            streams[0].wait_stream(st_)
        e1.record(streams[0])
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---------------- device-resident: inputs already in HBM, results stay in HBM ----------------
    for b_ in batches:
        b_.upload(imgs1, imgs2, kp1, kp2)
    resident_steps(args.warmup * S)
    for t in trks:
        t.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    launches0 = klt.kernel_launches()
    ms_total_local = resident_steps(args.steps)
    gpu_launches = klt.kernel_launches() - launches0
    barrier()

    # the same step looped for >= ~1.2 s: the K-step region above is a burst of a few tens of ms
    sustained = None
    if not args.no_sustained:
        ms_step_all = reduce_max([ms_total_local / args.steps])[0]     # the same step count on every rank
        n_sus = int(min(4000, max(args.steps, 1200.0 / max(ms_step_all, 1e-3))))
        sus_sampler = ClockSampler(local)
        if rank == 0:
            sus_sampler.start()
        barrier()
        ms_sus_local = resident_steps(n_sus)
        barrier()
        sus_clocks = sus_sampler.stop() if rank == 0 else None
        sustained = (n_sus, ms_sus_local, sus_clocks)

    # per-kernel launch durations: CUDA events recorded inside the library around each kernel group on the
    # launching stream, over the same number of steps run back to back on ONE stream (with several batches
    # in flight the brackets of one batch would include the other batch's kernels)
    for _ in range(args.steps):
        batch.run(params)
    ms_pyr, ms_sol = batch.timings(min(args.steps, 64))
    _, _, st = batch.download(kp2_io, succ)
    iters = [int(v) for v in st.gn_iters][:LEVELS]
    slow, deferred, n_success = int(st.n_slow_path), int(st.n_deferred), int(st.n_success)
    defer_reason = [int(v) for v in st.defer_reason]
    resident_kp, resident_succ = kp2_io.copy(), succ.copy()

    # ---------------- end to end: host buffers -> lego_klt_track_batched -> host buffers ----------------
    # kp2 is in/out (initial guess in, tracked position out): every timed step gets its own pre-filled pinned
    # buffer, so that no host-side refill of the guess sits inside the timed region.
    n_ring = max(2, args.steps) if args.steps <= 32 else 2
    kp2_ring = [kp2_io] + [klt.pinned_empty((B, n, 2), np.float32) for _ in range(n_ring - 1)]
    for _ in range(min(args.warmup, 3)):
        np.copyto(kp2_io, kp2)
        batch.track(imgs1, imgs2, kp1, kp2_io, succ, params)
    for buf in kp2_ring:
        np.copyto(buf, kp2)
    barrier()
    torch.cuda.synchronize()
    # (1) one call in flight: lego_klt_track_batched, synchronous
    with Timer(torch, stream) as tm:
        for i in range(args.steps):
            if n_ring < args.steps:
                np.copyto(kp2_ring[i % n_ring], kp2)
            batch.track(imgs1, imgs2, kp1, kp2_ring[i % n_ring], succ, params)   # returns after the D2H
    e2e_sync_ms_local = tm.ms
    e2e_kp, e2e_succ = kp2_ring[(args.steps - 1) % n_ring].copy(), succ.copy()
    barrier()
    # (2) two calls in flight: lego_klt_track_batched_begin / _end on two batch objects (two contexts, two stream sets):
    # the upload of step i + 1 starts while step i still computes and copies its results back -- the copy engine never idles
    if S >= 2:
        eb, est = [batches[0], batches[1]], [streams[0], streams[1]]
    else:
        t2 = klt.Tracker(local)
        s2 = torch.cuda.Stream(device=local)
        t2.set_stream(s2.cuda_stream)
        eb, est = [batch, t2.batch(B, ROWS, COLS, n, levels=LEVELS)], [stream, s2]
        trks.append(t2)
    succ2 = [succ, klt.pinned_empty((B, n), np.uint8)]
    for buf in kp2_ring:
        np.copyto(buf, kp2)
    eb[1].track(imgs1, imgs2, kp1, kp2_ring[0], succ2[1], params)    # (warm-up of the second object)
    np.copyto(kp2_ring[0], kp2)
    def e2e_pipelined(src_kp1, guess, steps):
        """`steps` end-to-end calls, two in flight; device time between the first enqueue and the last result copy."""
        ring = kp2_ring if steps <= len(kp2_ring) else kp2_ring[:2]
        for buf in ring:
            np.copyto(buf, guess)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(est[0])
        est[1].wait_event(e0)
        pending = [False, False]
        for i in range(steps):
            j = i % 2
            if pending[j]:
                eb[j].track_end()
                if len(ring) < steps:
                    np.copyto(ring[i % len(ring)], guess)
            eb[j].track_begin(imgs1, imgs2, src_kp1, ring[i % len(ring)], succ2[j], params)
            pending[j] = True
        for j in range(2):
            if pending[j]:
                eb[j].track_end()
        est[0].wait_stream(est[1])
        e1.record(est[0])
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), ring[(steps - 1) % len(ring)], succ2[(steps - 1) % 2]

    barrier()
    t0 = time.perf_counter()
    e2e_ms_local, last_kp, last_succ = e2e_pipelined(kp1, kp2, args.steps)
    e2e_wall_s = time.perf_counter() - t0
    pipe_kp, pipe_succ = last_kp.copy(), last_succ.copy()
    barrier()

    # the same pinned buffers, copy only (H2D of both image sets and the keypoints + the re-pitch kernels), all ranks at
    # once: what this host can deliver to N GPUs at a time -- the ceiling of the end-to-end number
    n_up = max(4, min(args.steps, 10))
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(est[0])
    est[1].wait_event(u0)
    for i in range(n_up):
        eb[i % 2].upload(imgs1, imgs2, kp1, kp2)     # (two objects, two streams: the copies run back to back)
    est[0].wait_stream(est[1])
    u1.record(est[0])
    torch.cuda.synchronize()
    h2d_bytes = int(imgs1.nbytes + imgs2.nbytes + kp1.nbytes + kp2.nbytes)
    h2d_gbs_local = h2d_bytes * n_up / (u0.elapsed_time(u1) * 1e-3) / 1e9
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-resident steps, sustained run, per-kernel timing pass and end-to-end steps"

    # ---------------- next step of the frontend on the tracked batch: triangulation ----------------
    left34 = np.hstack([np.eye(3), np.zeros((3, 1))])
    right34 = np.hstack([np.eye(3), np.array([[-0.537], [0.0], [0.0]])])
    cam_l = klt.make_camera(718.856, 718.856, 607.1928, 185.2157, left34)
    cam_r = klt.make_camera(718.856, 718.856, 607.1928, 185.2157, right34)
    tri_pt = klt.pinned_empty((B, n, 3), np.float64)
    tri_ok = klt.pinned_empty((B, n), np.uint8)
    batch.upload(imgs1, imgs2, kp1, kp2)
    batch.run(params)
    batch.triangulate(cam_l, cam_r, 1e-3, tri_pt, tri_ok)
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        batch.triangulate(cam_l, cam_r, 1e-3, tri_pt, tri_ok)
    tri_ms_local = (time.perf_counter() - t0) / 3 * 1e3

    # ---------------- side records that run on every rank (weak scaling like the headline) ----------------
    side_ms = {}

    def timed_runs(bt, p, reps=3, key=None):
        bt.run(p)
        trk.sync()
        with Timer(torch, stream) as tmr:
            for _ in range(reps):
                bt.run(p)
        if key:
            side_ms[key] = tmr.ms / reps
        return tmr.ms / reps

    if not args.no_side:
        # (a) the literal 8x8 patch of the metric text and the 11x11 stress patch on the headline batch
        for name, (plo, phi) in (("8x8", (-4, 3)), ("11x11", (-5, 5))):
            timed_runs(batch, klt.make_params(levels=LEVELS, patch_lo=plo, patch_hi=phi), key=f"patch_{name}")
        # (b) sub-pixel source keypoints (tracked points fed back by TrackLastFrame, src/frontend_g2o.cpp:453-492)
        kp1s = klt.pinned_empty((B, n, 2), np.float32)
        np.copyto(kp1s, kp1)
        if not args.subpixel:
            kp1s += np.random.default_rng(77 + rank).uniform(-0.5, 0.5, kp1s.shape).astype(np.float32)
        batch.upload(imgs1, imgs2, kp1s, kp1s)
        timed_runs(batch, params, key="subpixel_resident")
        _, _, st_s = batch.download(kp2_io, succ)
        sub_kp, sub_succ = kp2_io.copy(), succ.copy()
        e2e_pipelined(kp1s, kp1s, 2)
        side_ms["subpixel_e2e"] = e2e_pipelined(kp1s, kp1s, 6)[0] / 6
        # (b') the projected-map-point branch of the callers (src/frontend_g2o.cpp:461-463, 504-505): the initial guess is
        # the true position + N(0, 2 px) instead of the source pixel (second variant)
        kp2n = klt.pinned_empty((B, n, 2), np.float32)
        np.copyto(kp2n, truth + np.random.default_rng(99 + rank).normal(0.0, 2.0, truth.shape).astype(np.float32))
        batch.upload(imgs1, imgs2, kp1, kp2n)
        timed_runs(batch, params, key="guess_projected")
        _, _, st_g = batch.download(kp2_io, succ)
        guess_acc = accuracy_vs_truth(kp2_io, succ, truth)
        for b_ in eb[1:] + batches[1:]:     # (memory back before the sweep batches)
            b_.close()
        # (c) C5: feature-count sweep x patch on 64 pairs, per-kernel times
        sweep_B = min(64, B)
        sweep_counts = [100, 500, 2000, 5000, 20000]
        sweep_base = make_workload(max(sweep_counts), 8, 4000 + rank * 8)
        sweep_iters = {}
        for cnt in sweep_counts:
            s1, s2, sk1, sk2 = fill_batch(sweep_base, sweep_B, cnt, klt.pinned_empty)
            sb = trk.batch(sweep_B, ROWS, COLS, cnt, levels=LEVELS)
            sb.upload(s1, s2, sk1, sk2)
            for pname, (plo, phi) in (("7x7", (-3, 3)), ("11x11", (-5, 5))):
                pp = klt.make_params(levels=LEVELS, patch_lo=plo, patch_hi=phi)
                for _ in range(4):
                    sb.run(pp)
                mp_, ms_ = sb.timings(3)
                _, _, sst = sb.download()
                side_ms[f"sweep_{cnt}_{pname}_pyr"], side_ms[f"sweep_{cnt}_{pname}_sol"] = mp_, ms_
                sweep_iters[(cnt, pname)] = [int(v) for v in sst.gn_iters][:LEVELS]
            sb.close()
        # (d) C4: 1920x1080, 5 levels, 5000 features, forward and the reference's inverse mode, 32 pairs
        c4_B, c4_n, c4_rows, c4_cols, c4_L = min(32, B), 5000, 1080, 1920, 5
        c4_base = make_workload(c4_n, 4, 3 + rank * 4, c4_rows, c4_cols)
        h1, h2, hk1, hk2 = fill_batch(c4_base, c4_B, c4_n, klt.pinned_empty, c4_rows, c4_cols)
        cb = trk.batch(c4_B, c4_rows, c4_cols, c4_n, levels=c4_L)
        cb.upload(h1, h2, hk1, hk2)
        c4_iters = {}
        for mode in ("forward", "inverse"):
            pp = klt.make_params(levels=c4_L, inverse=(mode == "inverse"))
            for _ in range(4):
                cb.run(pp)
            mp_, ms_ = cb.timings(3)
            _, _, cst = cb.download()
            side_ms[f"c4_{mode}_pyr"], side_ms[f"c4_{mode}_sol"] = mp_, ms_
            c4_iters[mode] = [int(v) for v in cst.gn_iters][:c4_L]
        cb.close()

    # ---------------- max over ranks of every timed region ----------------
    keys = list(side_ms)
    red = reduce_max([ms_total_local, e2e_ms_local, tri_ms_local, sustained[1] if sustained else 0.0, e2e_sync_ms_local] +
                     [side_ms[k] for k in keys])
    ms_total, e2e_ms, tri_ms, ms_sus, e2e_sync_ms = red[:5]
    side_ms = dict(zip(keys, red[5:]))
    per_rank_resident = gather_all(ms_total_local / args.steps)
    per_rank_e2e = gather_all(e2e_ms_local / args.steps)
    per_rank_h2d = gather_all(h2d_gbs_local)

    # final gather (outside the timed regions): the only exchange of the sharded path
    if world > 1:
        kp_full, su_full = sharding.gather_results(torch.from_numpy(e2e_kp).cuda(), torch.from_numpy(e2e_succ).cuda(),
                                                   B * world)
        assert kp_full.shape[0] == B * world and su_full.shape[0] == B * world
        lo, hi = sharding.shard_range(B * world, rank, world)     # this rank's block comes back byte for byte
        assert torch.equal(kp_full[lo:hi].cpu(), torch.from_numpy(e2e_kp)) and torch.equal(su_full[lo:hi].cpu(), torch.from_numpy(e2e_succ))

    if rank != 0:
        barrier()   # rank 0 runs its single-call side records and the CPU arm now
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    props = torch.cuda.get_device_properties(local)
    sm_count = props.multi_processor_count
    sm_max_mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    fp32_probe = load_profile_json("r02_fp32_peak.json")
    if fp32_probe.get("fp32_unfused_flop_per_cycle_per_sm"):
        fp32_peak_tflops = fp32_probe["fp32_unfused_flop_per_cycle_per_sm"] * sm_count * sm_max_mhz * 1e6 / 1e12
        peak_source = (f"measured: {fp32_probe['fp32_unfused_flop_per_cycle_per_sm']:.1f} un-fused fp32 flop/cycle/SM "
                       f"(profiles/r02_fp32_peak.json, tools/probe/pipe_probe.cu) x {sm_count} SMs x {sm_max_mhz:.0f} MHz")
    else:
        fp32_peak_tflops = sm_count * 128 * sm_max_mhz * 1e6 / 1e12
        peak_source = (f"computed (no probe result committed): {sm_count} SMs x 128 lanes x {sm_max_mhz:.0f} MHz un-fused fp32")
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    prof = load_profile_json("r02_profile.json")     # ncu: instructions executed / DRAM bytes per launch at this config
    default_cfg = (args.pairs, args.features, args.kernel, bool(args.subpixel)) == (256, 2000, 0, False)

    def prof_of(kernel, key):
        v = (prof.get(kernel) or {}).get(key) if default_cfg else None
        return v if v else None

    def solver_roofline(iter_list, n_feat_levels, patch_w, inverse, ms):
        fl = flops_of(iter_list, n_feat_levels, patch_w, inverse)
        ach = fl / (ms * 1e-3) / 1e12
        return {"achieved": ach, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": ach / fp32_peak_tflops,
                "algorithmic_flop_per_launch": fl, "ms_per_launch": ms}

    def pyramid_roofline(n_images, bytes_per_image, ms):
        gbs = n_images * bytes_per_image / (ms * 1e-3) / 1e9
        return {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "ms_per_launch": ms,
                "algorithmic_bytes_per_launch": n_images * bytes_per_image}

    # ---------------- rank 0 only: single-call configurations, CPU arm, parity ----------------
    threads = os.cpu_count() or 1
    seq = c1 = chain = None
    if not args.no_side and world == 1:
        seq = sequence_mode(trk, n, args)
        c1 = single_call_record(trk, args)
        chain = frontend_chain_record(trk, imgs1, imgs2, B, cam_l, cam_r)
    cpu = None
    parity = {}
    # the headline batch's first pairs once more through the bit-exact EXACT kernel (the on-GPU checker)
    Bc = min(B, 16)
    chk = trk.batch(Bc, ROWS, COLS, n, levels=LEVELS)
    chk.upload(imgs1[:Bc], imgs2[:Bc], kp1[:Bc], kp2[:Bc])
    chk.run(klt.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI, kernel=klt.KERNEL_EXACT))
    ex_kp, ex_succ, ex_st = chk.download()
    chk.run(params)                       # (the same pairs alone through the headline kernel: its Gauss-Newton pass counts)
    _, _, ln_st = chk.download()
    chk.close()
    d = np.abs(resident_kp[:Bc].astype(np.float64) - ex_kp).max(axis=2)
    parity["vs_exact_kernel"] = {
        "gn_passes_per_level_exact_kernel": [int(v) for v in ex_st.gn_iters][:LEVELS],
        "gn_passes_per_level_this_kernel": [int(v) for v in ln_st.gn_iters][:LEVELS],
        "p99.9_abs_dpos_px": float(np.quantile(d, 0.999)),
        "checked_features": int(Bc * n), "checker": "EXACT kernel (bit-identical to the CPU oracle, tests/)",
        "flag_mismatches": int((resident_succ[:Bc] != ex_succ).sum()), "max_abs_dpos_px": float(d.max()),
        "n_over_1e-3_px": int((d > 1e-3).sum()),
        "bit_identical_fraction": float((resident_kp[:Bc].view(np.uint32) == ex_kp.view(np.uint32)).all(axis=2).mean()),
        "e2e_equals_resident_bytes": bool(np.array_equal(e2e_kp.view(np.uint32), resident_kp.view(np.uint32)) and
                                          np.array_equal(e2e_succ, resident_succ) and
                                          np.array_equal(pipe_kp.view(np.uint32), resident_kp.view(np.uint32)) and
                                          np.array_equal(pipe_succ, resident_succ))}
    if not args.no_cpu_baseline and world == 1:
        arm = CpuArm()
        v, n_pairs, dt = arm.throughput(imgs1, imgs2, kp1, kp2, args.cpu_budget, threads)
        cpu = {"value": v, "unit": "tracks/s", "cores": threads, "kind": arm.kind,
               "sample": f"{n_pairs} pairs x {n} features of the same workload ({dt:.1f} s wall), full pyramids + "
                         f"{LEVELS} levels, one pair per thread; {arm.what}"}
        # the WHOLE headline batch against the CPU arm (the reference's own code when kind == "reference")
        t0 = time.perf_counter()
        ref_kp, ref_ok = arm.track_pairs(imgs1, imgs2, kp1, kp2, threads)
        dref = np.abs(resident_kp.astype(np.float64) - ref_kp).max(axis=2)
        border = np.minimum.reduce([np.abs(ref_kp[..., 0]), np.abs(ref_kp[..., 1]), np.abs(ref_kp[..., 0] - COLS),
                                    np.abs(ref_kp[..., 1] - ROWS)])
        mism = resident_succ.astype(bool) != ref_ok.astype(bool)
        parity["vs_cpu_arm_full_batch"] = {
            "checker": arm.kind, "features": int(B * n), "seconds": time.perf_counter() - t0,
            "flag_mismatches": int(mism.sum()), "flag_mismatches_within_1e-4_px_of_a_border": int((mism & (border < 1e-4)).sum()),
            "max_abs_dpos_px": float(dref.max()), "p99.9_abs_dpos_px": float(np.quantile(dref, 0.999)),
            "n_over_1e-3_px": int((dref > 1e-3).sum()),
            "bit_identical_fraction": float((resident_kp.view(np.uint32) == ref_kp.view(np.uint32)).all(axis=2).mean())}
        if not args.no_side:
            ref_s, ok_s = arm.track_pairs(imgs1[:32], imgs2[:32], kp1s[:32], kp1s[:32], threads)
            ds = np.abs(sub_kp[:32].astype(np.float64) - ref_s).max(axis=2)
            parity["subpixel_vs_cpu_arm_32_pairs"] = {
                "flag_mismatches": int((sub_succ[:32].astype(bool) != ok_s.astype(bool)).sum()),
                "max_abs_dpos_px": float(ds.max()), "n_over_1e-3_px": int((ds > 1e-3).sum()),
                "bit_identical_fraction": float((sub_kp[:32].view(np.uint32) == ref_s.view(np.uint32)).all(axis=2).mean())}
    barrier()

    # ---------------- the line ----------------
    value = world * n_tracks * args.steps / (ms_total * 1e-3)
    e2e_value = world * n_tracks * args.steps / (e2e_ms * 1e-3)
    h2d_floor = min(per_rank_h2d)
    e2e_gbs_per_rank = h2d_bytes / (e2e_ms / args.steps * 1e-3) / 1e9
    lane_path = args.kernel in (0, 3)
    sol = solver_roofline(iters, n_tracks * LEVELS, PATCH_HI - PATCH_LO + 1, False, ms_sol)
    inst = prof_of("klt_lane_kernel", "inst_executed")
    executed = None
    if inst:
        issue_peak = sm_count * 4 * sm_max_mhz * 1e6     # warp instructions per second: one per sub-partition per cycle
        inst_all = sum((prof.get(k) or {}).get("inst_executed", 0) for k in ("klt_lane_kernel", "klt_template_kernel"))
        executed = {"warp_instructions_per_launch": int(inst_all),
                    "issue_rate_frac": inst_all / (ms_sol * 1e-3) / issue_peak,
                    "warp_instructions_per_32_gn_passes": inst_all / max(1, sum(iters)) * 32,
                    "what": "smsp__inst_executed.sum of klt_lane_kernel + klt_template_kernel from the committed ncu capture "
                            "(profiles/r02_profile.json) over the live solver time: the executed-instruction view beside the "
                            "reference-formulation flop count (SURVEY.md 8d)"}
    line = {
        "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": "tracks/s",
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(kp2_io.nbytes + succ.nbytes + 12 * 8),
                "ms_per_step": e2e_ms / args.steps,
                "api": "lego_klt_track_batched_begin / _end, two calls in flight on two batch objects (pinned host buffers)",
                "one_call_in_flight": {"value": world * n_tracks * args.steps / (e2e_sync_ms * 1e-3), "ms_per_step": e2e_sync_ms / args.steps,
                                       "api": "lego_klt_track_batched, synchronous"},
                "h2d_ceiling_gbs": h2d_floor, "h2d_gbs_in_e2e": e2e_gbs_per_rank, "frac_of_ceiling": e2e_gbs_per_rank / h2d_floor,
                "h2d_ceiling_what": "the same pinned buffers uploaded without tracking (lego_klt_batch_upload alternating on the two "
                                    "batch objects), all ranks at once; slowest rank"},
        "gpu_launches": int(gpu_launches),
        "roofline": dict(sol, **{
            "kernel": "klt_template_kernel + klt_lane_kernel (+ klt_warp_kernel on deferred features): fused 4-level GN solver"
                      if lane_path else "solver kernel %d" % args.kernel,
            "bound": "fp32-issue (non-tensor)", "traffic": prof_of("klt_lane_kernel", "dram_bytes"),
            "traffic_template_kernel": prof_of("klt_template_kernel", "dram_bytes"), "peak_source": peak_source,
            "gn_iters_per_level": iters, "share_of_step": ms_sol / (ms_sol + ms_pyr), "executed": executed}),
        "roofline_pyramid": dict(pyramid_roofline(2 * B, PYR_BYTES_PER_IMAGE, ms_pyr), **{
            "kernel": "pyramid_l01_kernel + pyramid_x2x2_kernel (all levels + row aprons; the band kernel on other shapes)", "bound": "hbm",
            "traffic": ((prof_of("pyramid_l01_kernel", "dram_bytes") or 0) + (prof_of("pyramid_x2x2_kernel", "dram_bytes") or prof_of("pyramid_band_kernel", "dram_bytes") or 0)) or None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}),
        "per_rank": {"resident_ms_per_step": per_rank_resident, "e2e_ms_per_step": per_rank_e2e, "h2d_ceiling_gbs": per_rank_h2d},
        "triangulation": {"what": "lego_klt_batch_triangulate on the tracked batch (keypoints in HBM, world points to "
                                  "pinned host memory): legoslam::triangulation, SURVEY.md 8f N3",
                          "points_per_s": world * n_tracks / (tri_ms * 1e-3), "ms_per_call": tri_ms,
                          "d2h_bytes_per_call": int(tri_pt.nbytes + tri_ok.nbytes), "n_accepted": int(tri_ok.sum())},
        "parity": parity,
        "accuracy_vs_truth": accuracy_vs_truth(resident_kp, resident_succ, truth),
        "host": {"numa_binding": numa, "e2e_wall_ms_per_step": e2e_wall_s * 1e3 / args.steps},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "solver": {"n_success": n_success, "n_slow_path_passes": slow, "n_deferred_features": deferred,
                   "stats[deferred_irregular, two_family_passes, of_which_split, masked_warp_trips]": defer_reason,
                   "kernel": {0: "auto (lane + warp for deferred)", 1: "exact", 2: "warp", 3: "lane"}[int(args.kernel)]},
    }
    if sustained:
        n_sus, _, sus_clocks = sustained
        line["sustained"] = {"steps": n_sus, "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sus,
                             "value": world * n_tracks * n_sus / (ms_sus * 1e-3), "unit": "tracks/s", "clocks": sus_clocks}
    if not args.no_side:
        line["sequence_mode"] = seq
        line["single_call"] = c1
        line["frontend_chain"] = chain
        line["other_patches"] = {
            name: {"patch": list(pb), "value": world * n_tracks / (side_ms[f"patch_{name}"] * 1e-3), "unit": "tracks/s",
                   "ms_per_step": side_ms[f"patch_{name}"]} for name, pb in (("8x8", (-4, 3)), ("11x11", (-5, 5)))}
        line["subpixel"] = {
            "what": "the headline batch with source keypoints jittered by +-0.5 px (tracked points as fed back by "
                    "TrackLastFrame), kp2 = kp1; one batch in flight",
            "value": world * n_tracks / (side_ms["subpixel_resident"] * 1e-3), "unit": "tracks/s",
            "ms_per_step": side_ms["subpixel_resident"],
            "e2e": {"value": world * n_tracks / (side_ms["subpixel_e2e"] * 1e-3), "ms_per_step": side_ms["subpixel_e2e"]},
            "n_slow_path_passes": int(st_s.n_slow_path), "n_deferred_features": int(st_s.n_deferred),
            "two_family_passes": int(st_s.defer_reason[1]), "two_family_passes_that_split": int(st_s.defer_reason[2]),
            "masked_warp_trips": int(st_s.defer_reason[3]), "gn_iters_per_level": [int(v) for v in st_s.gn_iters][:LEVELS]}
        line["guess_projected"] = {
            "what": "the headline batch with the initial guess = ground truth + N(0, 2 px) (the callers' projected-map-point "
                    "branch, src/frontend_g2o.cpp:461-463) instead of kp2 = kp1; device-resident, one batch in flight",
            "value": world * n_tracks / (side_ms["guess_projected"] * 1e-3), "unit": "tracks/s",
            "ms_per_step": side_ms["guess_projected"], "gn_iters_per_level": [int(v) for v in st_g.gn_iters][:LEVELS],
            "accuracy_vs_truth": guess_acc}
        sweep = []
        for cnt in sweep_counts:
            for pname, pw in (("7x7", 7), ("11x11", 11)):
                msol, mpyr = side_ms[f"sweep_{cnt}_{pname}_sol"], side_ms[f"sweep_{cnt}_{pname}_pyr"]
                sweep.append({"features_per_pair": cnt, "patch": pname, "pairs_per_gpu": sweep_B,
                              "value": world * sweep_B * cnt / ((msol + mpyr) * 1e-3), "unit": "tracks/s",
                              "ms_solver": msol, "ms_pyramid": mpyr,
                              "roofline_solver_frac": solver_roofline(sweep_iters[(cnt, pname)], sweep_B * cnt * LEVELS, pw, False, msol)["frac"],
                              "roofline_pyramid_frac": pyramid_roofline(2 * sweep_B, PYR_BYTES_PER_IMAGE, mpyr)["frac"]})
        line["sweep_c5"] = {"what": "BASELINE config C5: 64 pairs 1241x376 per GPU, device-resident, one batch in flight; "
                                    "AUTO kernel selection (LANE above 3000 features per launch)", "points": sweep}
        c4 = {"what": f"BASELINE config C4: {c4_B} pairs {c4_cols}x{c4_rows} per GPU x {c4_n} features, {c4_L} levels, "
                      "device-resident; inverse = the reference's inverse mode with its stale Jacobian (SURVEY.md F4)"}
        c4_pyr_bytes = 2_762_040   # SURVEY.md 8d, 1920x1080, 5 levels
        for mode in ("forward", "inverse"):
            msol, mpyr = side_ms[f"c4_{mode}_sol"], side_ms[f"c4_{mode}_pyr"]
            c4[mode] = {"value": world * c4_B * c4_n / ((msol + mpyr) * 1e-3), "unit": "tracks/s", "ms_solver": msol,
                        "ms_pyramid": mpyr, "gn_iters_per_level": c4_iters[mode],
                        "roofline_solver_frac": solver_roofline(c4_iters[mode], c4_B * c4_n * c4_L, 7, mode == "inverse", msol)["frac"],
                        "roofline_pyramid_frac": pyramid_roofline(2 * c4_B, c4_pyr_bytes, mpyr)["frac"]}
        line["config_c4"] = c4
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def frontend_chain_record(trk, imgs1, imgs2, B, cam_l, cam_r, n_feat: int = 150):
    """The frontend's per-frame chain on a batch of frames whose images are in HBM, keypoints never crossing PCIe:
    DetectFeatures on the left images (cv::GFTTDetector::create(150, 0.01, 20), src/frontend_g2o.cpp:16,279-297) ->
    FindFeaturesInRight (:495-535, guess = the same pixel) -> triangulation of the matches (:310-349)."""
    import torch
    import lego_slam_b200 as klt
    fb = trk.batch(B, ROWS, COLS, n_feat, levels=LEVELS)
    z = klt.pinned_empty((B, n_feat, 2), np.float32)
    z[:] = 0
    fb.upload(imgs1, imgs2, z, z)
    p = klt.make_params(levels=LEVELS)
    tri_pt = klt.pinned_empty((B, n_feat, 3), np.float64)
    tri_ok = klt.pinned_empty((B, n_feat), np.uint8)

    def once():
        t0 = time.perf_counter()
        _, cnt, _ = fb.detect_features(0, n_feat, 0.01, 20.0)
        t1 = time.perf_counter()
        fb.use_detected_features()
        fb.run(p)
        trk.sync()
        t2 = time.perf_counter()
        fb.triangulate(cam_l, cam_r, 1e-3, tri_pt, tri_ok)
        t3 = time.perf_counter()
        return cnt, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3

    once()
    reps = [once() for _ in range(5)]
    cnt = reps[-1][0]
    det, trackms, tri = (float(np.median([r[i] for r in reps])) for i in (1, 2, 3))
    _, succ, st = fb.download()
    fb.close()
    total = det + trackms + tri
    return {"what": f"{B} frames 1241x376 in HBM: lego_klt_batch_detect_features ({n_feat} corners, 0.01, 20) -> "
                    "lego_klt_batch_use_detected_features -> lego_klt_batch_run (stereo matching) -> lego_klt_batch_triangulate; "
                    "host wall clock incl. the result copies of detection and triangulation",
            "ms_detect": det, "ms_track": trackms, "ms_triangulate": tri, "ms_total": total,
            "frames_per_s": B / (total * 1e-3), "corners_per_frame_mean": float(cnt.mean()),
            "detect_images_per_s": B / (det * 1e-3), "tracked": int(st.n_success), "triangulated_ok": int(tri_ok.sum())}


def single_call_record(trk, args):
    """BASELINE config C1: ONE stereo pair 1241x376, 150 features, one lego_klt_track call (host buffers in and out:
    2 image uploads, 2 pyramids, the solver, results back), and the 1080p / 5000-feature call of C4."""
    from lego_slam_b200 import synth
    import lego_slam_b200 as klt
    out = {}
    for name, (rows, cols, nf, lv, seed) in (("c1_150_features", (ROWS, COLS, 150, 4, 1)),
                                             ("c4_5000_features_1080p", (1080, 1920, 5000, 5, 3))):
        L, R, kp1, kp2, _ = synth.stereo_case(rows, cols, nf, seed=seed)
        rec = {"workload": f"one pair {cols}x{rows}, {nf} features, {lv} levels, lego_klt_track (host in, host out)"}
        for mode in ("forward", "inverse"):
            p = klt.make_params(levels=lv, inverse=(mode == "inverse"))
            for _ in range(5):
                trk.track(L, R, kp1, kp2, p)
            reps = 40
            t0 = time.perf_counter()
            for _ in range(reps):
                trk.track(L, R, kp1, kp2, p)
            dt = (time.perf_counter() - t0) / reps
            rec[mode] = {"ms_per_call": dt * 1e3, "tracks_per_s": nf / dt}
        if not args.no_cpu_baseline:
            arm = CpuArm(levels=lv)
            t0 = time.perf_counter()
            reps = 3
            for _ in range(reps):
                arm.track_pairs(L[None], R[None], kp1[None], kp2[None], 1)
            dt = (time.perf_counter() - t0) / reps
            rec["cpu_baseline_forward"] = {"ms_per_call": dt * 1e3, "tracks_per_s": nf / dt, "cores": 1, "kind": arm.kind,
                                           "sample": f"{reps} calls; the reference's parallel_for_ run as one stripe"}
        out[name] = rec
    # the step before tracking on the device as well: Frontend::DetectFeatures (src/frontend_g2o.cpp:279-297)
    L, R, kp1, kp2, _ = synth.stereo_case(ROWS, COLS, 150, seed=1)
    h = trk.image(ROWS, COLS, LEVELS).upload(L)
    det = {"workload": f"cv::goodFeaturesToTrack(150, 0.01, 20) on one {COLS}x{ROWS} image, 40 existing features masked out"}
    existing = kp1[:40]
    for name, fn in (("host_image", lambda: trk.detect_features(L, 150, 0.01, 20.0, exclude=existing)),
                     ("image_handle", lambda: trk.detect_features(h, 150, 0.01, 20.0, exclude=existing))):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(20):
            pts, _ = fn()
        det[name] = {"ms_per_call": (time.perf_counter() - t0) / 20 * 1e3, "corners": int(len(pts))}
    try:
        import cv2
        from oracle import gftt_np
        mask = gftt_np.exclusion_mask(ROWS, COLS, existing, 10.0)
        t0 = time.perf_counter()
        for _ in range(5):
            cv2.goodFeaturesToTrack(L, 150, 0.01, 20.0, mask=mask)
        det["cpu_cv2"] = {"ms_per_call": (time.perf_counter() - t0) / 5 * 1e3, "threads": cv2.getNumThreads()}
    except Exception:  # noqa: BLE001
        pass
    out["detect_features"] = det
    return out


def sequence_mode(trk, n_feat, args):
    """BASELINE config C2 beside the headline: ONE camera, frame after frame -- per frame the temporal track
    (last left -> current left) and the stereo match (current left -> current right), 2000 features each, the way
    Frontend::Track calls them (src/frontend_g2o.cpp:453-535).  Latency-bound (two synchronous calls per frame), so
    tracks/s is far below the batched figure; timed by wall clock around the calls a user makes (host buffers in,
    host buffers out).  `handles`: image handles with cached pyramids (each frame uploads 2 images, builds 2 pyramids);
    `pairwise`: lego_klt_track as the reference's signature implies (4 uploads, 4 pyramids per frame)."""
    from lego_slam_b200 import synth
    import lego_slam_b200 as klt
    frames = []
    L, R, kps, _, _ = synth.stereo_case(ROWS, COLS, n_feat, seed=2)
    for f in (1, 2, 3):
        P, Cur, kt, _, _ = synth.temporal_case(ROWS, COLS, n_feat, seed=2, frame=f)
        frames.append((P, Cur, kt))
    params = klt.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI)
    prev_h, cur_h, right_h = (trk.image(ROWS, COLS, LEVELS) for _ in range(3))
    n_frames = 60

    def run_handles(count):
        nonlocal prev_h, cur_h
        for i in range(count):
            P, Cur, kt = frames[i % len(frames)]
            if i == 0:
                prev_h.upload(P)
            cur_h.upload(Cur)
            right_h.upload(R)
            # (no counters requested, like the C++ shim: the reference's signature has none)
            trk.track_images(prev_h, cur_h, kt, kt, params, want_stats=False)     # Frontend::TrackLastFrame...4LayerSelf
            trk.track_images(cur_h, right_h, kps, kps, params, want_stats=False)  # Frontend::FindFeaturesInRight...4LayerSelf
            prev_h, cur_h = cur_h, prev_h

    def run_fused(count):
        # the whole frame in ONE call: temporal track + device-chained stereo match of the kept features
        nonlocal prev_h, cur_h
        for i in range(count):
            P, Cur, kt = frames[i % len(frames)]
            if i == 0:
                prev_h.upload(P)
            cur_h.upload(Cur)
            right_h.upload(R)
            trk.track_frame(prev_h, cur_h, right_h, kt, kt, params)
            prev_h, cur_h = cur_h, prev_h

    def run_pairwise(count):
        for i in range(count):
            P, Cur, kt = frames[i % len(frames)]
            trk.track(P, Cur, kt, kt, params)
            trk.track(L, R, kps, kps, params)

    res = {"workload": f"C2: sequence, per frame temporal + stereo track of {n_feat} features, {COLS}x{ROWS}, {LEVELS} levels"}
    for name, fn in (("handles", run_handles), ("fused_frame_call", run_fused), ("pairwise", run_pairwise)):
        fn(5)
        trk.sync()
        t0 = time.perf_counter()
        fn(n_frames)
        trk.sync()
        dt = time.perf_counter() - t0
        res[name] = {"ms_per_frame": dt / n_frames * 1e3, "frames_per_s": n_frames / dt,
                     "tracks_per_s": 2 * n_feat * n_frames / dt}
    if not args.no_cpu_baseline:
        arm = CpuArm()
        P, Cur, kt = frames[0]
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            arm.track_pairs(P[None], Cur[None], kt[None], kt[None], 1)
            arm.track_pairs(L[None], R[None], kps[None], kps[None], 1)
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"ms_per_frame": dt * 1e3, "tracks_per_s": 2 * n_feat / dt, "cores": 1, "kind": arm.kind,
                               "sample": f"{reps} frames, each call one stripe on one core (the reference splits the features "
                                         "of a call over its cv::parallel_for_ threads)"}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=256, help="stereo pairs per GPU")
    ap.add_argument("--features", type=int, default=2000)
    ap.add_argument("--distinct", type=int, default=64, help="distinct generated pairs (tiled to --pairs)")
    ap.add_argument("--kernel", type=int, default=0, help="LEGO_KLT_KERNEL_* (0 = auto)")
    ap.add_argument("--streams", type=int, default=2,
                    help="device-resident batches in flight (value only; 2 = double buffering: the pyramid / template "
                         "kernels of one batch fill the tail of the other batch's persistent solver kernel)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-side", "--no-sequence", dest="no_side", action="store_true",
                    help="skip the side records (C1, C2, C4, C5 sweep, other patches, sub-pixel keypoints)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s sustained loop")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--subpixel", action="store_true",
                    help="jitter the source keypoints of the HEADLINE by +-0.5 px (the side record does it anyway)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
