#!/usr/bin/env python
"""bench.py -- KLT feature-tracks/s of the B200-native pyramid Gauss-Newton KLT path.

Metric (BASELINE.json): KLT feature-tracks/sec, 4-level pyramid, at 1/2/4/8 B200, with the CPU
ParallelLoopBody-style path timed beside it, and pyramid GB/s.

Workload (config C3, the batched multi-GPU one the metric is quoted on): per GPU, B=256 independent
synthetic KITTI-shaped stereo pairs 1241x376 x 2000 features, 4 levels, the reference's 7x7 patch
(src/algorithm.cpp:40: half_patch_size=3 -- SURVEY.md F1; "8x8" in the metric text is not what the
reference computes), forward mode, initial guess kp2 = kp1.  A step = pyramids of all 512 images +
the fused 4-level solver over 512,000 features.  Weak scaling: every rank owns its own 256 pairs, no
data-path collective, final gather of (x,y,flag) outside the timed region.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS, LEVELS = 376, 1241, 4
PATCH_LO, PATCH_HI = -3, 3
FLOP_PER_PIXEL_ITER = 123          # SURVEY.md 8d, reference formulation, forward mode
PYR_BYTES_PER_IMAGE = 619_601      # level 0 read once + levels 1..3 written once
METRIC = "KLT feature-tracks/sec (4-level pyramid, 1241x376, 2000 features/pair, batched pairs)"


def _make_pair(args):
    seed, n = args
    from lego_slam_b200 import synth
    L, R, kp1, kp2, _ = synth.stereo_case(ROWS, COLS, n, seed=seed)
    return L, R, kp1, kp2


def make_workload(n_pairs: int, n_feat: int, distinct: int, seed0: int):
    """`distinct` generated pairs (seeds seed0..), tiled cyclically to n_pairs; odd repeats are
    vertically flipped so that repeated slots are not byte-identical."""
    distinct = max(1, min(distinct, n_pairs))
    jobs = [(seed0 + i, n_feat) for i in range(distinct)]
    workers = min(len(jobs), max(1, (os.cpu_count() or 2) // 2), 32)
    if workers > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            base = pool.map(_make_pair, jobs)
    else:
        base = [_make_pair(j) for j in jobs]
    return base


def fill_batch(base, n_pairs, n_feat, alloc):
    imgs1 = alloc((n_pairs, ROWS, COLS), np.uint8)
    imgs2 = alloc((n_pairs, ROWS, COLS), np.uint8)
    kp1 = alloc((n_pairs, n_feat, 2), np.float32)
    kp2 = alloc((n_pairs, n_feat, 2), np.float32)
    for b in range(n_pairs):
        L, R, a, g = base[b % len(base)]
        if (b // len(base)) % 2 == 1:  # vertical flip: still a valid stereo pair
            imgs1[b], imgs2[b] = L[::-1], R[::-1]
            f = a.copy()
            f[:, 1] = (ROWS - 1) - f[:, 1]
            kp1[b] = f
            kp2[b] = f
        else:
            imgs1[b], imgs2[b], kp1[b], kp2[b] = L, R, a, g
    return imgs1, imgs2, kp1, kp2


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (the same counters
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
            return
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self._stop.set()
        self.t.join(timeout=2)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.samples), "reasons": reasons}


def cpu_oracle_throughput(base, n_feat, budget_s: float, threads: int):
    """The reference's CPU path (oracle port: pyramids + 4 levels, src/algorithm.cpp:128-206) on the host
    cores: pairs run concurrently on `threads` worker threads (ctypes releases the GIL), one
    single-threaded LKOpticalFlow4Layer-equivalent per pair."""
    from oracle import binding as ob
    p = ob.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI)
    t0 = time.perf_counter()
    ob.track(base[0][0], base[0][1], base[0][2], base[0][3], p, threads=1)
    t1 = time.perf_counter() - t0
    n_pairs = int(max(threads, min(4096, budget_s * threads / max(t1, 1e-4))))
    n_pairs = (n_pairs + threads - 1) // threads * threads

    def one(i):
        L, R, a, g = base[i % len(base)]
        ob.track(L, R, a, g, p, threads=1)

    with ThreadPoolExecutor(threads) as ex:
        t0 = time.perf_counter()
        list(ex.map(one, range(n_pairs)))
        dt = time.perf_counter() - t0
    return n_pairs * n_feat / dt, n_pairs, dt


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  Its translation unit
    cannot be built offline (needs OpenCV/Eigen/Sophus/glog), so this is the oracle PORT, compiled
    with the reference's flags, on all host threads; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    base = make_workload(args.pairs, args.features, min(args.distinct, 16), 1000)
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    vals, sample = [], None
    for s in range(args.warmup + args.steps):
        v, n_pairs, dt = cpu_oracle_throughput(base, args.features, per_step_budget, threads)
        if s >= args.warmup:
            vals.append((v, dt))
        sample = f"{n_pairs} pairs x {args.features} features per step ({dt:.1f} s), full pyramids + 4 levels"
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 sampling + f64 normal equations",
        "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "tracks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tracks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(cuda_index: int):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, BEFORE the pinned host buffers are
    allocated (first touch puts them on that node).  With 8 ranks the end-to-end path is bound by host memory /
    PCIe root complexes; remote pinned buffers halve it.  Returns a short description (or the reason it did not)."""
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


_TRAFFIC = None


def traffic(kernel: str, args):
    """DRAM bytes per launch of `kernel` (ncu dram__bytes_read.sum + dram__bytes_write.sum) at the default
    configuration, from profiles/r01_traffic.json; None for any other configuration or if never captured."""
    global _TRAFFIC
    if (args.pairs, args.features, args.kernel, bool(args.subpixel)) != (256, 2000, 0, False):
        return None
    if _TRAFFIC is None:
        try:
            with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
                _TRAFFIC = json.load(f)
        except Exception:
            _TRAFFIC = {}
    v = _TRAFFIC.get(kernel)
    return int(v) if v else None


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

    def run_pairwise(count):
        for i in range(count):
            P, Cur, kt = frames[i % len(frames)]
            trk.track(P, Cur, kt, kt, params)
            trk.track(L, R, kps, kps, params)

    res = {"workload": f"C2: sequence, per frame temporal + stereo track of {n_feat} features, {COLS}x{ROWS}, {LEVELS} levels"}
    for name, fn in (("handles", run_handles), ("pairwise", run_pairwise)):
        fn(5)
        trk.sync()
        t0 = time.perf_counter()
        fn(n_frames)
        trk.sync()
        dt = time.perf_counter() - t0
        res[name] = {"ms_per_frame": dt / n_frames * 1e3, "frames_per_s": n_frames / dt,
                     "tracks_per_s": 2 * n_feat * n_frames / dt}
    if not args.no_cpu_baseline:
        from oracle import binding as ob   # the checker, timed as the CPU baseline of this configuration
        threads = os.cpu_count() or 1
        P, Cur, kt = frames[0]
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ob.track(P, Cur, kt, kt, threads=threads)
            ob.track(L, R, kps, kps, threads=threads)
        dt = (time.perf_counter() - t0) / reps
        res["cpu_baseline"] = {"ms_per_frame": dt * 1e3, "tracks_per_s": 2 * n_feat / dt, "cores": threads, "kind": "port",
                               "sample": f"{reps} frames, features split over {threads} threads like cv::parallel_for_"}
    return res


def workload_config(args):
    return {"workload": f"C3: {args.pairs} independent stereo pairs {COLS}x{ROWS} u8 per GPU x {args.features} "
                        f"features, {LEVELS}-level pyramid, 7x7 patch (reference half_patch_size=3), forward, kp2=kp1",
            "pairs_per_gpu": args.pairs, "features_per_pair": args.features, "levels": LEVELS,
            "patch": [PATCH_LO, PATCH_HI], "distinct_pairs": min(args.distinct, args.pairs),
            "l2_policy": "inputs larger than L2 (level-0 images of one step: %.0f MB > 126 MB)"
                         % (2 * args.pairs * ROWS * COLS / 1e6),
            "sharding": "block partition of pairs, one process per GPU, no data-path collective",
            "batches_in_flight": max(1, args.streams), "subpixel_keypoints": bool(args.subpixel)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import lego_slam_b200 as klt
    from lego_slam_b200 import build, sharding
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

    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    B, n = args.pairs, args.features
    base = make_workload(B, n, args.distinct, 1000 + rank * B)
    imgs1, imgs2, kp1, kp2 = fill_batch(base, B, n, klt.pinned_empty)
    if args.subpixel:
        kp1 += np.random.default_rng(77 + rank).uniform(-0.5, 0.5, kp1.shape).astype(np.float32)
        np.copyto(kp2, kp1)
    kp2_io = klt.pinned_empty((B, n, 2), np.float32)
    succ = klt.pinned_empty((B, n), np.uint8)
    params = klt.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI, kernel=args.kernel)
    n_tracks = B * n
    # `--streams S` batches in flight: S device-resident batch objects (same inputs), each on its own
    # stream, stepped round-robin -- the small kernels of one batch (pyramid, aprons, templates) fill
    # the issue slots the persistent solver kernel of the other leaves idle.
    S = max(1, args.streams)
    trks = [klt.Tracker(local) for _ in range(S)]
    streams = [torch.cuda.Stream(device=local) for _ in range(S)]
    batches = []
    for t, st_ in zip(trks, streams):
        t.set_stream(st_.cuda_stream)
        batches.append(t.batch(B, ROWS, COLS, n, levels=LEVELS))
    trk, stream, batch = trks[0], streams[0], batches[0]

    # ---------------- device-resident: inputs already in HBM, results stay in HBM ----------------
    for b_ in batches:
        b_.upload(imgs1, imgs2, kp1, kp2)
    for i in range(args.warmup * S):
        batches[i % S].run(params)
    for t in trks:
        t.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(ev0)
    for i in range(args.steps):
        batches[i % S].run(params)
    for st_ in streams[1:]:
        streams[0].wait_stream(st_)
    ev1.record(streams[0])
    torch.cuda.synchronize()
    barrier()
    ms_total = ev0.elapsed_time(ev1)

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

    # ---------------- end to end: host buffers -> lego_klt_track_batched -> host buffers ----------------
    # kp2 is in/out (initial guess in, tracked position out): every timed step gets its own pre-filled pinned
    # buffer, so that no host-side refill of the guess sits inside the timed region.
    n_ring = args.steps if args.steps <= 32 else 1
    kp2_ring = [kp2_io] + [klt.pinned_empty((B, n, 2), np.float32) for _ in range(n_ring - 1)]
    for _ in range(min(args.warmup, 3)):
        np.copyto(kp2_io, kp2)
        batch.track(imgs1, imgs2, kp1, kp2_io, succ, params)
    for buf in kp2_ring:
        np.copyto(buf, kp2)
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        if n_ring == 1:
            np.copyto(kp2_io, kp2)
        batch.track(imgs1, imgs2, kp1, kp2_ring[i % n_ring], succ, params)   # synchronous: returns after the D2H
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_wall_s = time.perf_counter() - t0
    e2e_s = e0.elapsed_time(e1) * 1e-3     # device clock on the library's stream; the call blocks, so wall == device
    barrier()
    # (one NVML query takes several ms: the sampler spans both timed regions -- device-resident and end-to-end -- and
    # the per-kernel timing pass between them)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-resident steps, per-kernel timing pass and end-to-end steps"

    # ---------------- next step of the frontend on the tracked batch: triangulation ----------------
    left34 = np.hstack([np.eye(3), np.zeros((3, 1))])
    right34 = np.hstack([np.eye(3), np.array([[-0.537], [0.0], [0.0]])])
    cam_l = klt.make_camera(718.856, 718.856, 607.1928, 185.2157, left34)
    cam_r = klt.make_camera(718.856, 718.856, 607.1928, 185.2157, right34)
    tri_pt = klt.pinned_empty((B, n, 3), np.float64)
    tri_ok = klt.pinned_empty((B, n), np.uint8)
    batch.triangulate(cam_l, cam_r, 1e-3, tri_pt, tri_ok)
    t0 = time.perf_counter()
    for _ in range(3):
        batch.triangulate(cam_l, cam_r, 1e-3, tri_pt, tri_ok)
    tri_ms = (time.perf_counter() - t0) / 3 * 1e3

    # max over ranks of the timed regions
    t = torch.tensor([ms_total, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    # final gather (outside the timed regions): the only exchange of the sharded path
    if world > 1:
        kp_full, su_full = sharding.gather_results(torch.from_numpy(kp2_io).cuda(), torch.from_numpy(succ).cuda(),
                                                   B * world)
        assert kp_full.shape[0] == B * world and su_full.shape[0] == B * world

    if rank != 0:
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
    fp32_peak_tflops = sm_count * 128 * sm_max_mhz * 1e6 / 1e12   # un-fused fp32 lane-ops/s
    P = (PATCH_HI - PATCH_LO + 1) ** 2
    algo_flop = sum(iters) * P * FLOP_PER_PIXEL_ITER
    achieved_tflops = algo_flop / (ms_sol * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    pyr_bytes = 2 * B * PYR_BYTES_PER_IMAGE
    pyr_gbs = pyr_bytes / (ms_pyr * 1e-3) / 1e9

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, n_pairs, dt = cpu_oracle_throughput(base, n, args.cpu_budget, threads)
        cpu = {"value": v, "unit": "tracks/s", "cores": threads, "kind": "port",
               "sample": f"{n_pairs} pairs x {n} features of the same workload ({dt:.1f} s wall), full pyramids + "
                         f"{LEVELS} levels, oracle built -std=c++11 -O3 (reference flags), one pair per thread"}

    seq = sequence_mode(trk, n, args) if world == 1 and not args.no_sequence else None

    # ---------------- parity report of this run, outside the timed regions ----------------
    # The first pairs of the batch once more through the bit-exact EXACT kernel (thread per feature, reference
    # operation order: the on-GPU checker, itself pinned to the CPU oracle by the tests), and -- at N = 1 -- pair 0
    # through the CPU oracle (the checker; never on the product path).
    parity = None
    if rank == 0:
        Bc = min(B, 16)
        chk = trk.batch(Bc, ROWS, COLS, n, levels=LEVELS)
        chk.upload(imgs1[:Bc], imgs2[:Bc], kp1[:Bc], kp2[:Bc])
        res = {}
        for name, k in (("exact", klt.KERNEL_EXACT), ("fast", args.kernel)):
            chk.run(klt.make_params(levels=LEVELS, patch_lo=PATCH_LO, patch_hi=PATCH_HI, kernel=k))
            o, s_, st_ = chk.download()
            res[name] = (o.copy(), s_.copy(), [int(v) for v in st_.gn_iters][:LEVELS])
        d = np.abs(res["fast"][0].astype(np.float64) - res["exact"][0]).max(axis=2)
        parity = {"checked_features": int(Bc * n), "checker": "EXACT kernel (bit-identical to the CPU oracle)",
                  "flag_mismatches": int((res["fast"][1] != res["exact"][1]).sum()),
                  "max_abs_dpos_px": float(d.max()), "n_over_1e-3_px": int((d > 1e-3).sum()),
                  "bit_identical_fraction": float((res["fast"][0].view(np.uint32) == res["exact"][0].view(np.uint32)).all(axis=2).mean()),
                  "gn_iters_equal": res["fast"][2] == res["exact"][2]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import binding as ob
            ro, rs, rst = ob.track(np.ascontiguousarray(imgs1[0]), np.ascontiguousarray(imgs2[0]), kp1[0], kp2[0],
                                   threads=os.cpu_count() or 1)
            d0 = np.abs(res["fast"][0][0].astype(np.float64) - ro).max(axis=1)
            parity["vs_cpu_oracle_pair0"] = {"features": int(n), "flag_mismatches": int((res["fast"][1][0] != rs).sum()),
                                             "max_abs_dpos_px": float(d0.max()),
                                             "exact_kernel_bit_identical": bool(np.array_equal(
                                                 res["exact"][0][0].view(np.uint32), ro.view(np.uint32)))}

    # The metric text of BASELINE.json says "8x8 patch"; the reference computes 7x7 (src/algorithm.cpp:40,63-64,
    # SURVEY.md F1), which is what `value` measures.  The literal 8x8 patch (-4..3) and the 11x11 patch of the stress
    # configuration (-5..5) beside it, same batch, device-resident (one batch in flight): the LANE solver is compiled
    # for all three patches (klt_solver_lane.cu, _p8.cu, _p11.cu).
    other_patches = None
    if world == 1 and not args.no_sequence:
        other_patches = {}
        for name, (plo, phi) in (("8x8", (-4, 3)), ("11x11", (-5, 5))):
            pp = klt.make_params(levels=LEVELS, patch_lo=plo, patch_hi=phi)
            batch.run(pp)
            trk.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                batch.run(pp)
            e1.record(stream)
            torch.cuda.synchronize()
            msp = e0.elapsed_time(e1) / 3
            other_patches[name] = {"patch": [plo, phi], "value": n_tracks / (msp * 1e-3), "unit": "tracks/s",
                                   "ms_per_step": msp}

    value = world * n_tracks * args.steps / (ms_total * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "tracks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 sampling + f64 normal equations", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": world * n_tracks * args.steps / (e2e_ms * 1e-3), "unit": "tracks/s",
                "h2d_bytes_per_step": int(imgs1.nbytes + imgs2.nbytes + kp1.nbytes + kp2.nbytes),
                "d2h_bytes_per_step": int(kp2_io.nbytes + succ.nbytes + 12 * 8),
                "ms_per_step": e2e_ms / args.steps, "api": "lego_klt_track_batched (pinned host buffers)"},
        # kernels of this library launched inside the device-resident timed region, per step: pyramid_l01_kernel,
        # pyramid_band_kernel, then klt_template_kernel + klt_warp_kernel (deferred features) + klt_lane_kernel x2
        # (LANE path) or one solver kernel
        "gpu_launches": (6 if args.kernel in (0, 3) else 3) * args.steps,
        "roofline": {"kernel": "klt_template_kernel + klt_lane_kernel (+ klt_warp_kernel on deferred features): fused 4-level GN solver"
                     if args.kernel in (0, 3) else "klt_warp_kernel (fused 4-level GN solver)", "bound": "fp32-issue (non-tensor)",
                     "achieved": achieved_tflops, "peak": fp32_peak_tflops, "unit": "TFLOP/s",
                     "frac": achieved_tflops / fp32_peak_tflops,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch from one `ncu --set full` capture at this
                     # exact configuration (profiles/r01_traffic.json, written by tools/summarise_ncu.py); null otherwise
                     "traffic": traffic("klt_lane_kernel", args),
                     "traffic_template_kernel": traffic("klt_template_kernel", args),
                     "peak_source": f"computed: {sm_count} SMs x 128 lanes x {sm_max_mhz:.0f} MHz un-fused fp32 "
                                    "(not in MEASURED_PEAKS.json, which has only HBM and bf16 tensor peaks)",
                     "algorithmic_flop_per_launch": algo_flop, "gn_iters_per_level": iters,
                     "ms_per_launch": ms_sol, "share_of_step": ms_sol / (ms_sol + ms_pyr)},
        "roofline_pyramid": {"kernel": "pyramid_l01_kernel + pyramid_band_kernel (all levels + row aprons)", "bound": "hbm",
                             "achieved": pyr_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": pyr_gbs / hbm_peak,
                             "traffic": (traffic("pyramid_l01_kernel", args) or 0) + (traffic("pyramid_band_kernel", args) or 0) or None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                             "algorithmic_bytes_per_launch": pyr_bytes, "ms_per_launch": ms_pyr},
        "triangulation": {"what": "lego_klt_batch_triangulate on the tracked batch (keypoints in HBM, world points to "
                                  "pinned host memory): legoslam::triangulation, SURVEY.md 8f N3",
                          "points_per_s": world * n_tracks / (tri_ms * 1e-3), "ms_per_call": tri_ms,
                          "d2h_bytes_per_call": int(tri_pt.nbytes + tri_ok.nbytes), "n_accepted": int(tri_ok.sum())},
        "parity": parity,
        "sequence_mode": seq,
        "other_patches": other_patches,
        "host": {"numa_binding": numa, "e2e_wall_ms_per_step": e2e_wall_s * 1e3 / args.steps},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "solver": {"n_success": n_success, "n_slow_path_passes": slow, "n_deferred_features": deferred, "defer_reason[inexact,margin,nominal,range]": defer_reason,
                   "kernel": {0: "auto (lane + warp for deferred)", 1: "exact", 2: "warp", 3: "lane"}[int(args.kernel)]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--no-sequence", action="store_true", help="skip the C2 sequence-mode side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--subpixel", action="store_true",
                    help="jitter the source keypoints by +-0.5 px (tracked points as fed back by TrackLastFrame)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
