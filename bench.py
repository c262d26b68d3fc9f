#!/usr/bin/env python
"""Benchmark of the hot path: frame pairs -> Farneback flow -> flow-graph clustering -> 3D boxes.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores
    python bench.py --config {stream1080,single1080,uhd60}   # BASELINE configs [2]/[4] (default), [1], [3]

A step = one batch of B consecutive frame pairs (B+1 frames) of the synthetic video through dofs3d_process*.
`value` times the device-pointer entry point with the frames already in HBM (every step its own frames); `e2e` times
the streaming entry points with HOST buffers: pinned frames in, run-length labels + boxes out, the upload of a chunk
under the kernels of the previous one, and — with more than one rank — the NCCL gather of the last step's boxes, all
inside the timed region.  Frames are sharded contiguously over ranks; there is no collective on the data path (weak
scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "frame-pairs/s"
SEED = 1234
FULL_PAIR_CPU_SECONDS = 40.0  # rough cost of one 1080p pair through the reference on one core (33 s)

CONFIGS = {
    # BASELINE.json configs[2] / configs[4]: streamed synthetic 1080p video — the configuration the metric is quoted on
    "stream1080": dict(width=1920, height=1080, objects=8, batch=192, contexts=6,
                       metric="1080p frame-pairs/s (flow->segment->3D)",
                       what="BASELINE configs[2]/[4]: synthetic 1920x1080 video streamed in batches"),
    # configs[1]: ONE synthetic 1080p pair per call: the latency of the path (its serial merge-chain replay is exposed)
    "single1080": dict(width=1920, height=1080, objects=8, batch=1, contexts=1,
                       metric="1080p frame-pairs/s, one pair per call (latency)",
                       what="BASELINE configs[1]: a single synthetic 1920x1080 pair per call"),
    # configs[3]: 4K with 60 overlapping objects (union-find replay and per-cluster lifting under stress)
    "uhd60": dict(width=3840, height=2160, objects=60, batch=24, contexts=3,
                  metric="4K (60 objects) frame-pairs/s (flow->segment->3D)",
                  what="BASELINE configs[3]: synthetic 3840x2160 video with 60 overlapping moving objects"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="stream1080", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="frame pairs per step per GPU (default: the config's)")
    ap.add_argument("--contexts", type=int, default=None, help="contexts (streams) per GPU; each takes batch/contexts pairs of a step")
    ap.add_argument("--labels", default="rle", choices=["rle", "u16", "i32"], help="label format of the end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=150.0, help="seconds the reference arm may spend in its steps")
    ap.add_argument("--no-calibration", action="store_true", help="reference arm: skip the full-frame calibration run")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    a.width, a.height, a.objects = cfg["width"], cfg["height"], cfg["objects"]
    a.batch = a.batch or cfg["batch"]
    a.contexts = a.contexts or cfg["contexts"]
    a.metric = cfg["metric"]
    return a


def workload_config(args, world):
    return {
        "workload": f"{CONFIGS[args.config]['what']} ({args.objects} moving textured objects; {args.batch} frame pairs per GPU "
                    f"per step; pair i = frames i, i+1)",
        "name": args.config,
        "pairs_per_step_per_gpu": args.batch,
        "frame": [args.width, args.height],
        "sharding": f"contiguous blocks of frames over {world} rank(s), no collective on the data path",
        "cache": "inputs larger than L2 (one step reads %.0f MB of BGR frames and streams >10 GB of intermediates); every "
                 "timed step of the resident leg has its own frames" % ((args.batch + 1) * args.width * args.height * 3 / 1e6),
    }


# --------------------------------------------------------------------------------------------------
# the reference's CPU path (cv2 for the OpenCV calls, oracle/_ref = the unchanged reference sources)
# --------------------------------------------------------------------------------------------------
_worker_state = {}


def _cpu_worker_init(width, height, rows, seed_base, objects):
    """Each worker renders its own frame pair (untimed) and keeps the band of `rows` rows it will process."""
    import cv2
    from denseopticalflowsegmentation3d_b200 import synth
    from oracle import cpu
    cv2.setNumThreads(1)
    idx = int(os.environ.get("DOFS_WORKER_INDEX", "0"))
    wid = os.getpid()
    first = (wid * 7 + idx) % 16
    fr = synth.frames(seed_base, objects, first, 2, width, height)
    y0 = (height - rows) // 2
    _worker_state["frames"] = np.ascontiguousarray(fr[:, y0:y0 + rows])
    _worker_state["oracle"] = cpu.ref() if cpu.ref_available() else cpu.port()
    _worker_state["kind"] = "reference" if cpu.ref_available() else "port"
    _worker_state["mats"] = _worker_state["oracle"].get_mats()
    return _worker_state["kind"]


def _cpu_pair(_):
    """cvtColor x2 -> calcOpticalFlowFarneback -> GaussianBlur -> build_graph -> segment_graph -> get_best_segments
    (segment.cpp:97-101, 34-72) on the worker's frame pair; returns (seconds, number of segments)."""
    import cv2
    fr = _worker_state["frames"]
    o = _worker_state["oracle"]
    persp, inv, up = _worker_state["mats"]
    t0 = time.perf_counter()
    g0 = cv2.cvtColor(fr[0], cv2.COLOR_BGR2GRAY)
    g1 = cv2.cvtColor(fr[1], cv2.COLOR_BGR2GRAY)
    flow = cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    fb = cv2.GaussianBlur(flow, (0, 0), 3.0)
    res = o.segment(fb, persp, inv, up)
    return time.perf_counter() - t0, len(res["entries"])


def usable_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample_rows(height, seconds_per_step, full_seconds=FULL_PAIR_CPU_SECONDS):
    frac = min(1.0, max(seconds_per_step / full_seconds, 0.125))
    rows = int(round(height * frac / 8)) * 8
    return max(min(rows, height), 64)


def run_cpu_reference(args, cores, rows, steps, warmup):
    """`cores` worker processes, each pushing one frame pair (band of `rows` rows) per step through the reference."""
    if cores == 1:  # in-process (used for the cpu_baseline key of the GPU arm: no fork after CUDA start-up)
        kind = _cpu_worker_init(args.width, args.height, rows, SEED, args.objects)
        for _ in range(warmup):
            _cpu_pair(0)
        t0 = time.perf_counter()
        n_seg = sum(_cpu_pair(0)[1] for _ in range(steps))
        dt = time.perf_counter() - t0
        return {"kind": kind, "seconds": dt, "pairs_per_s": steps * (rows / args.height) / dt,
                "ms_per_step": 1e3 * dt / steps, "segments": n_seg}
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(args.width, args.height, rows, SEED, args.objects)) as pool:
        kind = pool.apply(_cpu_worker_kind)
        for _ in range(warmup):
            pool.map(_cpu_pair, range(cores), chunksize=1)
        t0 = time.perf_counter()
        n_seg = 0
        for _ in range(steps):
            out = pool.map(_cpu_pair, range(cores), chunksize=1)
            n_seg += sum(o[1] for o in out)
        dt = time.perf_counter() - t0
    pairs = cores * steps * (rows / args.height)  # a band of rows/H of the frame counts as that fraction of a pair
    return {"kind": kind, "seconds": dt, "pairs_per_s": pairs / dt, "ms_per_step": 1e3 * dt / steps, "segments": n_seg}


def _cpu_worker_kind():
    return _worker_state.get("kind", "port")


def calibrate_band(args, rows):
    """How much a band under-costs the reference: the same core runs a band of `rows` rows and ONE FULL frame pair (the
    reference's multiset sort and set copies are super-linear in the pixel count).  Returns the seconds of both."""
    code = ("import sys, json; sys.path.insert(0, %r); import bench, types; a = types.SimpleNamespace(width=%d, height=%d, "
            "objects=%d); b = bench.run_cpu_reference(a, 1, %d, 1, 0); f = bench.run_cpu_reference(a, 1, %d, 1, 0); "
            "print(json.dumps({'band_s': b['seconds'], 'full_s': f['seconds'], 'kind': f['kind']}))"
            % (ROOT, args.width, args.height, args.objects, rows, args.height))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        return None
    return json.loads(out.stdout.strip().splitlines()[-1])


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = usable_cores()
    try:
        avail_gb = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) / 1e6
        cores = max(1, min(cores, int(avail_gb * 0.5 / 2.0)))  # ~1.7 GB resident per 1080p reference process
    except Exception:
        pass
    scale = (args.width * args.height) / (1920 * 1080)
    full_guess = FULL_PAIR_CPU_SECONDS * scale * max(scale, 1.0) ** 0.2
    if args.config == "single1080":  # the latency configuration: ONE full pair on one core, nothing sampled
        r = run_cpu_reference(args, 1, args.height, 1, 0)
        cores, rows, cal, factor = 1, args.height, None, 1.0
        sample = "1 process, 1 full synthetic pair (no sampling): the reference's latency on one core"
    else:
        rows = cpu_sample_rows(args.height, args.cpu_budget / max(args.steps + args.warmup, 1), full_guess)
        cal = None if args.no_calibration or rows >= args.height else calibrate_band(args, rows)
        r = run_cpu_reference(args, cores, rows, args.steps, args.warmup)
        # value = full-frame pairs/s: the band throughput corrected by the measured band-vs-full cost ratio
        factor = 1.0
        if cal:
            factor = cal["full_s"] / (cal["band_s"] / (rows / args.height))
        sample = (f"{cores} processes x 1 synthetic pair per step, central band of {rows}/{args.height} rows at full width "
                  f"(a band counts as {rows / args.height:.3f} pair); cv2 flow+blur, then "
                  + ("the unchanged reference graph.cpp+lifting_3d.cpp (oracle/_ref)" if r["kind"] == "reference"
                     else "the oracle port (oracle/_ref not built)")
                  + (f"; calibrated on one core: a full pair costs {cal['full_s']:.1f} s, {factor:.2f}x the band's "
                     f"pro-rata {cal['band_s'] / (rows / args.height):.1f} s, so value = band throughput / {factor:.2f}"
                     if cal else "; uncalibrated (a band under-costs the super-linear reference)"))
    value = r["pairs_per_s"] / factor
    line = {
        "impl": "reference", "metric": args.metric, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 geometry / f64 weights", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample,
                         "band_pairs_per_s_uncalibrated": r["pairs_per_s"], "calibration": cal},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic():
    """The committed per-kernel DRAM traffic of one 32-pair 1080p call (ncu, tools/summarize_traffic.py)."""
    for name in ("r02_traffic.json", "r02_traffic_start.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                d["file"] = "profiles/" + name
                return d
            except Exception:
                pass
    return None


def ours(args):
    import ctypes as C
    import torch
    import denseopticalflowsegmentation3d_b200 as dofs
    from denseopticalflowsegmentation3d_b200 import capi
    from denseopticalflowsegmentation3d_b200.capi import BOX_DTYPE, RUN_DTYPE, STATS_DTYPE

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        # NCCL writes its version banner to stdout when the communicator is created: keep fd 1 clean for the JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        devnull = os.open(os.devnull, os.O_WRONLY)
        os.dup2(devnull, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            torch.cuda.set_device(local)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            os.close(devnull)
    torch.cuda.set_device(local)
    W, H, B, K, Wm, NC = args.width, args.height, args.batch, args.steps, args.warmup, args.contexts
    assert B % NC == 0, "--batch must be a multiple of --contexts"
    Bc = B // NC  # pairs per context per step
    N = W * H
    MAXB, MAXR = 256, 32 * H  # boxes / label runs per frame the outputs can hold
    dev = torch.device("cuda", local)
    params = None  # the reference's constants as they are (its 640x360 calibration, unscaled: SURVEY.md section 8d)
    # NC contexts = NC streams: the latency-bound phases of one sub-batch (merge-chain replay, late Boruvka levels)
    # overlap the bandwidth-bound phases of the other.  Each context is driven through the plain C ABI.
    ctxs = [dofs.Context(W, H, max_pairs=Bc, device=local, params=params) for _ in range(NC)]
    streams = [torch.cuda.ExternalStream(c.stream, device=dev) for c in ctxs]

    # this rank's block of the video, cut into steps of B pairs; context c takes pairs [c*Bc, (c+1)*Bc) of the step.
    # Every warm-up, probe and timed step has its own frames (the replay's work depends on the data).
    n_inputs = K + Wm + 1
    frames, outs = [], []
    for c, ctx in enumerate(ctxs):
        fl = []
        for i in range(n_inputs):
            f = torch.empty((Bc + 1, H, W, 3), dtype=torch.uint8, device=dev)
            ctx.synth_frames_dev(SEED, args.objects, (rank * n_inputs + i) * B + c * Bc, Bc + 1, f.data_ptr())
            fl.append(f)
        frames.append(fl)
        outs.append(dict(labels=torch.empty((Bc, H, W), dtype=torch.int32, device=dev),
                         boxes=torch.empty((Bc, MAXB * BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev),
                         nbox=torch.empty((Bc,), dtype=torch.int32, device=dev),
                         stats=torch.empty((Bc, STATS_DTYPE.itemsize), dtype=torch.uint8, device=dev)))
        ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def enqueue(c, i):  # asynchronous: no host wait
        o = outs[c]
        ctxs[c].process_dev(frames[c][i].data_ptr(), Bc + 1, o["labels"].data_ptr(), o["boxes"].data_ptr(),
                            o["nbox"].data_ptr(), MAXB, o["stats"].data_ptr())

    def sync_all():
        for ctx in ctxs:
            ctx.sync()  # also reports any deferred failure (overflow / non-convergence) of EVERY call since the last sync

    for i in range(Wm):
        for c in range(NC):
            enqueue(c, i)
        sync_all()

    # stagger the contexts by 1/NC of a sub-batch, once, at the start of the timed region (its cost is part of the
    # measured time): context c first sleeps c/NC of the time one sub-batch takes alone.  Nothing couples the streams
    # afterwards, so the phase difference persists and the latency-bound tail of one sub-batch keeps running under the
    # bandwidth-bound kernels of the others.
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    pe0.record(streams[0])
    enqueue(0, Wm)
    pe1.record(streams[0])
    ctxs[0].sync()
    alone_ms = min(max(pe0.elapsed_time(pe1), 1.0), 2000.0)  # one sub-batch alone on the GPU (device time)
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = sum(c.launch_count for c in ctxs) + 0
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
    ctxs[0].set_timing(True)  # per-kernel CUDA events on context 0's stream, live in the timed region
    for c in range(NC):
        ev0[c].record(streams[c])  # the timed region starts here for every context; the stagger is inside it
        if c > 0:
            with torch.cuda.stream(streams[c]):
                torch.cuda._sleep(int(alone_ms * 1e-3 * c / NC * 1.9e9))
    for i in range(K):  # every step of every context is enqueued without a host wait in between
        for c in range(NC):
            enqueue(c, Wm + 1 + i)
            if i == K - 1:
                ev1[c].record(streams[c])
    sync_all()
    stage_ms = {}
    for k, (ms, cnt) in ctxs[0].timing().items():  # the marks of context 0's last step
        stage_ms[k] = [ms * K, cnt * K]
    barrier()
    ctxs[0].set_timing(False)
    clk = clocks.stop()
    launches = sum(c.launch_count for c in ctxs) - launches0
    # first start to last end over all contexts: K steps of B pairs plus the stagger
    starts = [ev0[0].elapsed_time(e) for e in ev0]
    ends = [ev0[0].elapsed_time(e) for e in ev1]
    ms_total = max(ends) - min(min(starts), 0.0)
    timeline = {"sub_batch_alone_ms": alone_ms, "context_start_ms": [round(v, 2) for v in starts],
                "context_end_ms": [round(v, 2) for v in ends]}
    n_boxes_dev = int(sum(int(o["nbox"].sum().item()) for o in outs))

    # ---- the same kernels alone on the GPU (one context, nothing overlapping): the per-kernel roofline numbers
    iso_ms = {}
    n_iso = max(2, min(K, 3))
    ctxs[0].set_timing(True)
    for i in range(n_iso):
        enqueue(0, i)
        ctxs[0].sync()
        for k, (ms, cnt) in ctxs[0].timing().items():
            a = iso_ms.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += cnt
    ctxs[0].set_timing(False)

    # ---- end to end through the streaming entry points with HOST buffers: pinned frames in, labels + boxes out.  One
    # host thread; every context has two chunks in flight, the upload of a chunk runs under the kernels of the previous one.
    fmt = {"rle": capi.LABELS_RLE, "u16": capi.LABELS_U16, "i32": capi.LABELS_I32}[args.labels]
    n_host = 3  # distinct host chunks per context, cycled (the pair at the wrap-around is a scene cut)
    host = []
    for c in range(NC):
        chunks = []
        for j in range(n_host):
            hf = torch.empty((Bc + 1, H, W, 3), dtype=torch.uint8).pin_memory()
            hf.copy_(frames[c][j])
            chunks.append(hf)
        slots = []
        for _ in range(2):
            if fmt == capi.LABELS_RLE:
                lab = torch.empty((Bc, MAXR * RUN_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
            else:
                lab = torch.empty((Bc, H, W), dtype=torch.int32 if fmt == capi.LABELS_I32 else torch.int16).pin_memory()
            s = dict(labels=lab, n_runs=torch.zeros((Bc,), dtype=torch.int32).pin_memory(),
                     boxes=torch.empty((Bc, MAXB * BOX_DTYPE.itemsize), dtype=torch.uint8).pin_memory(),
                     nbox=torch.zeros((Bc,), dtype=torch.int32).pin_memory(),
                     stats=torch.empty((Bc, STATS_DTYPE.itemsize), dtype=torch.uint8).pin_memory())
            s["out"] = capi.Outputs(fmt, s["labels"].data_ptr(), s["n_runs"].data_ptr(), MAXR, s["boxes"].data_ptr(),
                                    s["nbox"].data_ptr(), MAXB, s["stats"].data_ptr())
            slots.append(s)
        host.append(dict(chunks=chunks, slots=slots, submitted=0, inflight=0, boxes_seen=0))
    # device buffers of the NCCL gather: each context packs the boxes of its last chunk into its slice
    CAPC = Bc * 64
    g_boxes = torch.zeros((NC, CAPC * BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    g_counts = torch.zeros((NC,), dtype=torch.int32, device=dev)
    pack_done = [torch.cuda.Event() for _ in ctxs]

    def ck(ctx, rc):
        if rc != 0:
            raise dofs.DofsError(rc, ctx.L.dofs3d_last_error(ctx.h).decode())

    def collect(c):
        ctx, h = ctxs[c], host[c]
        n = C.c_int(0)
        ck(ctx, ctx.L.dofs3d_stream_collect(ctx.h, C.byref(n)))
        slot = h["slots"][(h["submitted"] - h["inflight"]) & 1]
        h["boxes_seen"] += int(slot["nbox"][:n.value].sum())
        h["inflight"] -= 1
        return n.value

    def submit(c):
        ctx, h = ctxs[c], host[c]
        if h["inflight"] == 2:
            collect(c)
        k = h["submitted"]
        chunk = h["chunks"][k % n_host]
        first = k == 0
        ptr = chunk.data_ptr() if first else chunk.data_ptr() + N * 3  # later chunks: the carried frame is on the device
        ck(ctx, ctx.L.dofs3d_stream_submit(ctx.h, C.c_void_p(ptr), Bc + 1 if first else Bc, C.byref(h["slots"][k & 1]["out"])))
        h["submitted"] += 1
        h["inflight"] += 1

    def host_steps(n_steps, gather):
        for _ in range(n_steps):
            for c in range(NC):
                submit(c)
        pairs = 0
        for c in range(NC):
            while host[c]["inflight"]:
                pairs += collect(c)
        gms = None
        if gather and world > 1:
            from denseopticalflowsegmentation3d_b200 import shard
            tg = time.perf_counter()
            for c in range(NC):
                ctxs[c].pack_boxes_dev(Bc, g_boxes[c].data_ptr(), CAPC, g_counts[c:c + 1].data_ptr())
                pack_done[c].record(streams[c])
            for c in range(NC):
                torch.cuda.current_stream().wait_event(pack_done[c])  # the collectives run behind the packs, no host wait
            all_counts, all_boxes = shard.gather_packed(g_counts, g_boxes)
            total = int(all_counts.sum().item())  # device -> host read of the gathered result
            torch.cuda.synchronize()
            gms = (1e3 * (time.perf_counter() - tg), total)
        return gms

    for c in range(NC):
        ck(ctxs[c], ctxs[c].L.dofs3d_stream_begin(ctxs[c].h))
    host_steps(min(Wm, 2), True)  # also the first collective of the communicator (channel set-up) stays out of the timing
    barrier()
    for h in host:
        h["boxes_seen"] = 0
    t0 = time.perf_counter()
    gathered = host_steps(K, True)
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    # the same gather once more, all ranks lined up by the barrier and timed on the device: the figure inside the e2e
    # region above also contains the ranks' arrival skew (they finish their last chunk a few ms apart)
    gather_aligned_ms = None
    if world > 1:
        from denseopticalflowsegmentation3d_b200 import shard
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for c in range(NC):
            ctxs[c].pack_boxes_dev(Bc, g_boxes[c].data_ptr(), CAPC, g_counts[c:c + 1].data_ptr())
            pack_done[c].record(streams[c])
        for c in range(NC):
            torch.cuda.current_stream().wait_event(pack_done[c])
        shard.gather_packed(g_counts, g_boxes)
        ev1.record()
        torch.cuda.synchronize()
        gather_aligned_ms = ev0.elapsed_time(ev1)
    n_boxes_host = int(sum(h["boxes_seen"] for h in host))
    label_bytes = {"rle": MAXR * RUN_DTYPE.itemsize + 4, "u16": N * 2, "i32": N * 4}[args.labels]
    h2d = B * N * 3  # steady state: the boundary frame of a chunk stays on the device
    d2h = B * (label_bytes + MAXB * BOX_DTYPE.itemsize + 4 + STATS_DTYPE.itemsize)

    # ---- max over ranks
    t = torch.tensor([ms_total, e2e_ms, float(launches), gathered[0] if gathered else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_boxes = gathered[1]
    else:
        total_boxes = n_boxes_dev
    ms_total, e2e_ms = float(t[0].item()), float(t[1].item())
    gather_ms = float(t[3].item()) if world > 1 else None

    if rank == 0:
        value = world * B * K / (ms_total / 1e3)
        e2e_value = world * B * K / (e2e_ms / 1e3)
        peak, peak_src = measured_peak_gbs()
        # Dominant bandwidth-bound kernel of the path: k_box_solve7f (FarnebackUpdateFlow_Blur: 15x15 box mean of the five
        # matrix channels + the 2x2 solve).  The roofline figure is taken on its full-resolution launches (3 per call, one per
        # iteration of pyramid level 0): algorithmic bytes per launch = (20 B of M read once + 8 B of flow written) per pixel.
        box_bytes = 28 * N * Bc

        def roof_of(stages):
            sc = stages.get("flow.box_solve.L0", [0.0, 0])
            if sc[1] == 0:
                return None
            ms_launch = sc[0] / sc[1]
            achieved = box_bytes / (ms_launch / 1e3) / 1e9
            return {"achieved": achieved, "frac": achieved / peak, "ms_per_launch": ms_launch, "launches_timed": sc[1]}

        def stage_gbs(stages, name, bytes_per_launch):
            sc = stages.get(name, [0.0, 0])
            if sc[1] == 0:
                return None
            g = bytes_per_launch / (sc[0] / sc[1] / 1e3) / 1e9
            return {"achieved": g, "frac": g / peak, "ms_per_launch": sc[0] / sc[1], "algorithmic_bytes_per_launch": bytes_per_launch}

        live, iso = roof_of(stage_ms), roof_of(iso_ms)
        traffic = committed_traffic()
        box_traffic = None
        box_keys = [k for k in (traffic or {}).get("kernels", {}) if k.startswith("k_box_solve7")]
        if traffic and traffic.get("pairs") and box_keys:
            # ncu DRAM bytes of the three full-resolution launches: the kernel's bytes scale with the pixel count
            kb = traffic["kernels"][box_keys[0]]
            box_traffic = (kb["dram_read_MB"] + kb["dram_write_MB"]) * 1e6 / traffic["pairs"] * Bc / (3 * 1.328125) * N / (1920 * 1080)
        roof = None
        if iso:
            # the per-kernel figure is the kernel alone on the GPU (CUDA events around every launch, one context, right after
            # the timed region): inside the timed region NC contexts overlap, so an event-bracketed launch there also contains
            # the time it spends sharing the SMs and HBM with the other contexts' kernels (reported as in_timed_region)
            roof = {"bound": "hbm", "kernel": "k_box_solve7f (15x15 box mean of M + 2x2 solve), full-resolution launches",
                    "achieved": iso["achieved"], "peak": peak, "unit": "GB/s", "frac": iso["frac"],
                    "traffic": box_traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": box_bytes,
                    "ms_per_launch": iso["ms_per_launch"], "launches_timed": iso["launches_timed"],
                    "how": "CUDA events on the launching stream around each launch, bench.py, one context alone on the GPU",
                    "in_timed_region": dict(live or {}, note=f"context 0's launches while {NC - 1} other context(s) share the GPU"),
                    "other_kernels": {
                        "k_blur_fused_tma<12> (flow blur, tiles staged by TMA bulk copies; 16 B per pixel)":
                            stage_gbs(iso_ms, "flow_blur", 16 * N * Bc),
                        "k_radix_onesweep<u32> (one 8-bit pass of the merge-time sort, key+payload read and written)":
                            stage_gbs(iso_ms, "time_sort.scatter", 16 * N * Bc),
                        "k_radix_onesweep<u32> (one 8-bit pass of the event sort, 32-bit (wave, winner) keys)":
                            stage_gbs(iso_ms, "event_sort.scatter", 16 * N * Bc)}}
        stages = {k: round(v[0] / K, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1][0])}
        stages_iso = {k: round(v[0] / n_iso, 4) for k, v in sorted(iso_ms.items(), key=lambda kv: -kv[1][0])}
        # whole path: the DRAM bytes the path actually moves (ncu, per kernel, profiles/r02_traffic*.json), not SURVEY's
        # 1476 B/px estimate of the reference-shaped pipeline
        whole = None
        if traffic:
            bpp = traffic["dram_bytes_per_pixel_per_pair"]
            gbs = bpp * N * B / (ms_total / K / 1e3) / 1e9
            whole = {"measured_dram_bytes_per_pixel_per_pair": bpp, "source": traffic["file"],
                     "achieved_dram_GBps": gbs, "frac_of_peak": gbs / peak,
                     "note": "per-kernel ncu dram__bytes_read+write of one 32-pair 1080p call, scaled by pixels x pairs"}
        cfg = workload_config(args, world)
        cfg["contexts_per_gpu"] = NC
        line = {
            "metric": args.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 flow / f64 edge weights / u32+u64 keys", "data": "synthetic", "config": cfg,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K, "labels": args.labels,
                    "api": f"dofs3d_stream_submit / dofs3d_stream_collect (host pointers, pinned), {NC} contexts, one host "
                           f"thread, two chunks in flight per context; labels as "
                           + {"rle": "run-length records", "u16": "uint16 images", "i32": "int32 images"}[args.labels]
                           + ("; NCCL gather of the last step's boxes inside the timed region" if world > 1 else "")},
            "gpu_launches": int(t[2].item()), "clocks": clk, "roofline": roof, "whole_path": whole,
            "timeline": timeline, "stage_ms_per_step_context0_live": stages, "stage_ms_per_call_alone": stages_iso,
            "boxes_found": total_boxes, "boxes_found_e2e_rank0": n_boxes_host, "nccl_gather_boxes_ms": gather_ms, "nccl_gather_boxes_ms_ranks_aligned": gather_aligned_ms,
            "device_bytes": sum(c.device_bytes for c in ctxs),
        }
        if world == 1 and not args.no_cpu_baseline:
            rows = cpu_sample_rows(H, 20.0, FULL_PAIR_CPU_SECONDS * (N / (1920 * 1080)))
            for c in ctxs:
                c.close()
            r = run_cpu_reference(args, 1, rows, 1, 0)
            line["cpu_baseline"] = {
                "value": r["pairs_per_s"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                "sample": f"1 process, 1 synthetic pair, central band of {rows}/{H} rows at full width (counted as "
                          f"{rows / H:.3f} pair, {r['seconds']:.1f} s; a band under-costs the super-linear reference — "
                          f"`--impl reference` calibrates it against a full frame); cv2 flow+blur + "
                          + ("unchanged reference sources (oracle/_ref)" if r["kind"] == "reference" else "oracle port")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
