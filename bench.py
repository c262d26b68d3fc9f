#!/usr/bin/env python
"""Benchmark of the hot path: 1080p frame pairs -> Farneback flow -> flow-graph clustering -> 3D boxes.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step = one batch of B consecutive frame pairs (B+1 frames) of the synthetic 1080p video through
dofs3d_process*.  `value` times the device-pointer entry point with the frames already in HBM;
`e2e` times the host-pointer entry point (pinned host frames in, labels + boxes out, copies inside the
timed region).  Frames are sharded contiguously over ranks; there is no collective on the data path
(weak scaling), only a final gather of per-rank box counts.
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

METRIC = "1080p frame-pairs/s (flow->segment->3D)"
UNIT = "frame-pairs/s"
SEED = 1234
N_OBJECTS = 8
FULL_PAIR_CPU_SECONDS = 40.0  # rough cost of one 1080p pair through the reference on one core (33 s)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frame pairs per step per GPU")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=200.0, help="seconds the reference arm may spend in total")
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"synthetic {args.width}x{args.height} video, {N_OBJECTS} moving textured objects (BASELINE configs[2]/[4] "
                    f"streamed in batches of {args.batch} frame pairs per GPU per step; pair i = frames i, i+1)",
        "pairs_per_step_per_gpu": args.batch,
        "frame": [args.width, args.height],
        "sharding": f"contiguous blocks of frames over {world} rank(s), no collective on the data path",
        "cache": "inputs larger than L2 (one step reads %.0f MB of BGR frames and streams >10 GB of intermediates)"
                 % ((args.batch + 1) * args.width * args.height * 3 / 1e6),
    }


# --------------------------------------------------------------------------------------------------
# the reference's CPU path (cv2 for the OpenCV calls, oracle/_ref = the unchanged reference sources)
# --------------------------------------------------------------------------------------------------
_worker_state = {}


def _cpu_worker_init(width, height, rows, seed_base):
    """Each worker renders its own frame pair (untimed) and keeps the band of `rows` rows it will process."""
    import cv2
    from denseopticalflowsegmentation3d_b200 import synth
    from oracle import cpu
    cv2.setNumThreads(1)
    idx = int(os.environ.get("DOFS_WORKER_INDEX", "0"))
    wid = os.getpid()
    first = (wid * 7 + idx) % 16
    fr = synth.frames(seed_base, N_OBJECTS, first, 2, width, height)
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


def cpu_sample_rows(height, seconds_per_step):
    frac = min(1.0, max(seconds_per_step / FULL_PAIR_CPU_SECONDS, 0.125))
    rows = int(round(height * frac / 8)) * 8
    return max(min(rows, height), 64)


def run_cpu_reference(args, cores, rows, steps, warmup):
    """`cores` worker processes, each pushing one frame pair (band of `rows` rows) per step through the reference."""
    if cores == 1:  # in-process (used for the cpu_baseline key of the GPU arm: no fork after CUDA start-up)
        kind = _cpu_worker_init(args.width, args.height, rows, SEED)
        for _ in range(warmup):
            _cpu_pair(0)
        t0 = time.perf_counter()
        n_seg = sum(_cpu_pair(0)[1] for _ in range(steps))
        dt = time.perf_counter() - t0
        return {"kind": kind, "seconds": dt, "pairs_per_s": steps * (rows / args.height) / dt,
                "ms_per_step": 1e3 * dt / steps, "segments": n_seg}
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_worker_init, initargs=(args.width, args.height, rows, SEED)) as pool:
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
    rows = cpu_sample_rows(args.height, args.cpu_budget / max(args.steps + args.warmup, 1))
    r = run_cpu_reference(args, cores, rows, args.steps, args.warmup)
    sample = (f"{cores} processes x 1 synthetic pair per step, central band of {rows}/{args.height} rows at full width "
              f"(counted as {rows / args.height:.3f} pair); cv2 flow+blur, then "
              + ("the unchanged reference graph.cpp+lifting_3d.cpp (oracle/_ref)" if r["kind"] == "reference"
                 else "the oracle port (oracle/_ref not built)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 geometry / f64 weights", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": r["pairs_per_s"], "unit": UNIT, "cores": cores, "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


def ncu_traffic_bytes(kernel):
    """dram bytes per launch of `kernel` from the committed ncu summary, if there is one."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(kernel)
        except Exception:
            return None
    return None


def ours(args):
    import torch
    import denseopticalflowsegmentation3d_b200 as dofs
    from denseopticalflowsegmentation3d_b200.capi import BOX_DTYPE, STATS_DTYPE

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    W, H, B, K, Wm = args.width, args.height, args.batch, args.steps, args.warmup
    N = W * H
    MAXB = 256
    ctx = dofs.Context(W, H, max_pairs=B, device=local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))

    # this rank's block of the video: frames [rank*T, rank*T + T], T = B pairs per step; a few distinct steps, cycled
    n_inputs = min(K + Wm, 4)
    frames = [torch.empty((B + 1, H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(n_inputs)]
    for i, f in enumerate(frames):
        ctx.synth_frames_dev(SEED, N_OBJECTS, (rank * (K + Wm) + i) * B, B + 1, f.data_ptr())
    d_labels = torch.empty((B, H, W), dtype=torch.int32, device="cuda")
    d_boxes = torch.empty((B, MAXB * BOX_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    d_nbox = torch.empty((B,), dtype=torch.int32, device="cuda")
    d_stats = torch.empty((B, STATS_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    ctx.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(i):
        ctx.process_dev(frames[i % n_inputs].data_ptr(), B + 1, d_labels.data_ptr(), d_boxes.data_ptr(), d_nbox.data_ptr(),
                        MAXB, d_stats.data_ptr())

    ctx.set_timing(True)
    for i in range(Wm):
        step_dev(i)
    stage_ms = {}
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for i in range(K):
        step_dev(Wm + i)
        for k, (ms, cnt) in ctx.timing().items():
            a = stage_ms.setdefault(k, [0.0, 0])
            a[0] += ms
            a[1] += cnt
    ev1.record(stream)
    barrier()
    clk = clocks.stop()
    launches = ctx.launch_count - launches0
    ms_total = ev0.elapsed_time(ev1)
    n_boxes_dev = int(d_nbox.sum().item())

    # ---- end to end through the host-pointer entry point: pinned frames in, labels + boxes out
    h_frames = torch.empty((B + 1, H, W, 3), dtype=torch.uint8).pin_memory()
    h_frames.copy_(frames[0])
    h_labels = torch.empty((B, H, W), dtype=torch.int32).pin_memory()
    h_boxes = torch.empty((B, MAXB * BOX_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    h_nbox = torch.empty((B,), dtype=torch.int32).pin_memory()
    h_stats = torch.empty((B, STATS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    import ctypes as C

    def step_host():
        rc = ctx.L.dofs3d_process(ctx.h, C.c_void_p(h_frames.data_ptr()), B + 1, C.c_void_p(h_labels.data_ptr()),
                                  C.c_void_p(h_boxes.data_ptr()), C.c_void_p(h_nbox.data_ptr()), MAXB,
                                  C.c_void_p(h_stats.data_ptr()))
        if rc != 0:
            raise dofs.DofsError(rc, ctx.L.dofs3d_last_error(ctx.h).decode())

    ctx.set_timing(False)
    for _ in range(min(Wm, 2)):
        step_host()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(K):
        step_host()
    e1.record(stream)
    barrier()
    e2e_wall_ms = 1e3 * (time.perf_counter() - t0)
    e2e_ms = max(e0.elapsed_time(e1), e2e_wall_ms if world == 1 else e0.elapsed_time(e1))
    n_boxes_host = int(h_nbox.sum().item())
    h2d = (B + 1) * N * 3
    d2h = B * N * 4 + B * MAXB * BOX_DTYPE.itemsize + B * 4 + B * STATS_DTYPE.itemsize

    # ---- max over ranks
    t = torch.tensor([ms_total, e2e_ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts = torch.tensor([n_boxes_dev], dtype=torch.int64, device="cuda")
        gathered = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts)  # per-rank box counts to every rank over NCCL/NVLink (results, not data path)
        total_boxes = int(sum(int(g.item()) for g in gathered))
    else:
        total_boxes = n_boxes_dev
    ms_total, e2e_ms = float(t[0].item()), float(t[1].item())

    if rank == 0:
        value = world * B * K / (ms_total / 1e3)
        e2e_value = world * B * K / (e2e_ms / 1e3)
        peak, peak_src = measured_peak_gbs()
        # dominant kernel: the radix-sort scatter pass over the 4N edge slots of all B frames
        # algorithmic bytes per launch: read (8 B key + 4 B payload) + write (8 + 4) per slot
        slots = 4 * N * B
        sc = stage_ms.get("edge_sort.scatter", [0.0, 0])
        roof = None
        if sc[1] > 0:
            ms_launch = sc[0] / sc[1]
            achieved = slots * 24 / (ms_launch / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": "k_radix_scatter (edge sort pass)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_traffic_bytes("k_radix_scatter"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": slots * 24, "ms_per_launch": ms_launch, "launches_timed": sc[1]}
        stages = {k: round(v[0] / K, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1][0])}
        whole_bytes = 1476 * N * B  # algorithmic HBM bytes per pixel per pair, whole path
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 flow / f64 edge weights / u64 keys", "data": "synthetic", "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / K, "api": "dofs3d_process (host pointers, pinned)"},
            "gpu_launches": int(t[2].item()), "clocks": clk, "roofline": roof,
            "whole_path": {"algorithmic_GBps": whole_bytes / (ms_total / K / 1e3) / 1e9,
                           "frac_of_peak": whole_bytes / (ms_total / K / 1e3) / 1e9 / peak,
                           "bytes_per_pixel_per_pair": 1476},
            "stage_ms_per_step": stages, "boxes_found": total_boxes, "boxes_found_e2e_rank0": n_boxes_host,
            "device_bytes": ctx.device_bytes,
        }
        if world == 1 and not args.no_cpu_baseline:
            rows = cpu_sample_rows(H, 20.0)
            ctx.close()
            r = run_cpu_reference(args, 1, rows, 1, 0)
            line["cpu_baseline"] = {
                "value": r["pairs_per_s"], "unit": UNIT, "cores": 1, "kind": r["kind"],
                "sample": f"1 process, 1 synthetic pair, central band of {rows}/{H} rows at full width (counted as "
                          f"{rows / H:.3f} pair, {r['seconds']:.1f} s); cv2 flow+blur + "
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
