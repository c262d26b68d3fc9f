"""Randomised parity sweep on the GPU: many small flow fields of random shapes and tie structures through
dofs3d_segment, every one compared with the CPU oracle (roots, pixel sets, sizes, boxes, mean-flow bits, gate counters).
    python tools/fuzz_parity.py [cases=300] [seed=0]  > profiles/r02_fuzz.log
Not part of the test suite (it needs a GPU and a minute); its log is kept with the profiles."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import denseopticalflowsegmentation3d_b200 as dofs  # noqa: E402
from denseopticalflowsegmentation3d_b200.capi import box_pixel_sets, LABELS_RLE, runs_to_labels  # noqa: E402
from conftest import random_flow  # noqa: E402
from oracle import cpu  # noqa: E402
from test_gpu_parity import compare_boxes, near_tie_field  # noqa: E402


def make_field(rng, W, H):
    kind = int(rng.integers(0, 7))
    seed = int(rng.integers(0, 1 << 30))
    if kind == 0:
        f = random_flow(seed, W, H, scale=float(rng.uniform(1, 8)), flat=bool(rng.integers(0, 2)))
    elif kind == 1:
        f = np.round(random_flow(seed, W, H, scale=float(rng.uniform(2, 10)), flat=False) * 2) / 2   # huge tie classes
    elif kind == 2:
        f = near_tie_field(W, H, seed)
    elif kind == 3:
        f = np.zeros((H, W, 2), np.float32)
        for _ in range(int(rng.integers(1, 6))):                                                  # plateaus
            x0, y0 = int(rng.integers(0, W)), int(rng.integers(0, H))
            f[y0:y0 + int(rng.integers(1, H + 1)), x0:x0 + int(rng.integers(1, W + 1))] = rng.integers(-4, 5, 2).astype(np.float32)
    elif kind == 4:
        f = (rng.normal(size=(H, W, 2)) * rng.uniform(0.1, 5)).astype(np.float32)               # white noise
    elif kind == 5:
        f = np.zeros((H, W, 2), np.float32)
        f[..., 1] = np.linspace(0, float(rng.uniform(1, 9)), H, dtype=np.float32)[:, None]            # ramp
        f[..., 0] = np.linspace(0, float(rng.uniform(0, 3)), W, dtype=np.float32)[None, :]
    else:
        f = random_flow(seed, W, H, scale=3.0)
        f[rng.random((H, W)) < 0.3] = 0.0                                                         # scattered exact zeros
    return np.ascontiguousarray(f, np.float32)


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    port = cpu.port()
    persp, inv, up = port.get_mats()
    t0 = time.time()
    n_boxes = n_px = 0
    for k in range(cases):
        W, H = int(rng.integers(2, 201)), int(rng.integers(2, 141))
        nb = 8 if rng.integers(0, 4) else 4
        min_size = int(rng.integers(8, max(9, W * H // 20)))
        thr = float(rng.choice([0.3, 0.0, -1.0, 0.6]))
        n = int(rng.integers(1, 4))
        fields = np.stack([make_field(rng, W, H) for _ in range(n)])
        p = dofs.default_params()
        p.neighbors, p.min_size, p.score_threshold = nb, min_size, thr
        try:
            with dofs.Context(W, H, max_pairs=n, params=p) as c:
                out = c.segment(fields, already_blurred=True, max_boxes=4096)
                rle = c.segment_ex(fields, True, LABELS_RLE, max_boxes=4096, max_runs=W * H)
        except dofs.DofsError as e:
            if e.status != -4:
                raise
            print(f"case {k}: more than 4096 boxes or candidates overflow, skipped ({e})")
            continue
        for i in range(n):
            res = port.segment(fields[i], persp, inv, up, neighbors=nb, score_threshold=thr, min_size=min_size)
            boxes = out["boxes"][i]
            compare_boxes(boxes, box_pixel_sets(out["labels"][i], boxes), res["entries"], W)
            st, cn = out["stats"][i], res["counters"]
            assert st["n_candidates"] == cn["get_score"] and st["n_merges"] == cn["merges"] == W * H - 1, (k, i)
            assert np.array_equal(runs_to_labels(rle["labels"][i], rle["n_runs"][i], W * H), out["labels"][i].reshape(-1))
            n_boxes += len(boxes)
            n_px += W * H
        if (k + 1) % 50 == 0:
            print(f"{k + 1} cases ok ({n_boxes} segments, {n_px} pixels, {time.time() - t0:.0f} s)", flush=True)
    print(f"fuzz parity: {cases} cases, {n_boxes} segments, {n_px} pixels: all identical to the oracle")


if __name__ == "__main__":
    main()
