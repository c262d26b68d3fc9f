"""Golden vectors for the Felzenszwalb mode: runs the REFERENCE's own Python (/root/reference/graph.py: build_graph,
segment_graph_flow; main.py's diff and threshold restated below because main.py opens GUI windows on import) on small
flow fields, in the authoring container, and stores inputs + component labels in tests/golden/fh_cases.npz.

    python tools/make_golden_fh.py

The reference addresses its array as img[x][y] (graph.py:19, main.py:311) with node id = y * width + x, so it is given
flow.transpose(1, 0, 2): then x is the column of the usual [H][W][2] field and ids are row * W + col.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
import graph as ref_graph  # noqa: E402  (the reference's module)


def diff(img, x1, y1, x2, y2):  # main.py:310-312
    _out = np.sum((img[x1, y1] - img[x2, y2]) ** 2)
    return np.sqrt(_out)


def threshold(size, const):  # main.py:314-315
    return (const * 1.0 / size)


def run_reference(flow, K, min_size, neighbors8):
    H, W = flow.shape[:2]
    img = np.ascontiguousarray(flow.transpose(1, 0, 2))  # img[x][y]
    edges = ref_graph.build_graph(img, W, H, diff, neighbors8)
    with contextlib.redirect_stdout(io.StringIO()):  # merge_components prints every candidate weight
        forest = ref_graph.segment_graph_flow(img, edges, W * H, K, min_size, threshold, diff, W)
        # the first two passes alone are the reference's segment_graph (graph.py:133-153)
        forest2 = ref_graph.segment_graph(img, edges, W * H, K, min_size, threshold, W)
    lab = lambda fo: np.array([fo.find(i) for i in range(W * H)], np.int32).reshape(H, W)  # noqa: E731
    return lab(forest), forest.num_sets, lab(forest2), forest2.num_sets


def blocks(seed, W, H, n_blocks, step, noise):
    """Piecewise-constant flow (rigidly moving rectangles `step` apart in flow space) plus smooth noise."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import random_flow
    rng = np.random.default_rng(seed)
    f = random_flow(seed, W, H, scale=noise, flat=False)
    for k in range(n_blocks):
        x0, y0 = int(rng.integers(0, W - 2)), int(rng.integers(0, H - 2))
        w, h = int(rng.integers(3, max(W // 2, 4))), int(rng.integers(3, max(H // 2, 4)))
        f[y0:y0 + h, x0:x0 + w] += np.float32(step) * rng.integers(-3, 4, size=2).astype(np.float32)
    return np.ascontiguousarray(f.astype(np.float32))


def fields():
    g = np.load(os.path.join(ROOT, "tests", "golden", "synth_320x180.npz"))
    fb = g["flow_blurred"]
    yield "synth_crop", np.ascontiguousarray(fb[60:132, 100:196]) * np.float32(4.0), 10.0, 40, True
    yield "blocks_8", blocks(5, 64, 40, 6, 7.0, 1.5), 3.0, 20, True
    yield "blocks_4", blocks(6, 48, 36, 5, 9.0, 2.0), 3.0, 20, False
    yield "ties", np.round(blocks(7, 40, 30, 5, 6.0, 3.0)), 2.0, 10, True
    yield "thin", blocks(8, 70, 9, 3, 8.0, 1.0), 5.0, 8, True


def main():
    out = {}
    for name, f, K, ms, n8 in fields():
        labels, n, labels2, n2 = run_reference(f, K, ms, n8)
        print(name, f.shape, "components after segment_graph", n2, "after segment_graph_flow", n)
        out[name + "_flow"] = f.astype(np.float32)
        out[name + "_labels"] = labels
        out[name + "_labels_stage2"] = labels2
        out[name + "_params"] = np.array([K, ms, 8 if n8 else 4], np.float64)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fh_cases.npz"), **out)


if __name__ == "__main__":
    main()
