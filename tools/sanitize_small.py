"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel of the path on small frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import denseopticalflowsegmentation3d_b200 as d
from denseopticalflowsegmentation3d_b200 import synth

W, H, n = 160, 96, 2
fr = synth.frames(3, 4, 0, n + 1, W, H)
with d.Context(W, H, max_pairs=n) as c:
    out = c.process(fr)
    print("process boxes", out["n_boxes"].tolist(), "levels", out["stats"]["n_levels"].tolist())
    g = c.gray(fr)
    f = c.flow(g[:-1], g[1:])
    seg = c.segment(f, already_blurred=False)
    c.paint(n, 0.7)
    rng = np.random.default_rng(0)
    fb = (rng.normal(size=(H, W, 2)) * 2).astype(np.float32)
    c.edges_sorted(fb)
    c.lift([[1.0, 2.0]], [[10, 20, 60, 70]], [1])
    # compact label formats, per-node accessors, the streaming entry points (carried frame, two staging buffers)
    from denseopticalflowsegmentation3d_b200 import capi
    rle = c.process_ex(fr, capi.LABELS_RLE, max_runs=4096)
    c.process_ex(fr, capi.LABELS_U16)
    c.scored_merges(0)
    c.node_state(0, int(out["stats"][0]["final_root"]))
    clip = synth.frames(3, 4, 0, 7, W, H)
    st = c.process_stream(clip, label_format=capi.LABELS_RLE, max_runs=4096)
    print("runs", rle["n_runs"].tolist(), "stream pairs", len(st["boxes"]))
# near-tie field: long prefix runs and the 64-bit fallback
fb = np.zeros((200, 300, 2), np.float32)
fb[..., 1] = np.arange(200, dtype=np.float32)[:, None]
fb[..., 0] = (np.random.default_rng(1).random((200, 300)) * 3e-4).astype(np.float32)
with d.Context(300, 200) as c:
    st = c.segment(fb, already_blurred=True)["stats"][0]
    print("fallback", st["sort_fallback"])
print("ok")
