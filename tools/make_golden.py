"""Generates tests/golden/*.npz IN THE AUTHORING CONTAINER (needs /root/reference, cv2 and oracle/_ref):
inputs and outputs of the reference itself, so that the GPU box (where /root/reference does not
exist) can check both the CPU restatement (oracle port) and the CUDA path against them.

  pair_1052_1053.npz  the repo's own frame pair (config 1 of BASELINE.json): gray frames, cv2 4.13
                      Farneback flow and blurred flow (segment.cpp:97-101,52), and what the UNCHANGED
                      reference sources (oracle/_ref) return for build_graph + segment_graph +
                      get_best_segments on that blurred flow, plus its gate counters.
  synth_320x180.npz   same for a seeded synthetic pair small enough for quick tests.
  lift_random.npz        get_bottom_variants of the unchanged reference on the gtest golden input
                      (cpp/tests/test_liftig_3d.cpp:179-227) and on seeded random problems.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import cpu  # noqa: E402
from denseopticalflowsegmentation3d_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_ref(flow_blurred, R):
    persp, inv, up = R.get_mats()
    res = R.segment(flow_blurred, persp, inv, up)
    ents = res["entries"]
    out = dict(
        n_edges=res["n_edges"], num_sets=res["num_sets"],
        root=np.array([e["root"] for e in ents], np.int32),
        score=np.array([e["score"] for e in ents], np.float64),
        move=np.array([e["move"] for e in ents], np.float64),
        size=np.array([e["size"] for e in ents], np.int32),
        cls=np.array([e["sol"]["cls"] for e in ents], np.int32),
        w_error=np.array([e["sol"]["w_error"] for e in ents], np.float64),
        h_error=np.array([e["sol"]["h_error"] for e in ents], np.float64),
        orient=np.array([e["sol"]["orient"] for e in ents], np.float64),
        ps_bev=np.array([e["sol"]["ps_bev"] for e in ents], np.float32).reshape(-1, 4, 2),
        rectangle=np.array([e["sol"]["rectangle"] for e in ents], np.float32).reshape(-1, 4, 2),
        lower_face=np.array([e["sol"]["lower_face"] for e in ents], np.float32).reshape(-1, 4, 2),
        upper_face=np.array([e["sol"]["upper_face"] for e in ents], np.float32).reshape(-1, 4, 2),
        pixels=np.concatenate([e["pixels"] for e in ents]).astype(np.int32) if ents else np.zeros(0, np.int32),
        pixel_offsets=np.cumsum([0] + [e["size"] for e in ents]).astype(np.int64),
    )
    s, e, w = R.build_graph(flow_blurred)
    out.update(edges_sha_start=sha(s), edges_sha_end=sha(e), edges_sha_weight=sha(w),
               edges_head_start=s[:4096], edges_head_end=e[:4096], edges_head_weight=w[:4096],
               edges_tail_start=s[-4096:], edges_tail_end=e[-4096:], edges_tail_weight=w[-4096:])
    R.set_counting(True)
    R.segment(flow_blurred, persp, inv, up)
    cnt = R.get_counts()
    R.set_counting(False)
    out["counter_names"] = np.array(sorted(cnt))
    out["counter_values"] = np.array([cnt[k] for k in sorted(cnt)], np.int64)
    return out


def flow_chain(g0, g1):
    flow = cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    return flow, cv2.GaussianBlur(flow, (0, 0), 3.0)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = cpu.ref()
    # --- the repo's own pair
    im1 = cv2.imread("/root/reference/data/frame_1052.png")
    im2 = cv2.imread("/root/reference/data/frame_1053.png")
    g0, g1 = cv2.cvtColor(im1, cv2.COLOR_BGR2GRAY), cv2.cvtColor(im2, cv2.COLOR_BGR2GRAY)
    flow, fb = flow_chain(g0, g1)
    np.savez_compressed(os.path.join(OUT, "pair_1052_1053.npz"), gray0=g0, gray1=g1, bgr0_head=im1[:8], gray0_head=g0[:8],
                        flow=flow, flow_blurred=fb, cv2_version=cv2.__version__, **run_ref(fb, R))
    # --- small synthetic pair
    W, H = 320, 180
    fr = synth.frames(77, 5, 0, 2, W, H)
    s0, s1 = cv2.cvtColor(fr[0], cv2.COLOR_BGR2GRAY), cv2.cvtColor(fr[1], cv2.COLOR_BGR2GRAY)
    sflow, sfb = flow_chain(s0, s1)
    sfb = sfb * np.float32(2.0)  # larger motion so that several clusters pass the gates at this size
    np.savez_compressed(os.path.join(OUT, "synth_320x180.npz"), bgr=fr, gray0=s0, gray1=s1, flow=sflow,
                        flow_blurred=sfb, cv2_version=cv2.__version__, **run_ref(sfb, R))
    # --- lifting
    persp, inv, up = R.get_mats()
    rng = np.random.default_rng(5)
    n = 512
    xmin = rng.integers(0, 560, n)
    ymin = rng.integers(36, 300, n)
    box = np.stack([xmin, ymin, xmin + rng.integers(8, 200, n), ymin + rng.integers(8, 150, n)], 1).astype(np.int32)
    d = rng.normal(size=(n, 2)).astype(np.float32) * 3
    d[::7, 0] = 0
    cls = rng.integers(0, 3, n).astype(np.int32)
    sols = [R.get_bottom_variants(d[i], box[i], persp, inv, up[cls[i]], cls[i]) for i in range(n)]
    # the gtest golden input (test_liftig_3d.cpp:181-186)
    kat_mat = np.array([20.137783, -13.474492, 402.17429, 5.1163507, 800.33502, -62251.332, 0.00039356545, 0.039720595, 1.0], np.float32)
    np.savez_compressed(
        os.path.join(OUT, "lift_random.npz"), persp=persp, inv=inv, upper=up, dir=d, box=box, cls=cls,
        has_rect=np.array([s["has_rectangle"] for s in sols]), w_error=np.array([s["w_error"] for s in sols]),
        h_error=np.array([s["h_error"] for s in sols]), orient=np.array([s["orient"] for s in sols]),
        ps_bev=np.array([s["ps_bev"] for s in sols]), rectangle=np.array([s["rectangle"] for s in sols]),
        lower_face=np.array([s["lower_face"] for s in sols]), upper_face=np.array([s["upper_face"] for s in sols]))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
