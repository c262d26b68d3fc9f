"""GPU parity at the benchmarked sizes (BASELINE configs 2 and 4) against live cv2 and against the UNCHANGED
reference (oracle/_ref), through the C ABI.

Flow bar.  The reference pins nothing at the OpenCV boundary and cv2 4.13 does not agree with
ITSELF to the 1e-3 px bar at these sizes: its SIMD/IPP build and its plain C++ build (cv2.setUseOptimized(False),
cv2.ipp.setUseIPP(False)) differ by up to 0.04 px in the border band of the 1080p pair and by up to 0.5 px inside the
4K/60-object pair (0.3 % of the pixels above 1e-3 px; profiles/r02_flow_scatter.json, tools/flow_scatter.py), wherever
the 2x2 system of FarnebackUpdateFlow is ill-conditioned or floor(x + dx) of FarnebackUpdateMatrices sits on a branch.
The same holds for any second implementation: oracle/farneback_np.py with OpenCV's own running box sums differs from cv2
by up to 7e-3 px on pixels where the two cv2 builds agree to 1e-4, and with direct window sums (what the CUDA kernel
computes; more accurate than running sums, but another rounding sequence) by up to 1.2e-2 px there — the very pixels and
values the CUDA path shows (profiles/r02_flow_scatter.json).  So the bar has three parts, all measured live on the same pair:
  * against the oracle with the kernel's summation order (farneback_np.farneback(sliding=False), itself pinned to cv2 by
    tests/test_flow_oracle.py; 1080p only, at 4K the numpy restatement needs too much memory and time): mean <= 1e-5 px
    and at least 99.9 % of ALL pixels within 1e-3 px.  (Measured on the B200: mean 2.3e-6, 99.88 % within 1e-4, max
    9e-3 — at the ill-conditioned pixels even two implementations of one summation order part ways, which is why no
    bar on the maximum can hold at this size; at 640x360 and 320x180 the strict bar max <= 1e-3 does hold and is
    asserted in tests/test_gpu_parity.py.)
  * against cv2 on STABLE pixels = no cv2-vs-cv2 difference above 1e-4 px anywhere in the (2*REACH+1)^2 neighbourhood
    (REACH = 16 px covers the 15-px window of the full-resolution iterations; >= 80 % of the frame): mean <= 1e-5 px and
    at most 0.1 % of them above 1e-3 px;
  * against cv2 everywhere: no further from cv2 than cv2's own second build is, distributionally: the fraction of
    pixels above 1e-3 / 1e-2 px is at most twice cv2's own fraction (+1e-4), and the 99.9th percentile at most 4x cv2's.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_EPE_MAX, TOL_EPE_MEAN = 1e-3, 1e-5
REACH = 16  # px


@pytest.fixture(scope="module")
def dofs():
    import denseopticalflowsegmentation3d_b200 as d
    return d


def epe(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d ** 2).sum(-1))


def cv2_flow(cv2, g0, g1, optimized):
    cv2.setUseOptimized(optimized)
    try:
        cv2.ipp.setUseIPP(optimized)
    except Exception:
        pass
    try:
        return cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    finally:
        cv2.setUseOptimized(True)
        try:
            cv2.ipp.setUseIPP(True)
        except Exception:
            pass


def record(name, payload):
    """Numbers of this run, kept beside the test log (gpurun_out/ travels back from the GPU box)."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, "r02_flow_parity.json")
        cur = json.load(open(p)) if os.path.exists(p) else {}
        cur[name] = payload
        json.dump(cur, open(p, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def synth_pair(W, H, n_objects):
    from denseopticalflowsegmentation3d_b200 import synth
    return synth.frames(1234, n_objects, 0, 2, W, H)


@pytest.mark.parametrize("W,H,n_objects", [(1920, 1080, 8), (3840, 2160, 60)])
def test_flow_full_size_vs_cv2(dofs, W, H, n_objects):
    cv2 = pytest.importorskip("cv2")
    fr = synth_pair(W, H, n_objects)
    g0, g1 = (cv2.cvtColor(x, cv2.COLOR_BGR2GRAY) for x in fr)
    ref = cv2_flow(cv2, g0, g1, True)
    ref_plain = cv2_flow(cv2, g0, g1, False)
    with dofs.Context(W, H, max_pairs=1) as c:
        gray = c.gray(fr)
        assert np.array_equal(gray[0], g0) and np.array_equal(gray[1], g1)
        f = c.flow(g0, g1)[0]
    e, e_cv = epe(f, ref), epe(ref_plain, ref)
    vs_oracle = None
    if W * H <= 1920 * 1080:
        from oracle import farneback_np
        e_or = epe(f, farneback_np.farneback(g0, g1, sliding=False))
        vs_oracle = {"max": float(e_or.max()), "mean": float(e_or.mean()), "frac_gt_1e-4": float((e_or > 1e-4).mean()),
                     "frac_gt_1e-3": float((e_or > 1e-3).mean())}
    unstable = cv2.dilate((e_cv > 1e-4).astype(np.uint8), np.ones((2 * REACH + 1, 2 * REACH + 1), np.uint8)) > 0
    stable = ~unstable
    q = lambda a, p: float(np.percentile(a, p))  # noqa: E731
    stats = {
        "frame": [W, H], "objects": n_objects, "stable_fraction": float(stable.mean()),
        "ours_vs_oracle_direct_sums": vs_oracle,
        "ours_vs_cv2": {"stable_max": float(e[stable].max()), "stable_mean": float(e[stable].mean()),
                        "stable_frac_gt_1e-3": float((e[stable] > 1e-3).mean()),
                        "all_max": float(e.max()), "all_mean": float(e.mean()), "p99.9": q(e, 99.9),
                        "frac_gt_1e-3": float((e > 1e-3).mean()), "frac_gt_1e-2": float((e > 1e-2).mean())},
        "cv2_plain_vs_cv2": {"stable_max": float(e_cv[stable].max()), "all_max": float(e_cv.max()),
                             "all_mean": float(e_cv.mean()), "p99.9": q(e_cv, 99.9),
                             "frac_gt_1e-3": float((e_cv > 1e-3).mean()), "frac_gt_1e-2": float((e_cv > 1e-2).mean())},
        "flow_magnitude_max": float(np.sqrt((ref.astype(np.float64) ** 2).sum(-1)).max()),
    }
    print(json.dumps(stats))
    record(f"{W}x{H}", stats)
    o, cv = stats["ours_vs_cv2"], stats["cv2_plain_vs_cv2"]
    assert stats["stable_fraction"] >= 0.80
    if vs_oracle:
        assert vs_oracle["mean"] <= TOL_EPE_MEAN and vs_oracle["frac_gt_1e-3"] <= 1e-3
    assert o["stable_mean"] <= TOL_EPE_MEAN and o["stable_frac_gt_1e-3"] <= 1e-3
    assert o["frac_gt_1e-3"] <= 2 * cv["frac_gt_1e-3"] + 1e-4
    assert o["frac_gt_1e-2"] <= 2 * cv["frac_gt_1e-2"] + 1e-4
    assert o["p99.9"] <= 4 * cv["p99.9"] + 1e-4


# ---------------------------------------------------------------------------------------------------
def compare_with_ref(boxes, psets, entries, W):
    """Against the unchanged reference: roots, pixel sets, sizes, classes bit-identical; the float outputs within the
    tolerances of tests/test_gpu_parity.py."""
    from test_gpu_parity import TOL_BEV, TOL_ERR, TOL_IMG, TOL_YAW
    assert [int(b["root"]) for b in boxes] == [e["root"] for e in entries]
    for b, px, e in zip(boxes, psets, entries):
        assert int(b["size"]) == e["size"] == len(px)
        assert np.array_equal(px, e["pixels"]), f"pixel set of root {e['root']}"
        sol = e["sol"]
        assert int(b["cls"]) == sol["cls"]
        assert abs(b["score"] - e["score"]) <= TOL_ERR and abs(b["move"] - e["move"]) <= 1e-12
        assert abs(b["w_error"] - sol["w_error"]) <= TOL_ERR and abs(b["h_error"] - sol["h_error"]) <= TOL_ERR
        assert abs(b["orient"] - sol["orient"]) <= TOL_YAW
        assert np.abs(b["ps_bev"] - sol["ps_bev"]).max() <= TOL_BEV
        assert np.abs(b["rectangle"] - sol["rectangle"]).max() <= TOL_BEV
        assert np.abs(b["lower_face"] - sol["lower_face"]).max() <= TOL_IMG
        assert np.abs(b["upper_face"] - sol["upper_face"]).max() <= TOL_IMG


def test_repo_pair_against_unchanged_reference_live(dofs, ref, golden_pair):
    """data/frame_1052-1053 (BASELINE config 1): the unchanged reference sources (oracle/_ref) run HERE, on the GPU box,
    on the same blurred-flow bits as the CUDA path."""
    from denseopticalflowsegmentation3d_b200.capi import box_pixel_sets
    fb = golden_pair["flow_blurred"]
    H, W = fb.shape[:2]
    persp, inv, up = ref.get_mats()
    res = ref.segment(fb, persp, inv, up)
    with dofs.Context(W, H) as c:
        out = c.segment(fb, already_blurred=True)
    boxes = out["boxes"][0]
    compare_with_ref(boxes, box_pixel_sets(out["labels"][0], boxes), res["entries"], W)
    assert len(boxes) == len(golden_pair["root"]) > 0


def test_1080p_pair_against_unchanged_reference_live(dofs, ref):
    """BASELINE config 2: whole path on the GPU (gray, Farneback, blur, graph, segmentation, lifting), then the
    unchanged reference (about half a minute on one core) on the GPU's own blurred flow."""
    from denseopticalflowsegmentation3d_b200.capi import box_pixel_sets
    W, H = 1920, 1080
    fr = synth_pair(W, H, 8)
    with dofs.Context(W, H, max_pairs=1) as c:
        whole = c.process(fr)
        gray = c.gray(fr)
        flow = c.flow(gray[:1], gray[1:])
        out = c.segment(flow, already_blurred=False, want_blurred=True)
    assert np.array_equal(whole["labels"], out["labels"]) and whole["boxes"][0].tobytes() == out["boxes"][0].tobytes()
    persp, inv, up = ref.get_mats()
    res = ref.segment(out["flow_blurred"][0], persp, inv, up)
    boxes = out["boxes"][0]
    compare_with_ref(boxes, box_pixel_sets(out["labels"][0], boxes), res["entries"], W)
    assert len(boxes) > 0
    print("1080p vs unchanged reference: %d segments identical (reference build_graph %.1f s, segment_graph %.1f s)"
          % (len(boxes), res["t_build"], res["t_segment"]))
