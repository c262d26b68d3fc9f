"""GPU tests of the round-2 boundary additions, through the C ABI: compact label formats, the streaming driver,
per-node accessors against the UNCHANGED reference, deferred errors of queued asynchronous calls, several contexts
with different Farneback windows in one process."""
import ctypes as C

import numpy as np
import pytest

from conftest import random_flow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dofs():
    import denseopticalflowsegmentation3d_b200 as d
    return d


def test_label_formats_agree(dofs, golden_synth):
    """u16 and run-length labels carry exactly the int32 label image."""
    from denseopticalflowsegmentation3d_b200 import capi
    fr = golden_synth["bgr"]
    n, H, W = fr.shape[0] - 1, fr.shape[1], fr.shape[2]
    with dofs.Context(W, H, max_pairs=n) as c:
        ref = c.process(fr)
        u16 = c.process_ex(fr, capi.LABELS_U16)
        rle = c.process_ex(fr, capi.LABELS_RLE, max_runs=4096)
        painted_rle, _ = c.paint(n, 0.5)   # the paint kernel reads whichever dense format the last call left
        again = c.process(fr)
        painted_i32, _ = c.paint(n, 0.5)
    assert np.array_equal(ref["labels"], again["labels"])
    assert np.array_equal(painted_rle, painted_i32)
    assert u16["labels"].dtype == np.uint16
    dense = u16["labels"].astype(np.int32)
    dense[u16["labels"] == 0xFFFF] = -1
    assert np.array_equal(dense, ref["labels"])
    assert (ref["labels"] >= 0).any()
    for i in range(n):
        assert 0 < rle["n_runs"][i] <= 4096
        runs = rle["labels"][i, :rle["n_runs"][i]]
        assert runs["start"][0] == 0 and np.all(np.diff(runs["start"].astype(np.int64)) > 0)
        assert np.all(runs["label"][1:] != runs["label"][:-1])      # maximal runs
        assert np.array_equal(capi.runs_to_labels(rle["labels"][i], rle["n_runs"][i], W * H), ref["labels"][i].reshape(-1))
        assert u16["boxes"][i].tobytes() == ref["boxes"][i].tobytes() == rle["boxes"][i].tobytes()
    bytes_rle = int(rle["n_runs"].sum()) * 8
    print("labels per pair: int32 %d B, u16 %d B, run-length %d B" % (W * H * 4, W * H * 2, bytes_rle // n))
    # too small a run capacity is an overflow, not a truncation
    with dofs.Context(W, H, max_pairs=n) as c:
        with pytest.raises(dofs.DofsError) as ei:
            c.process_ex(fr, capi.LABELS_RLE, max_runs=4)
        assert ei.value.status == -4
        ok = c.process_ex(fr, capi.LABELS_RLE, max_runs=4096)   # the context stays usable
        assert np.array_equal(ok["n_runs"], rle["n_runs"])


def test_segment_ex_formats_on_random_fields(dofs):
    from denseopticalflowsegmentation3d_b200 import capi
    W, H = 131, 77   # ragged: odd width, label rows not aligned
    fields = np.stack([random_flow(31 + k, W, H, scale=4.0) for k in range(3)])
    p = dofs.default_params()
    p.min_size = 40
    with dofs.Context(W, H, max_pairs=3, params=p) as c:
        ref = c.segment(fields, already_blurred=True)
        rle = c.segment_ex(fields, True, capi.LABELS_RLE, max_runs=W * H)
    for i in range(3):
        assert np.array_equal(capi.runs_to_labels(rle["labels"][i], rle["n_runs"][i], W * H), ref["labels"][i].reshape(-1))


@pytest.mark.parametrize("chunk", [32, 7])
def test_stream_equals_one_batch(dofs, chunk):
    """A 200-frame clip through dofs3d_stream_* in chunks equals one dofs3d_process call over the whole clip bit for bit
    (VERDICT r01 item 7): the carried frame's gray image and polynomial expansion stand in for re-expanding it."""
    from denseopticalflowsegmentation3d_b200 import capi, synth
    W, H, n = 192, 108, 199
    fr = synth.frames(9, 5, 0, n + 1, W, H)
    p = dofs.default_params()
    p.min_size = 120
    with dofs.Context(W, H, max_pairs=n, params=p) as c:
        whole = c.process_ex(fr, capi.LABELS_U16)
    with dofs.Context(W, H, max_pairs=chunk, params=p) as c:
        c.process_ex(fr[:2], capi.LABELS_RLE, max_runs=8192)   # the flow / run-length buffers exist from here on
        mem = c.device_bytes
        st = c.process_stream(fr, label_format=capi.LABELS_U16)
        st_rle = c.process_stream(fr, label_format=capi.LABELS_RLE, max_runs=8192)
        assert c.device_bytes - mem == 2 * (chunk + 1) * W * H * 3   # the two staging buffers: nothing grows with the clip
    assert len(st["boxes"]) == n and st["labels"].shape == whole["labels"].shape
    assert np.array_equal(st["labels"], whole["labels"])
    assert np.array_equal(st["n_boxes"], whole["n_boxes"]) and int(whole["n_boxes"].sum()) > 0
    for a, b in zip(st["boxes"], whole["boxes"]):
        assert a.tobytes() == b.tobytes()
    for k in ("n_merges", "n_candidates", "n_scored", "final_root"):
        assert np.array_equal(st["stats"][k], whole["stats"][k])
    for i in (0, chunk - 1, chunk, n - 1):
        dense = capi.runs_to_labels(st_rle["labels"][i], st_rle["n_runs"][i], W * H)
        want = whole["labels"][i].reshape(-1).astype(np.int32)
        want[want == 0xFFFF] = -1
        assert np.array_equal(dense, want)


def test_stream_argument_errors(dofs):
    from denseopticalflowsegmentation3d_b200 import synth
    W, H = 96, 64
    fr = synth.frames(3, 2, 0, 6, W, H)
    with dofs.Context(W, H, max_pairs=2) as c:
        c.stream_begin()
        with pytest.raises(dofs.DofsError):
            c.stream_submit(fr[:1])           # a stream cannot start with a single frame
        with pytest.raises(dofs.DofsError):
            c.stream_submit(fr[:4])           # 3 pairs > max_pairs
        c.stream_submit(fr[:3])
        c.stream_submit(fr[3:5])
        with pytest.raises(dofs.DofsError):
            c.stream_submit(fr[5:6])          # two chunks outstanding
        assert len(c.stream_collect()["boxes"]) == 2
        c.stream_submit(fr[5:6])
        assert len(c.stream_collect()["boxes"]) == 2
        assert len(c.stream_collect()["boxes"]) == 1
        with pytest.raises(dofs.DofsError):
            c.stream_collect()                # nothing outstanding


def test_node_accessors_against_unchanged_reference(dofs, ref, golden_pair):
    """Forest::get_segment_best_score (graph.cpp:386-389) = the score of the root's LATEST scored merge, before the
    convexity / threshold gates — from dofs3d_scored_merges, against the unchanged reference on the repo's pair."""
    fb = golden_pair["flow_blurred"]
    H, W = fb.shape[:2]
    persp, inv, up = ref.get_mats()
    res = ref.segment(fb, persp, inv, up, nodes=True)
    with dofs.Context(W, H) as c:
        out = c.segment(fb, already_blurred=True)
        last = c.last_scores(0)
        merges = c.scored_merges(0)
        final = c.node_state(0, int(out["stats"][0]["final_root"]))
    want = {int(i): float(res["node_score"][i]) for i in np.nonzero(res["node_score"])[0]}
    assert set(last) == set(want) and len(want) > 0
    for r, s in want.items():
        assert abs(last[r] - s) <= 1e-5
    assert merges["kept"].sum() == out["stats"][0]["n_scored"] and len(merges["root"]) >= merges["kept"].sum()
    # the only box the reference still holds at the end is the final root's (graph.cpp:207 clears the others)
    has_box = np.nonzero(res["node_bbox"][:, 0] >= 0)[0]
    assert list(has_box) == [int(out["stats"][0]["final_root"])]
    assert list(final["bbox"]) == list(res["node_bbox"][has_box[0]]) == [0, 0, W - 1, H - 1] and final["size"] == W * H


def test_node_state_against_port_trace(dofs, port):
    """dofs3d_node_state = the state a root had when it was absorbed = the port's merge trace at the root's last win."""
    W, H = 96, 64
    f = random_flow(77, W, H, scale=3.0)
    persp, inv, up = port.get_mats()
    res = port.segment(f, persp, inv, up, min_size=30, trace=True)
    tr = res["trace"]
    with dofs.Context(W, H) as c:
        c.segment(f, already_blurred=True)
        last_win = {}
        for k in range(len(tr["winner"])):
            last_win[int(tr["winner"][k])] = k
        rng = np.random.default_rng(0)
        for node in rng.choice(sorted(last_win), size=40, replace=False).tolist():
            k = last_win[node]
            st = c.node_state(0, node)
            assert st["size"] == tr["size"][k] and list(st["bbox"]) == list(tr["bbox"][k])
            assert np.array_equal(st["mean_flow"].view(np.uint32), tr["flow"][k].view(np.uint32))
        never = [p for p in range(W * H) if p not in last_win][:10]
        for node in never:
            st = c.node_state(0, node)
            assert st["size"] == 1 and list(st["bbox"]) == [node % W, node // W, node % W, node // W]


def test_queued_async_calls_report_any_failure(dofs):
    """Several *_dev calls before one dofs3d_sync: a failure of an EARLIER call is still reported (the
    context used to remember the last call only)."""
    import torch
    W, H = 64, 48
    bad = random_flow(5, W, H)
    bad[10:14, 20:24] = np.nan
    good = random_flow(6, W, H)
    with dofs.Context(W, H) as c:
        d_bad, d_good = torch.from_numpy(bad).cuda(), torch.from_numpy(good).cuda()
        stats = torch.zeros(40, dtype=torch.uint8, device="cuda")
        c.segment_dev(d_bad.data_ptr(), True, 1, d_stats=stats.data_ptr())
        c.segment_dev(d_good.data_ptr(), True, 1, d_stats=stats.data_ptr())
        with pytest.raises(dofs.DofsError) as ei:
            c.sync()
        assert ei.value.status == -5
        c.segment_dev(d_good.data_ptr(), True, 1, d_stats=stats.data_ptr())
        c.sync()   # reported once, then clear
    # max_boxes too small for an earlier call of the queue
    fields = random_flow(11, 160, 96, scale=4.0)
    p = dofs.default_params()
    p.min_size = 60
    with dofs.Context(160, 96, params=p) as c:
        d = torch.from_numpy(fields).cuda()
        boxes = torch.zeros(216, dtype=torch.uint8, device="cuda")
        nb = torch.zeros(1, dtype=torch.int32, device="cuda")
        c.segment_dev(d.data_ptr(), True, 1, d_boxes=boxes.data_ptr(), d_n_boxes=nb.data_ptr(), max_boxes=1)
        c.segment_dev(d.data_ptr(), True, 1)
        if int(nb.item()) > 1:
            with pytest.raises(dofs.DofsError) as ei:
                c.sync()
            assert ei.value.status == -4
        else:
            c.sync()


def test_contexts_with_different_windows_share_a_process(dofs):
    """the shared-memory opt-in of the box-filter kernels is set per context, not once per process:
    a later context with a larger Farneback window must work, and so must the first one afterwards."""
    cv2 = pytest.importorskip("cv2")
    from denseopticalflowsegmentation3d_b200 import synth
    W, H = 256, 144
    fr = synth.frames(5, 4, 0, 2, W, H)
    g = [cv2.cvtColor(x, cv2.COLOR_BGR2GRAY) for x in fr]

    def ctx(win):
        p = dofs.default_params()
        p.winsize = win
        return dofs.Context(W, H, params=p)

    with ctx(15) as a, ctx(31) as b, ctx(9) as s:
        for c, win in ((a, 15), (b, 31), (s, 9), (a, 15)):
            f = c.flow(g[0], g[1])[0]
            want = cv2.calcOpticalFlowFarneback(g[0], g[1], None, 0.5, 3, win, 3, 5, 1.2, 0)
            e = np.sqrt(((f.astype(np.float64) - want) ** 2).sum(-1))[16:-16, 16:-16]
            assert e.max() <= 1e-3, (win, e.max())


def test_negative_score_threshold_keeps_reference_semantics(dofs, port):
    """dofs3d_params.score_threshold < 0 is admissible: segment_history starts at -1 and keeps any score above the
    threshold (graph.cpp:348-352); the selection no longer uses 0 as its 'no score' value."""
    from test_gpu_parity import run_and_compare
    fields = [random_flow(61 + k, 160, 96, scale=4.0) for k in range(2)]
    assert run_and_compare(dofs, port, fields, min_size=60, score_threshold=-0.5) > 0


def test_kernel_variants_are_bit_identical(dofs, monkeypatch):
    """The A/B knobs select another staging of the same arithmetic: the shared-memory tiled pyramid levels against the
    per-thread kernel, the TMA (bulk copy) staged flow blur against LDG -> STS staging, the relabel pass folded into
    the Boruvka pixel kernel against the separate pass, the next iteration's update-matrices fused into the box filter +
    solve kernel against separate kernels; and the conditional graph nodes that skip the late Boruvka
    levels / replay waves against unconditional launches, with their IF bodies forced to run (two levels enqueued
    unconditionally, the rest behind the condition); the pyramid kernel with its filter radius at run time against the
    radius-templated instances.  Results must not change by a bit."""
    from denseopticalflowsegmentation3d_b200 import synth
    W, H, n = 328, 190, 2      # not a multiple of the tile sizes; W a multiple of 4 (word-aligned fast path taken)
    fr = synth.frames(21, 6, 0, n + 1, W, H)
    fr_odd = np.ascontiguousarray(fr[:, :, :-1])   # W = 327: every fast path falls back

    def run(frames):
        w = frames.shape[2]
        p = dofs.default_params()
        p.min_size = 150
        with dofs.Context(w, H, max_pairs=n, params=p) as c:
            g = c.gray(frames)
            f = c.flow(g[:-1], g[1:])
            out = c.segment(f, already_blurred=False, want_blurred=True)
        return f, out

    base = {k: run(v) for k, v in (("even", fr), ("odd", fr_odd))}
    for knob, val in (("DOFS3D_PYR_TILED", "1"), ("DOFS3D_BLUR_TMA", "0"), ("DOFS3D_BOR_FOLD", "1"), ("DOFS3D_STRIDE_BLOCKS", "5"),
                      ("DOFS3D_FLOW_FUSE", "1"), ("DOFS3D_COND_GRAPHS", "0"), ("DOFS3D_SOFT_LEVELS", "2"),
                      ("DOFS3D_SOFT_LEVELS", "1"), ("DOFS3D_PYR_GENERIC", "1")):
        monkeypatch.setenv(knob, val)
        for k, v in (("even", fr), ("odd", fr_odd)):
            f, out = run(v)
            assert np.array_equal(f, base[k][0]), (knob, k, "flow")
            assert np.array_equal(out["flow_blurred"], base[k][1]["flow_blurred"]), (knob, k, "blur")
            assert np.array_equal(out["labels"], base[k][1]["labels"]), (knob, k, "labels")
            for a, b in zip(out["boxes"], base[k][1]["boxes"]):
                assert a.tobytes() == b.tobytes()
        monkeypatch.delenv(knob)


@pytest.mark.parametrize("name", ["synth_crop", "blocks_8", "blocks_4", "ties", "thin"])
def test_fh_mode_equals_reference_python_golden(dofs, name):
    """the Felzenszwalb adaptive-threshold mode (graph.py:156-177) on the device against outputs of the
    reference's own Python (tests/golden/fh_cases.npz, tools/make_golden_fh.py): root id of every pixel identical, after
    the whole of segment_graph_flow and after its first two passes (= segment_graph)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fh_cases.npz"))
    f = g[name + "_flow"]
    K, ms, nb = g[name + "_params"]
    H, W = f.shape[:2]
    p = dofs.default_params()
    p.neighbors = int(nb)
    with dofs.Context(W, H, params=p) as c:
        lab3, n3 = c.segment_fh(f, K, int(ms), int(nb), stage=3)
        lab2, n2 = c.segment_fh(f, K, int(ms), int(nb), stage=2)
    assert np.array_equal(lab3, g[name + "_labels"]) and n3 == len(np.unique(lab3))
    assert np.array_equal(lab2, g[name + "_labels_stage2"]) and n2 == len(np.unique(lab2))


def test_fh_mode_equals_oracle_on_a_real_flow_field(dofs, golden_pair):
    """The repo's own frame pair (640x360, 918 602 edges) through the Felzenszwalb mode: device == CPU restatement."""
    from oracle import fh
    fb = np.ascontiguousarray(golden_pair["flow_blurred"]) * np.float32(3.0)
    H, W = fb.shape[:2]
    with dofs.Context(W, H) as c:
        c.set_timing(True)
        lab, n = c.segment_fh(fb, 10.0, 100, 8)
        t = c.timing()
    want, wn = fh.segment_flow(fb, 10.0, 100, 8)
    assert n == wn and np.array_equal(lab, want) and n > 1
    print("Felzenszwalb mode 640x360: %d components; device ms %s" % (n, {k: round(v[0], 2) for k, v in t.items()}))


def test_bev_warp_equals_cv2(dofs, golden_pair):
    """the reference's transform() = cv::warpPerspective(INTER_CUBIC, BORDER_REPLICATE) to 2500 x 14000,
    bit-exact against cv2 live, plus small ragged cases (1 and 3 channels, replicated borders)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    img = cv2.GaussianBlur(rng.integers(0, 256, (360, 640, 3), dtype=np.uint8), (0, 0), 2.0)
    with dofs.Context(640, 360) as c:
        persp = np.array(list(dofs.default_params().persp), np.float32).reshape(3, 3)
        c.set_timing(True)
        bev = c.bev_transform(img)
        ms = c.timing()["bev.warp"][0]
        ref = cv2.warpPerspective(img, persp, (2500, 14000), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
        assert bev.shape == ref.shape and np.array_equal(bev, ref)
        print("BEV warp 2500x14000x3: %.3f ms on the device = %.0f GB/s of output" % (ms, ref.size / ms / 1e6))
        src = np.float32([[50, 70], [20, 30], [75, 30], [150, 70]])
        dst = np.float32([[20, 300], [20, 40], [180, 40], [180, 300]])
        M = cv2.getPerspectiveTransform(src, dst).astype(np.float32)
        small = np.ascontiguousarray(img[:90, :160])
        for im, (ow, oh) in ((small, (220, 340)), (np.ascontiguousarray(small[..., 1]), (101, 77)), (small, (67, 33))):
            want = cv2.warpPerspective(im, M, (ow, oh), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
            assert np.array_equal(c.warp_perspective(im, M, ow, oh), want)


def test_render_equals_reference_display_with_cv2(dofs, golden_pair):
    """dofs3d_render = plot_best_segments_simple (draw.cpp:102-160) — painting in ascending root order,
    the wireframe cube of every drawn segment (draw_cube, draw.cpp:85-100), the final blend — against the same OpenCV
    calls (cv2.line, cv2.addWeighted) issued here in the reference's order, on the repo's own pair."""
    cv2 = pytest.importorskip("cv2")
    from denseopticalflowsegmentation3d_b200.capi import box_pixel_sets
    fb = golden_pair["flow_blurred"]
    H, W = fb.shape[:2]
    rng = np.random.default_rng(3)
    frame = cv2.GaussianBlur(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), (0, 0), 2.0)
    for min_score in (0.7, 0.3):
        with dofs.Context(W, H) as c:
            out = c.segment(fb, already_blurred=True)
            got = c.render(frame, min_score)[0]
        boxes = out["boxes"][0]
        psets = box_pixel_sets(out["labels"][0], boxes)
        fr, seg = frame.copy(), frame.copy()
        drawn = 0

        def pt(v):
            return (int(np.rint(np.float32(v[0]))), int(np.rint(np.float32(v[1]))))

        for b, px in zip(boxes, psets):           # ascending root = the order of the reference's history vector
            if not b["score"] > min_score:
                continue
            drawn += 1
            seg.reshape(-1, 3)[px] = (0, 255, 0) if b["cls"] == 1 else (0, 255, 255)
            for img in (fr, seg):
                lo, up = b["lower_face"], b["upper_face"]
                for i in range(4):
                    cv2.line(img, pt(lo[i]), pt(lo[(i + 1) % 4]), (255, 0, 0), 1)
                    cv2.line(img, pt(up[i]), pt(up[(i + 1) % 4]), (255, 0, 0), 1)
                    cv2.line(img, pt(lo[i]), pt(up[i]), (255, 0, 0), 1)
        want = cv2.addWeighted(fr, 1.0 - 2.0 / 5.0, seg, 2.0 / 5.0, 0)
        assert drawn > 0 and np.array_equal(got, want), (min_score, int((got != want).sum()))


def test_repeated_runs_are_deterministic_on_tie_heavy_fields(dofs, port):
    """compute-sanitizer is closed on this pool (profiles/r02_sanitizer_closed.txt), so the lock-free protocols of the
    segmentation (exact minimum under concurrent offers in the Boruvka pixel kernel, path halving with benign races in
    the contraction, decoupled look-back in the one-sweep sort) are exercised the way a race would show: fields made of
    ties and near-ties, where every tie-break path runs thousands of times, repeated in one context and across fresh
    contexts — every repetition must reproduce the first one bit for bit, and the first one equals the oracle."""
    from test_gpu_parity import near_tie_field, compare_boxes
    from denseopticalflowsegmentation3d_b200.capi import box_pixel_sets
    W, H = 320, 200
    zero = np.zeros((H, W, 2), np.float32)
    blocks = np.zeros((H, W, 2), np.float32)
    blocks[40:160, 60:260] = (1.5, -2.0)                  # two plateaus: all ties inside, one ring of equal weights
    rounded = np.round(random_flow(91, W, H, scale=3.0) * 2) / 2   # a handful of distinct weights, huge tie classes
    fields = np.stack([zero, blocks, near_tie_field(W, H), rounded, random_flow(92, W, H, scale=4.0)])
    p = dofs.default_params()
    p.min_size = 200
    first = None
    for rep in range(4):
        with dofs.Context(W, H, max_pairs=len(fields), params=p) as c:
            for _ in range(4):
                out = c.segment(fields, already_blurred=True)
                sig = (out["labels"].tobytes(), b"".join(b.tobytes() for b in out["boxes"]),
                       out["stats"][["n_merges", "n_candidates", "n_scored", "n_boxes", "final_root"]].tobytes())
                if first is None:
                    first = (sig, out)
                assert sig == first[0], "a repetition differs from the first run"
    persp, inv, up = port.get_mats()
    out = first[1]
    for i in range(len(fields)):
        res = port.segment(fields[i], persp, inv, up, min_size=200)
        compare_boxes(out["boxes"][i], box_pixel_sets(out["labels"][i], out["boxes"][i]), res["entries"], W)
        assert out["stats"][i]["n_candidates"] == res["counters"]["get_score"]


def test_incremental_forest_drives_the_reference_loop(dofs, port):
    """Forest::Forest / find / merge / new_merge one call at a time (graph.hpp:72-114) over device state: the loop of the
    reference's segment_graph (graph.cpp:520-531) written against dofs3d_forest_*, on the port's sorted edge list, ends
    with the history the oracle's whole pass ends with — roots, pixel sets, sizes, mean-flow bits, scores."""
    from denseopticalflowsegmentation3d_b200.capi import Forest
    W, H = 56, 40
    f = random_flow(17, W, H, scale=4.0)
    persp, inv, up = port.get_mats()
    res = port.segment(f, persp, inv, up, min_size=40, trace=True)
    s, e, _ = port.build_graph(f, True)
    with dofs.Context(W, H) as c, Forest(c, f) as forest:
        assert forest.num_sets == W * H and forest.find(5) == 5 and list(forest.bbox(5)) == [5, 0, 5, 0]
        merges = 0
        for a0, b0 in zip(s.tolist(), e.tolist()):
            a, b = forest.find(a0), forest.find(b0)
            if a != b:
                forest.new_merge(a, b, 0.3, 40)
                assert forest.find(a0) == forest.find(b0) == int(res["trace"]["winner"][merges])
                merges += 1
        assert merges == W * H - 1 and forest.num_sets == 1
        boxes = forest.boxes()
        ents = res["entries"]
        assert [int(b["root"]) for b in boxes] == [en["root"] for en in ents] and len(ents) > 0
        for b, en in zip(boxes, ents):
            assert int(b["size"]) == en["size"] and abs(b["score"] - en["score"]) <= 1e-5 and int(b["time"]) == en["time"]
            assert np.array_equal(np.asarray(b["mean_flow"]).view(np.uint32), en["flow"].view(np.uint32))
            assert np.array_equal(forest.pixels(int(b["root"]), int(b["size"])), en["pixels"])
        final = forest.find(0)
        assert list(forest.bbox(final)) == [0, 0, W - 1, H - 1]
        absorbed = next(p for p in range(W * H) if p != final)
        assert forest.bbox(absorbed) is None          # Forest::merge clears the box of an absorbed root (graph.cpp:207)
    # plain merge (no gates, no history) keeps the union-find bookkeeping of graph.cpp:170-218
    with dofs.Context(W, H) as c, Forest(c, f) as forest:
        assert forest.merge(0, 1) == 1 and forest.num_sets == W * H - 1      # equal ranks: b's root survives
        assert forest.merge(2, 0) == 1 and forest.find(2) == 1                 # lower rank hangs under the higher
        assert len(forest.boxes()) == 0
