"""CPU tests (no GPU): pin the oracle.

* the reference's own known-answer tests (cpp/tests/test_liftig_3d.cpp:69-89,179-227) against the
  CPU restatement (port) and, where built, against the unchanged reference sources (oracle/_ref);
* the committed golden fixtures (outputs of the unchanged reference, tools/make_golden.py) against
  the port;
* port == _ref on seeded random fields.
"""
import hashlib

import numpy as np
import pytest

from conftest import random_flow


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


# ---- the reference's gtest cases ---------------------------------------------------------------
def check_intersection_exists(o):  # IntersectionTest.IntersectionExists, test_liftig_3d.cpp:69-78
    p = o.get_intersect((1, 1), (4, 4), (1, 8), (2, 4))
    assert abs(p[0] - 2.4) < 1e-2 and abs(p[1] - 2.4) < 1e-2


def check_intersection_missing(o):  # IntersectionTest.IntersectionDoesNotExist, :80-89
    p = o.get_intersect((1, 1), (1, 2), (3, 3), (3, 4))
    assert np.isnan(p[0]) and np.isnan(p[1])


# GetBottomVariantsTest.Test1 (test_liftig_3d.cpp:179-227): inputs :181-186, expected :190-205
KAT = dict(
    dir=(2.5470946, 1.9316475), box=(375, 92, 576, 286), cls=2,
    mat=[20.1377838, -13.4744920, 402.174272, 5.11635077, 800.335022, -62251.3321, 0.000393565444, 0.0397205947, 1.0],
    inv_mat=[0.202212552, 0.00181942728, 31.9370859, -0.00182975914, 0.00123437589, 77.5774258, -6.90475148e-06,
             -4.97462083e-05, 1.0],
    inv_upper=[0.203701900, 0.00169508037, 32.3672674, 0.0, 0.00146371164, 29.6614710, 0.0, -5.01704822e-05, 1.0],
    ps_bev=[(327.809749190909, 13476.772230116465), (1398.2414179174136, 2769.3562851313454),
            (2204.8576245955073, 2935.1653236816246), (647.3324083001849, 13473.77576520306)],
    lower_face=[(385.305, 286.0), (375.0, 269.47327), (555.45557, 270.2111), (576.0, 286.92706)],
    upper_face=[(385.305, 99.43571), (375.0, 92.75792), (555.45557, 92.0), (576.0, 98.69487)],
    w_error=0.5987518562843858, h_error=0.7156805292391223, orient=-1.6261444189491607)


def check_bottom_variants_kat(o, tol=1e-1):  # tolerance 1e-1 as in the reference's own test (:213-226)
    s = o.get_bottom_variants(KAT["dir"], KAT["box"], KAT["mat"], KAT["inv_mat"], KAT["inv_upper"], KAT["cls"])
    np.testing.assert_allclose(s["ps_bev"], KAT["ps_bev"], atol=tol, rtol=0)
    np.testing.assert_allclose(s["lower_face"], KAT["lower_face"], atol=tol, rtol=0)
    np.testing.assert_allclose(s["upper_face"], KAT["upper_face"], atol=tol, rtol=0)
    assert abs(s["w_error"] - KAT["w_error"]) < tol
    assert abs(s["h_error"] - KAT["h_error"]) < tol
    assert abs(s["orient"] - KAT["orient"]) < tol
    return s


def test_kat_port(port):
    check_intersection_exists(port)
    check_intersection_missing(port)
    check_bottom_variants_kat(port)


def test_kat_ref(ref):
    check_intersection_exists(ref)
    check_intersection_missing(ref)
    check_bottom_variants_kat(ref)


def test_kat_port_equals_ref_bitwise(port, ref):
    a, b = check_bottom_variants_kat(port), check_bottom_variants_kat(ref)
    for k in ("ps_bev", "lower_face", "upper_face", "rectangle"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("w_error", "h_error", "orient"):
        assert a[k] == b[k], k


# ---- golden fixtures (unchanged reference outputs) vs the port ----------------------------------
def entries_vs_golden(res, g, exact=True):
    ents = res["entries"]
    assert [e["root"] for e in ents] == list(g["root"])
    assert [e["size"] for e in ents] == list(g["size"])
    assert [e["sol"]["cls"] for e in ents] == list(g["cls"])
    cmp = (lambda a, b: np.array_equal(np.asarray(a), np.asarray(b))) if exact else \
        (lambda a, b: np.allclose(a, b, rtol=1e-12, atol=1e-12))
    assert cmp([e["score"] for e in ents], g["score"])
    assert cmp([e["move"] for e in ents], g["move"])
    assert cmp([e["sol"]["orient"] for e in ents], g["orient"])
    assert cmp([e["sol"]["w_error"] for e in ents], g["w_error"])
    assert cmp([e["sol"]["h_error"] for e in ents], g["h_error"])
    for k in ("ps_bev", "rectangle", "lower_face", "upper_face"):
        assert np.array_equal(np.array([e["sol"][k] for e in ents], np.float32).reshape(-1, 4, 2), g[k]), k
    off = g["pixel_offsets"]
    for i, e in enumerate(ents):
        assert np.array_equal(e["pixels"], g["pixels"][off[i]:off[i + 1]]), f"pixel set of entry {i}"


@pytest.mark.parametrize("name", ["golden_pair", "golden_synth"])
def test_port_matches_reference_golden(port, name, request):
    g = request.getfixturevalue(name)
    persp, inv, up = port.get_mats()
    fb = g["flow_blurred"]
    s, e, w = port.build_graph(fb)
    assert len(s) == int(g["n_edges"])
    assert sha(s) == str(g["edges_sha_start"]) and sha(e) == str(g["edges_sha_end"]) and sha(w) == str(g["edges_sha_weight"])
    res = port.segment(fb, persp, inv, up)
    assert res["num_sets"] == int(g["num_sets"]) == 1
    entries_vs_golden(res, g)
    # gate counters of the unchanged reference (its log lines, counted)
    cnt = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counter_values"]]))
    c = res["counters"]
    assert c["merges"] == cnt["new_merge"]
    assert c["fail_size"] == cnt["Low size"]
    assert c["fail_row"] == cnt.get("Low y", 0)
    assert c["fail_move"] == cnt["Low movement"]
    assert c["get_score"] == cnt["get_score"]
    assert c["fail_convexity"] == cnt["Low convexity"]
    assert c["fail_score"] == cnt.get("Low score", 0)
    assert c["history_writes"] == cnt["Add segment with score"]


def test_port_matches_reference_lifting_golden(port, golden_lift):
    g = golden_lift
    for i in range(len(g["cls"])):
        c = int(g["cls"][i])
        s = port.get_bottom_variants(g["dir"][i], g["box"][i], g["persp"], g["inv"], g["upper"][c], c)
        assert s["has_rectangle"] == bool(g["has_rect"][i])
        for k in ("w_error", "h_error", "orient"):
            assert s[k] == g[k][i] or (np.isnan(s[k]) and np.isnan(g[k][i])), (i, k)
        for k in ("ps_bev", "rectangle", "lower_face", "upper_face"):
            assert np.array_equal(s[k], g[k][i], equal_nan=True), (i, k)


# ---- port == unchanged reference on seeded random fields ------------------------------------------
@pytest.mark.parametrize("seed,W,H,n8", [(1, 96, 64, True), (2, 128, 80, False), (3, 200, 120, True)])
def test_port_equals_ref_random(port, ref, seed, W, H, n8):
    fb = random_flow(seed, W, H)
    a, b = port.build_graph(fb, n8), ref.build_graph(fb, n8)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    persp, inv, up = ref.get_mats()
    ra = port.segment(fb, persp, inv, up, neighbors=8 if n8 else 4)
    rb = ref.segment(fb, persp, inv, up, neighbors=8 if n8 else 4)
    assert len(ra["entries"]) == len(rb["entries"])
    for ea, eb in zip(ra["entries"], rb["entries"]):
        assert ea["root"] == eb["root"] and ea["score"] == eb["score"] and ea["move"] == eb["move"]
        assert np.array_equal(ea["pixels"], eb["pixels"])
        for k in ("ps_bev", "rectangle", "lower_face", "upper_face"):
            assert np.array_equal(ea["sol"][k], eb["sol"][k])


def test_get_mats_port_equals_ref(port, ref):
    for a, b in zip(port.get_mats(), ref.get_mats()):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------------
# The structural fact the CUDA segmentation rests on (csrc/dofs_seg.cuh K8): the merge loop only
# acts on the edges it accepts, so ordering THOSE edges — on one 32-bit key (the slot for a zero weight, else an
# order-preserving 32-bit prefix of the weight) with ties of a prefix broken by (weight, slot) — reproduces the order in
# which the reference accepts them.  Checked here on the CPU against the port's merge trace.
def _edge_prefix(w):
    """csrc/dofs_seg.cuh edge_prefix: exponent rebased to 2^-200, top 23 mantissa bits."""
    bits = np.ascontiguousarray(w, np.float64).view(np.uint64)
    base = np.uint64((1023 - 200) << 52)
    k = np.where(bits > base, bits - base, np.uint64(0))
    return np.minimum(k >> np.uint64(29), np.uint64(0xFFFFFFFE)).astype(np.uint64)


def _near_tie_columns(W, H, seed=0):
    rng = np.random.default_rng(seed)
    f = np.zeros((H, W, 2), np.float32)
    f[..., 1] = np.arange(H, dtype=np.float32)[:, None] + 1000.0 * np.arange(W, dtype=np.float32)[None, :]
    f[..., 0] = (rng.random((H, W)) * 3e-4).astype(np.float32)
    return f


@pytest.mark.parametrize("field", ["random", "zero", "near_tie"])
def test_merge_times_from_accepted_edges_only(port, field):
    W, H = 64, 40
    if field == "random":
        fb = random_flow(7, W, H)  # smooth field with exactly-flat patches (zero-weight ties)
    elif field == "zero":
        fb = np.zeros((H, W, 2), np.float32)
    else:
        fb = _near_tie_columns(W, H)  # thousands of accepted edges of weight 1 + O(1e-8): one prefix, different weights
    persp, inv, up = port.get_mats()
    start, end, w = port.build_graph(fb, True)
    tr = port.segment(fb, persp, inv, up, trace=True)["trace"]
    pos = tr["edge_pos"].astype(np.int64)  # positions of the accepted edges in the reference's sorted list, in merge order
    assert len(pos) == W * H - 1 and np.all(np.diff(pos) > 0)
    s, e, wa = start[pos].astype(np.int64), end[pos].astype(np.int64), w[pos]
    # slot = 4 * start + direction (0 left, 1 up, 2 up-left, 3 down-left) = the reference's insertion sequence number
    d = np.select([e == s - 1, e == s - W, e == s - W - 1, e == s + W - 1], [0, 1, 2, 3], -1)
    assert np.all(d >= 0)
    slot = 4 * s + d
    prefix = _edge_prefix(wa)
    assert np.all((prefix == 0) == (wa == 0)) and np.all(prefix[wa > 0] >= 0x19800000) and 4 * W * H < 0x19800000
    key = np.where(prefix == 0, slot.astype(np.uint64), prefix)
    # order by (key, weight, slot) — what the 4-pass radix sort + the repair of equal-prefix runs produce
    order = np.lexsort((slot, wa.view(np.uint64), key))
    assert np.array_equal(order, np.arange(len(pos))), "merge times from the accepted edges alone differ from the reference order"
    if field == "near_tie":
        runs = np.unique(key, return_counts=True)[1]
        assert runs.max() > 2048  # the case the exact fallback sort exists for
