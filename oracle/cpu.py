"""TEST INFRASTRUCTURE (oracle) — not part of the product.

ctypes bindings of the two CPU checkers, which export the same C functions:
  * ref()  = oracle/_ref/libdofs3d_ref.so      the reference's own, unchanged cpp/src/graph.cpp +
             cpp/src/lifting_3d.cpp (oracle/ref_driver.cpp, `make -C oracle ref`; needs
             /root/reference at build time, the built .so travels to the GPU box)
  * port() = oracle/_build/libdofs3d_oracle.so the CPU restatement oracle/dofs3d_oracle.cpp
             (`make -C oracle port`)
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libdofs3d_ref.so")
PORT_PATH = os.path.join(_HERE, "_build", "libdofs3d_oracle.so")

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)
_dp = C.POINTER(C.c_double)


class Solution(C.Structure):
    _fields_ = [
        ("cls", C.c_int),
        ("has_rectangle", C.c_int),
        ("ps_bev", C.c_float * 8),
        ("lower_face", C.c_float * 8),
        ("upper_face", C.c_float * 8),
        ("rectangle", C.c_float * 8),
        ("w_error", C.c_double),
        ("h_error", C.c_double),
        ("orient", C.c_double),
    ]

    def as_dict(self):
        d = {"cls": self.cls, "has_rectangle": bool(self.has_rectangle),
             "w_error": self.w_error, "h_error": self.h_error, "orient": self.orient}
        for k in ("ps_bev", "lower_face", "upper_face", "rectangle"):
            d[k] = np.array(getattr(self, k), dtype=np.float32).reshape(4, 2)
        return d


def _f(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
    if n is not None:
        assert a.size == n, (a.size, n)
    return a, a.ctypes.data_as(_fp)


class CpuOracle:
    def __init__(self, path, kind):
        self.path, self.kind = path, kind
        L = C.CDLL(path)
        L.ref_get_mats.argtypes = [_fp, _fp, _fp]
        L.ref_get_intersect.argtypes = [_fp, _fp, _fp, _fp, _fp]
        L.ref_get_bottom_variants.argtypes = [_fp, _ip, _fp, _fp, _fp, C.c_int, C.POINTER(Solution)]
        L.ref_build_graph.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _ip, _ip, _dp]
        L.ref_build_graph.restype = C.c_long
        L.ref_segment.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp]
        L.ref_segment.restype = C.c_void_p
        L.ref_result_count.argtypes = [C.c_void_p]
        L.ref_result_edges.argtypes = [C.c_void_p]
        L.ref_result_edges.restype = C.c_long
        L.ref_result_num_sets.argtypes = [C.c_void_p]
        L.ref_result_times.argtypes = [C.c_void_p, _dp, _dp]
        L.ref_result_entry.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, C.POINTER(Solution)]
        L.ref_result_pixels.argtypes = [C.c_void_p, C.c_int, _ip]
        L.ref_result_free.argtypes = [C.c_void_p]
        if kind == "ref":
            L.ref_result_nodes.argtypes = [C.c_void_p, _dp, _ip]
        L.ref_set_counting.argtypes = [C.c_int]
        L.ref_get_counts.argtypes = [C.c_char_p, C.c_int]
        if kind == "port":
            L.oracle_segment_ex.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_double, C.c_int, C.c_int]
            L.oracle_segment_ex.restype = C.c_void_p
            L.oracle_result_entry_extra.argtypes = [C.c_void_p, C.c_int, _ip, _fp, _ip]
            L.oracle_result_counters.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
            L.oracle_result_trace.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _ip, _fp]
            L.oracle_result_trace.restype = C.c_long
        self.L = L

    def get_mats(self):
        """(persp, inv, upper[3]) float32 3x3 — get_mat / get_mat_upper (lifting_3d.cpp:441-514)."""
        persp, inv, upper = np.zeros(9, np.float32), np.zeros(9, np.float32), np.zeros(27, np.float32)
        self.L.ref_get_mats(persp.ctypes.data_as(_fp), inv.ctypes.data_as(_fp), upper.ctypes.data_as(_fp))
        return persp.reshape(3, 3), inv.reshape(3, 3), upper.reshape(3, 3, 3)

    def get_intersect(self, a, b, c, d):
        """get_intersect (lifting_3d.cpp:63-88)."""
        out = np.zeros(2, np.float32)
        (_, pa), (_, pb), (_, pc), (_, pd) = _f(a, 2), _f(b, 2), _f(c, 2), _f(d, 2)
        self.L.ref_get_intersect(pa, pb, pc, pd, out.ctypes.data_as(_fp))
        return out

    def get_bottom_variants(self, direction, box, mat, inv_mat, inv_upper, cls):
        """get_bottom_variants (lifting_3d.cpp:350-439); box = (xmin, ymin, xmax, ymax)."""
        d, pd = _f(direction, 2)
        m, pm = _f(mat, 9)
        im, pim = _f(inv_mat, 9)
        u, pu = _f(inv_upper, 9)
        b = np.ascontiguousarray(box, dtype=np.int32)
        sol = Solution()
        self.L.ref_get_bottom_variants(pd, b.ctypes.data_as(_ip), pm, pim, pu, int(cls), C.byref(sol))
        return sol.as_dict()

    def build_graph(self, flow, neighbors8=True):
        """Sorted edge list of build_graph (graph.cpp:51-103): (start, end, weight)."""
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        h, w = flow.shape[:2]
        n = 4 * w * h
        start, end, weight = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float64)
        e = self.L.ref_build_graph(flow.ctypes.data_as(_fp), w, h, 1 if neighbors8 else 0,
                                   start.ctypes.data_as(_ip), end.ctypes.data_as(_ip), weight.ctypes.data_as(_dp))
        return start[:e].copy(), end[:e].copy(), weight[:e].copy()

    def segment(self, flow_blurred, persp, inv, upper, neighbors=8, score_threshold=0.3, min_size=500, trace=False, nodes=False):
        """build_graph + segment_graph + get_best_segments on an already blurred flow field
        (segment.cpp:54-63).  Returns dict(entries=[{root, score, move, size, pixels, sol, ...}], ...)."""
        flow = np.ascontiguousarray(flow_blurred, dtype=np.float32)
        h, w = flow.shape[:2]
        p, pp = _f(persp, 9)
        i, pi = _f(inv, 9)
        u, pu = _f(upper, 27)
        L = self.L
        if self.kind == "port":
            hnd = L.oracle_segment_ex(flow.ctypes.data_as(_fp), w, h, neighbors, pp, pi, pu,
                                      score_threshold, min_size, 1 if trace else 0)
        else:
            assert score_threshold == 0.3 and min_size == 500 and not trace, "fixed in the reference (graph.hpp:93-94)"
            hnd = L.ref_segment(flow.ctypes.data_as(_fp), w, h, neighbors, pp, pi, pu)
        try:
            out = {"entries": [], "n_edges": L.ref_result_edges(hnd), "num_sets": L.ref_result_num_sets(hnd)}
            tb, ts = C.c_double(), C.c_double()
            L.ref_result_times(hnd, C.byref(tb), C.byref(ts))
            out["t_build"], out["t_segment"] = tb.value, ts.value
            for k in range(L.ref_result_count(hnd)):
                root, score, move = C.c_int32(), C.c_double(), C.c_double()
                sol = Solution()
                size = L.ref_result_entry(hnd, k, C.byref(root), C.byref(score), C.byref(move), C.byref(sol))
                px = np.zeros(size, np.int32)
                L.ref_result_pixels(hnd, k, px.ctypes.data_as(_ip))
                ent = {"root": root.value, "score": score.value, "move": move.value,
                       "size": size, "pixels": px, "sol": sol.as_dict()}
                if self.kind == "port":
                    t = C.c_int32()
                    fl, bb = np.zeros(2, np.float32), np.zeros(4, np.int32)
                    L.oracle_result_entry_extra(hnd, k, C.byref(t), fl.ctypes.data_as(_fp), bb.ctypes.data_as(_ip))
                    ent.update(time=t.value, flow=fl, bbox=bb)
                out["entries"].append(ent)
            if self.kind == "ref" and nodes:
                # Forest::get_segment_best_score / get_bounding_box of every node of the finished forest
                sc, bb = np.zeros(h * w, np.float64), np.zeros((h * w, 4), np.int32)
                L.ref_result_nodes(hnd, sc.ctypes.data_as(_dp), bb.ctypes.data_as(_ip))
                out["node_score"], out["node_bbox"] = sc, bb
            if self.kind == "port":
                cnt = (C.c_long * 9)()
                L.oracle_result_counters(hnd, cnt)
                names = ["merges", "fail_size", "fail_row", "fail_move", "get_score", "fail_no_rect",
                         "fail_convexity", "fail_score", "history_writes"]
                out["counters"] = dict(zip(names, list(cnt)))
                if trace:
                    n = L.oracle_result_trace(hnd, None, None, None, None, None, None)
                    tr = {k: np.zeros(n, np.int32) for k in ("loser", "winner", "size", "edge_pos")}
                    tr["bbox"] = np.zeros((n, 4), np.int32)
                    tr["flow"] = np.zeros((n, 2), np.float32)
                    L.oracle_result_trace(hnd, *(tr[k].ctypes.data_as(_ip) for k in ("loser", "winner", "size", "edge_pos", "bbox")),
                                          tr["flow"].ctypes.data_as(_fp))
                    out["trace"] = tr
        finally:
            L.ref_result_free(hnd)
        return out

    def set_counting(self, on):
        self.L.ref_set_counting(1 if on else 0)

    def get_counts(self):
        n = self.L.ref_get_counts(None, 0)
        buf = C.create_string_buffer(n + 16)
        self.L.ref_get_counts(buf, n + 16)
        out = {}
        for line in buf.value.decode().splitlines():
            if "=" in line:
                k, v = line.rsplit("=", 1)
                out[k] = int(v)
        return out


_cache = {}


def build(ref_too=True):
    """Compile the checkers (the port always; _ref only where /root/reference exists)."""
    subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)
    if ref_too and os.path.isdir("/root/reference/cpp/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def ref_available():
    return os.path.exists(REF_PATH)


def ref():
    if "ref" not in _cache:
        if not ref_available():
            raise RuntimeError(f"{REF_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        _cache["ref"] = CpuOracle(REF_PATH, "ref")
    return _cache["ref"]


def port():
    if "port" not in _cache:
        if not os.path.exists(PORT_PATH):
            build(ref_too=False)
        _cache["port"] = CpuOracle(PORT_PATH, "port")
    return _cache["port"]
