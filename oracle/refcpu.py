"""TEST INFRASTRUCTURE (oracle) — not part of the product.

ctypes binding of oracle/_ref/libdofs3d_ref.so: the reference's own, unchanged
cpp/src/graph.cpp + cpp/src/lifting_3d.cpp (see oracle/ref_driver.cpp, oracle/Makefile).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdofs3d_ref.so")


class RefSolution(C.Structure):
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
        return {
            "cls": self.cls,
            "has_rectangle": bool(self.has_rectangle),
            "ps_bev": np.array(self.ps_bev, dtype=np.float32).reshape(4, 2),
            "lower_face": np.array(self.lower_face, dtype=np.float32).reshape(4, 2),
            "upper_face": np.array(self.upper_face, dtype=np.float32).reshape(4, 2),
            "rectangle": np.array(self.rectangle, dtype=np.float32).reshape(4, 2),
            "w_error": self.w_error,
            "h_error": self.h_error,
            "orient": self.orient,
        }


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int32)
        dp = C.POINTER(C.c_double)
        L.ref_get_mats.argtypes = [fp, fp, fp]
        L.ref_get_intersect.argtypes = [fp, fp, fp, fp, fp]
        L.ref_get_bottom_variants.argtypes = [fp, ip, fp, fp, fp, C.c_int, C.POINTER(RefSolution)]
        L.ref_build_graph.argtypes = [fp, C.c_int, C.c_int, C.c_int, ip, ip, dp]
        L.ref_build_graph.restype = C.c_long
        L.ref_segment.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, fp, fp]
        L.ref_segment.restype = C.c_void_p
        L.ref_result_count.argtypes = [C.c_void_p]
        L.ref_result_edges.argtypes = [C.c_void_p]
        L.ref_result_edges.restype = C.c_long
        L.ref_result_num_sets.argtypes = [C.c_void_p]
        L.ref_result_times.argtypes = [C.c_void_p, dp, dp]
        L.ref_result_entry.argtypes = [C.c_void_p, C.c_int, ip, dp, dp, C.POINTER(RefSolution)]
        L.ref_result_pixels.argtypes = [C.c_void_p, C.c_int, ip]
        L.ref_result_free.argtypes = [C.c_void_p]
        L.ref_set_counting.argtypes = [C.c_int]
        L.ref_get_counts.argtypes = [C.c_char_p, C.c_int]
        _lib = L
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def get_mats():
    """(persp, inv, upper[3]) as float32 3x3 — get_mat / get_mat_upper (lifting_3d.cpp:441-514)."""
    persp = np.zeros(9, np.float32)
    inv = np.zeros(9, np.float32)
    upper = np.zeros(27, np.float32)
    fp = C.POINTER(C.c_float)
    lib().ref_get_mats(persp.ctypes.data_as(fp), inv.ctypes.data_as(fp), upper.ctypes.data_as(fp))
    return persp.reshape(3, 3), inv.reshape(3, 3), upper.reshape(3, 3, 3)


def get_intersect(a, b, c, d):
    out = np.zeros(2, np.float32)
    aa, pa = _f(a)
    bb, pb = _f(b)
    cc, pc = _f(c)
    dd, pd = _f(d)
    lib().ref_get_intersect(pa, pb, pc, pd, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out


def get_bottom_variants(direction, box, mat, inv_mat, inv_upper, cls):
    d, pd = _f(direction)
    m, pm = _f(np.asarray(mat).reshape(9))
    im, pim = _f(np.asarray(inv_mat).reshape(9))
    u, pu = _f(np.asarray(inv_upper).reshape(9))
    b = np.ascontiguousarray(box, dtype=np.int32)
    sol = RefSolution()
    lib().ref_get_bottom_variants(pd, b.ctypes.data_as(C.POINTER(C.c_int32)), pm, pim, pu, int(cls), C.byref(sol))
    return sol.as_dict()


def build_graph(flow, neighbors8=True):
    """Sorted edge list of build_graph (graph.cpp:51-103): (start, end, weight)."""
    flow = np.ascontiguousarray(flow, dtype=np.float32)
    h, w = flow.shape[:2]
    n = 4 * w * h
    start = np.zeros(n, np.int32)
    end = np.zeros(n, np.int32)
    weight = np.zeros(n, np.float64)
    ip = C.POINTER(C.c_int32)
    e = lib().ref_build_graph(
        flow.ctypes.data_as(C.POINTER(C.c_float)), w, h, 1 if neighbors8 else 0,
        start.ctypes.data_as(ip), end.ctypes.data_as(ip), weight.ctypes.data_as(C.POINTER(C.c_double)))
    return start[:e].copy(), end[:e].copy(), weight[:e].copy()


def segment(flow_blurred, persp, inv, upper, neighbors=8):
    """build_graph + segment_graph + get_best_segments on an already blurred flow field.

    Returns dict(entries=[{root, score, move, size, pixels, sol}], n_edges, num_sets, t_build, t_segment).
    """
    flow = np.ascontiguousarray(flow_blurred, dtype=np.float32)
    h, w = flow.shape[:2]
    p, pp = _f(np.asarray(persp).reshape(9))
    i, pi = _f(np.asarray(inv).reshape(9))
    u, pu = _f(np.asarray(upper).reshape(27))
    L = lib()
    hnd = L.ref_segment(flow.ctypes.data_as(C.POINTER(C.c_float)), w, h, neighbors, pp, pi, pu)
    try:
        out = {"entries": [], "n_edges": L.ref_result_edges(hnd), "num_sets": L.ref_result_num_sets(hnd)}
        tb, ts = C.c_double(), C.c_double()
        L.ref_result_times(hnd, C.byref(tb), C.byref(ts))
        out["t_build"], out["t_segment"] = tb.value, ts.value
        for k in range(L.ref_result_count(hnd)):
            root, score, move = C.c_int32(), C.c_double(), C.c_double()
            sol = RefSolution()
            size = L.ref_result_entry(hnd, k, C.byref(root), C.byref(score), C.byref(move), C.byref(sol))
            px = np.zeros(size, np.int32)
            L.ref_result_pixels(hnd, k, px.ctypes.data_as(C.POINTER(C.c_int32)))
            out["entries"].append({"root": root.value, "score": score.value, "move": move.value,
                                   "size": size, "pixels": px, "sol": sol.as_dict()})
    finally:
        L.ref_result_free(hnd)
    return out


def set_counting(on):
    lib().ref_set_counting(1 if on else 0)


def get_counts():
    n = lib().ref_get_counts(None, 0)
    buf = C.create_string_buffer(n + 16)
    lib().ref_get_counts(buf, n + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        if "=" in line:
            k, v = line.rsplit("=", 1)
            out[k] = int(v)
    return out
