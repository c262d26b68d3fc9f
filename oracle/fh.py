"""TEST INFRASTRUCTURE (oracle) — not part of the product.

ctypes binding of oracle/fh_oracle.cpp: the CPU restatement of the reference's Felzenszwalb-style flow segmentation
(/root/reference/graph.py:77-177, main.py:310-315).  The reference addresses its array as img[x][y] with node id
= y * width + x; here `flow` is the usual [H][W][2] field and node id = row * W + col, i.e. the reference is given
flow.transpose(1, 0, 2) (tools/make_golden_fh.py does exactly that).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_build", "libdofs3d_fh_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE, "fh"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(PATH):
            build()
        _lib = C.CDLL(PATH)
        _lib.fh_segment_flow.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double,
                                         C.c_double, C.c_int, C.c_void_p]
    return _lib


def segment_flow(flow, K=10.0, min_size=100, neighbors=8, flow_dist=5.0, edge_dist=5.0, stage=3):
    """labels[H][W] = Forest.find(pixel) after segment_graph_flow (stage 3), or after its first / second pass."""
    f = np.ascontiguousarray(flow, np.float32)
    H, W = f.shape[:2]
    labels = np.empty((H, W), np.int32)
    n = lib().fh_segment_flow(f.ctypes.data, W, H, 1 if neighbors == 8 else 0, float(K), int(min_size), float(flow_dist),
                              float(edge_dist), int(stage), labels.ctypes.data)
    return labels, n
