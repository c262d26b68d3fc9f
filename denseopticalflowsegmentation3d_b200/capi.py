"""ctypes binding of include/dofs3d.h (the C ABI of libdofs3d.so)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    """The in-tree library; DOFS3D_LIB names another build of the same sources (kernel tuning experiments)."""
    return os.environ.get("DOFS3D_LIB") or os.path.join(_HERE, "libdofs3d.so")


class DofsError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"dofs3d status {status}: {message}")
        self.status = status


class Params(C.Structure):
    _fields_ = [
        ("persp", C.c_float * 9), ("inv", C.c_float * 9), ("inv_upper", (C.c_float * 9) * 3),
        ("pyr_scale", C.c_double), ("levels", C.c_int), ("winsize", C.c_int), ("iters", C.c_int),
        ("poly_n", C.c_int), ("poly_sigma", C.c_double), ("blur_sigma", C.c_double), ("neighbors", C.c_int),
        ("min_size", C.c_int), ("score_threshold", C.c_double), ("cls_size", (C.c_int * 2) * 3),
        ("cls_min_convexity", C.c_double * 3),
    ]


class Box(C.Structure):
    _fields_ = [
        ("root", C.c_int32), ("size", C.c_int32), ("cls", C.c_int32), ("parent_box", C.c_int32),
        ("bbox", C.c_int32 * 4), ("time", C.c_uint32), ("mean_flow", C.c_float * 2), ("pad_", C.c_float),
        ("score", C.c_double), ("move", C.c_double), ("orient", C.c_double), ("w_error", C.c_double),
        ("h_error", C.c_double), ("ps_bev", C.c_float * 8), ("rectangle", C.c_float * 8),
        ("lower_face", C.c_float * 8), ("upper_face", C.c_float * 8),
    ]


BOX_DTYPE = np.dtype([
    ("root", "<i4"), ("size", "<i4"), ("cls", "<i4"), ("parent_box", "<i4"), ("bbox", "<i4", (4,)),
    ("time", "<u4"), ("mean_flow", "<f4", (2,)), ("pad_", "<f4"), ("score", "<f8"), ("move", "<f8"),
    ("orient", "<f8"), ("w_error", "<f8"), ("h_error", "<f8"), ("ps_bev", "<f4", (4, 2)),
    ("rectangle", "<f4", (4, 2)), ("lower_face", "<f4", (4, 2)), ("upper_face", "<f4", (4, 2)),
], align=True)
assert BOX_DTYPE.itemsize == C.sizeof(Box) == 216, (BOX_DTYPE.itemsize, C.sizeof(Box))


class Stats(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_edges", "n_merges", "n_levels", "n_candidates", "n_scored", "n_boxes",
                                         "longest_chain", "final_root", "sort_fallback", "replay_exact_chunks")]


STATS_DTYPE = np.dtype([(k, "<i4") for k in ("n_edges", "n_merges", "n_levels", "n_candidates", "n_scored",
                                             "n_boxes", "longest_chain", "final_root", "sort_fallback", "replay_exact_chunks")])

class FhParams(C.Structure):
    """dofs3d_fh_params: the Felzenszwalb mode of the reference's Python twin (graph.py:156-177)."""
    _fields_ = [("k", C.c_double), ("min_size", C.c_int), ("neighbors", C.c_int), ("flow_dist", C.c_double),
                ("edge_dist", C.c_double), ("stage", C.c_int)]


BEV_WIDTH, BEV_HEIGHT = 2500, 14000


class Run(C.Structure):
    _fields_ = [("start", C.c_uint32), ("label", C.c_int32)]


RUN_DTYPE = np.dtype([("start", "<u4"), ("label", "<i4")])
LABELS_I32, LABELS_U16, LABELS_RLE = 0, 1, 2


class Outputs(C.Structure):
    """dofs3d_outputs: where the results of one call go (host or device addresses, depending on the entry point)."""
    _fields_ = [("label_format", C.c_int), ("labels", C.c_void_p), ("n_runs", C.c_void_p), ("max_runs", C.c_int),
                ("boxes", C.c_void_p), ("n_boxes", C.c_void_p), ("max_boxes", C.c_int), ("stats", C.c_void_p)]


def runs_to_labels(runs, n_runs, n_pixels):
    """Dense int32 label image of one frame from its run-length form."""
    r = runs[:n_runs]
    ends = np.append(r["start"][1:], n_pixels).astype(np.int64)
    return np.repeat(r["label"], ends - r["start"].astype(np.int64)).astype(np.int32)


# every symbol include/dofs3d.h declares
SYMBOLS = [
    "dofs3d_process_ex", "dofs3d_process_ex_dev", "dofs3d_segment_ex", "dofs3d_stream_begin", "dofs3d_stream_submit",
    "dofs3d_stream_collect", "dofs3d_fh_default_params", "dofs3d_segment_fh", "dofs3d_warp_perspective",
    "dofs3d_forest_create", "dofs3d_forest_destroy", "dofs3d_forest_find", "dofs3d_forest_merge", "dofs3d_forest_new_merge",
    "dofs3d_forest_num_sets", "dofs3d_forest_last_score", "dofs3d_forest_bbox", "dofs3d_forest_boxes", "dofs3d_forest_pixels",
    "dofs3d_bev_transform", "dofs3d_render", "dofs3d_pack_boxes_dev", "dofs3d_node_state", "dofs3d_scored_merges", "dofs3d_pinned_alloc", "dofs3d_pinned_free",
    "dofs3d_default_params", "dofs3d_params_for_size", "dofs3d_create", "dofs3d_destroy", "dofs3d_sync", "dofs3d_last_error", "dofs3d_stream",
    "dofs3d_launch_count", "dofs3d_device_bytes", "dofs3d_gray", "dofs3d_gray_dev", "dofs3d_flow", "dofs3d_blur",
    "dofs3d_segment", "dofs3d_paint", "dofs3d_lift", "dofs3d_edges_sorted", "dofs3d_process", "dofs3d_process_dev",
    "dofs3d_segment_dev", "dofs3d_flow_dev", "dofs3d_synth_frames_dev", "dofs3d_set_timing", "dofs3d_get_timing",
]

_lib = None


def load_library():
    """Loads libdofs3d.so (no compute happens at load time, so this works without a GPU)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise DofsError(-2, f"{path} is missing: run `python -m denseopticalflowsegmentation3d_b200.build` "
                            "(there is no CPU fallback)")
    L = C.CDLL(path)
    vp, ip, fp, u8p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    L.dofs3d_default_params.argtypes = [C.POINTER(Params)]
    L.dofs3d_default_params.restype = None
    L.dofs3d_params_for_size.argtypes = [C.POINTER(Params), C.c_int, C.c_int]
    L.dofs3d_params_for_size.restype = None
    L.dofs3d_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Params)]
    L.dofs3d_destroy.argtypes = [vp]
    L.dofs3d_destroy.restype = None
    L.dofs3d_sync.argtypes = [vp]
    L.dofs3d_last_error.argtypes = [vp]
    L.dofs3d_last_error.restype = C.c_char_p
    L.dofs3d_stream.argtypes = [vp]
    L.dofs3d_stream.restype = C.c_void_p
    L.dofs3d_launch_count.argtypes = [vp]
    L.dofs3d_launch_count.restype = C.c_longlong
    L.dofs3d_device_bytes.argtypes = [vp]
    L.dofs3d_device_bytes.restype = C.c_longlong
    L.dofs3d_gray.argtypes = [vp, u8p, C.c_int, u8p]
    L.dofs3d_gray_dev.argtypes = [vp, u8p, C.c_int, u8p]
    L.dofs3d_flow.argtypes = [vp, u8p, u8p, C.c_int, fp]
    L.dofs3d_flow_dev.argtypes = [vp, u8p, u8p, C.c_int, fp]
    L.dofs3d_blur.argtypes = [vp, fp, C.c_int, fp]
    L.dofs3d_segment.argtypes = [vp, fp, C.c_int, C.c_int, ip, vp, ip, C.c_int, vp, fp]
    L.dofs3d_segment_dev.argtypes = [vp, fp, C.c_int, C.c_int, ip, vp, ip, C.c_int, vp]
    L.dofs3d_lift.argtypes = [vp, fp, ip, ip, C.c_int, vp]
    L.dofs3d_paint.argtypes = [vp, C.c_int, C.c_double, ip, u8p]
    L.dofs3d_edges_sorted.argtypes = [vp, fp, ip, ip, vp]
    L.dofs3d_edges_sorted.restype = C.c_longlong
    L.dofs3d_process.argtypes = [vp, u8p, C.c_int, ip, vp, ip, C.c_int, vp]
    L.dofs3d_process_dev.argtypes = [vp, u8p, C.c_int, ip, vp, ip, C.c_int, vp]
    L.dofs3d_synth_frames_dev.argtypes = [vp, C.c_uint32, C.c_int, C.c_int, C.c_int, u8p]
    L.dofs3d_process_ex.argtypes = [vp, u8p, C.c_int, C.POINTER(Outputs)]
    L.dofs3d_process_ex_dev.argtypes = [vp, u8p, C.c_int, C.POINTER(Outputs)]
    L.dofs3d_segment_ex.argtypes = [vp, fp, C.c_int, C.c_int, C.POINTER(Outputs)]
    L.dofs3d_stream_begin.argtypes = [vp]
    L.dofs3d_stream_submit.argtypes = [vp, u8p, C.c_int, C.POINTER(Outputs)]
    L.dofs3d_stream_collect.argtypes = [vp, C.POINTER(C.c_int)]
    L.dofs3d_node_state.argtypes = [vp, C.c_int, C.c_int, ip, fp, ip]
    L.dofs3d_scored_merges.argtypes = [vp, C.c_int, C.c_int, ip, ip, vp, u8p]
    L.dofs3d_fh_default_params.argtypes = [C.POINTER(FhParams)]
    L.dofs3d_fh_default_params.restype = None
    L.dofs3d_segment_fh.argtypes = [vp, fp, C.POINTER(FhParams), ip, ip]
    L.dofs3d_warp_perspective.argtypes = [vp, u8p, C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, u8p]
    L.dofs3d_bev_transform.argtypes = [vp, u8p, u8p]
    L.dofs3d_pack_boxes_dev.argtypes = [vp, C.c_int, vp, C.c_int, ip]
    L.dofs3d_forest_create.argtypes = [vp, fp, C.POINTER(vp)]
    L.dofs3d_forest_destroy.argtypes = [vp]
    L.dofs3d_forest_destroy.restype = None
    L.dofs3d_forest_find.argtypes = [vp, C.c_int, C.POINTER(C.c_int32)]
    L.dofs3d_forest_merge.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.dofs3d_forest_new_merge.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_int]
    L.dofs3d_forest_num_sets.argtypes = [vp, C.POINTER(C.c_int32)]
    L.dofs3d_forest_last_score.argtypes = [vp, C.c_int, C.POINTER(C.c_double)]
    L.dofs3d_forest_bbox.argtypes = [vp, C.c_int, ip]
    L.dofs3d_forest_boxes.argtypes = [vp, C.c_int, vp]
    L.dofs3d_forest_pixels.argtypes = [vp, C.c_int, C.c_int, ip]
    L.dofs3d_render.argtypes = [vp, C.c_int, C.c_double, u8p]
    L.dofs3d_pinned_alloc.argtypes = [C.c_size_t]
    L.dofs3d_pinned_alloc.restype = C.c_void_p
    L.dofs3d_pinned_free.argtypes = [vp]
    L.dofs3d_pinned_free.restype = None
    L.dofs3d_set_timing.argtypes = [vp, C.c_int]
    L.dofs3d_get_timing.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]
    _lib = L
    return L


def default_params():
    p = Params()
    load_library().dofs3d_default_params(C.byref(p))
    return p


def params_for_size(width, height):
    """default_params with the reference's 640x360 calibration quads rescaled to width x height."""
    p = Params()
    load_library().dofs3d_params_for_size(C.byref(p), width, height)
    return p


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Context:
    """One dofs3d_ctx: frames of `width` x `height`, at most `max_pairs` frame pairs per call, on GPU `device`."""

    def __init__(self, width, height, max_pairs=1, device=0, params=None):
        self.L = load_library()
        self.W, self.H, self.N, self.max_pairs = width, height, width * height, max_pairs
        self.h = C.c_void_p()
        self._stream_pending = []
        rc = self.L.dofs3d_create(C.byref(self.h), device, width, height, max_pairs,
                                  C.byref(params) if params is not None else None)
        if rc != 0:
            msg = self.L.dofs3d_last_error(self.h).decode() if self.h else "dofs3d_create failed (no CUDA device?)"
            if self.h:
                self.L.dofs3d_destroy(self.h)
                self.h = C.c_void_p()
            raise DofsError(rc, msg)

    def close(self):
        if getattr(self, "h", None):
            self.L.dofs3d_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise DofsError(rc, self.L.dofs3d_last_error(self.h).decode())
        return rc

    # ---- stages ---------------------------------------------------------------------------
    def gray(self, bgr):
        bgr = np.ascontiguousarray(bgr, np.uint8).reshape(-1, self.H, self.W, 3)
        out = np.empty(bgr.shape[:3], np.uint8)
        self._ck(self.L.dofs3d_gray(self.h, _ptr(bgr), bgr.shape[0], _ptr(out)))
        return out

    def flow(self, gray0, gray1):
        g0 = np.ascontiguousarray(gray0, np.uint8).reshape(-1, self.H, self.W)
        g1 = np.ascontiguousarray(gray1, np.uint8).reshape(-1, self.H, self.W)
        assert g0.shape == g1.shape
        out = np.empty(g0.shape + (2,), np.float32)
        self._ck(self.L.dofs3d_flow(self.h, _ptr(g0), _ptr(g1), g0.shape[0], _ptr(out)))
        return out

    def blur(self, flow):
        f = np.ascontiguousarray(flow, np.float32).reshape(-1, self.H, self.W, 2)
        out = np.empty_like(f)
        self._ck(self.L.dofs3d_blur(self.h, _ptr(f), f.shape[0], _ptr(out)))
        return out

    def segment(self, flow, already_blurred=False, max_boxes=1024, want_labels=True, want_blurred=False):
        """get_segmented_array + get_best_segments for a batch of flow fields [n][H][W][2]."""
        f = np.ascontiguousarray(flow, np.float32).reshape(-1, self.H, self.W, 2)
        n = f.shape[0]
        labels = np.empty((n, self.H, self.W), np.int32) if want_labels else None
        boxes = np.zeros((n, max_boxes), BOX_DTYPE)
        n_boxes = np.zeros(n, np.int32)
        stats = np.zeros(n, STATS_DTYPE)
        blurred = np.empty_like(f) if want_blurred else None
        self._ck(self.L.dofs3d_segment(self.h, _ptr(f), 1 if already_blurred else 0, n, _ptr(labels), _ptr(boxes),
                                       _ptr(n_boxes), max_boxes, _ptr(stats), _ptr(blurred)))
        return {"labels": labels, "boxes": [boxes[i, :n_boxes[i]] for i in range(n)], "n_boxes": n_boxes,
                "stats": stats, "flow_blurred": blurred}

    def paint(self, n_pairs, min_score=0.7, bgr=None):
        """Display semantics of plot_best_segments_simple for the results of the last segment/process call."""
        painted = np.empty((n_pairs, self.H, self.W), np.int32)
        if bgr is not None:
            bgr = np.ascontiguousarray(bgr, np.uint8).reshape(n_pairs, self.H, self.W, 3)
        self._ck(self.L.dofs3d_paint(self.h, n_pairs, float(min_score), _ptr(painted), _ptr(bgr)))
        return painted, bgr

    def render(self, bgr_frames, min_score=0.7):
        """plot_best_segments_simple's returned frame for the results of the last segment/process call."""
        fr = np.array(bgr_frames, np.uint8, copy=True).reshape(-1, self.H, self.W, 3)
        self._ck(self.L.dofs3d_render(self.h, fr.shape[0], float(min_score), _ptr(fr)))
        return fr

    def lift(self, direction, bbox, cls):
        d = np.ascontiguousarray(direction, np.float32).reshape(-1, 2)
        b = np.ascontiguousarray(bbox, np.int32).reshape(-1, 4)
        c = np.ascontiguousarray(cls, np.int32).reshape(-1)
        out = np.zeros(d.shape[0], BOX_DTYPE)
        self._ck(self.L.dofs3d_lift(self.h, _ptr(d), _ptr(b), _ptr(c), d.shape[0], _ptr(out)))
        return out

    def edges_sorted(self, flow_blurred):
        f = np.ascontiguousarray(flow_blurred, np.float32).reshape(self.H, self.W, 2)
        start = np.empty(4 * self.N, np.int32)
        end = np.empty(4 * self.N, np.int32)
        wbits = np.empty(4 * self.N, np.uint64)
        e = self._ck(self.L.dofs3d_edges_sorted(self.h, _ptr(f), _ptr(start), _ptr(end), _ptr(wbits)))
        return start[:e], end[:e], wbits[:e].view(np.float64)

    def process(self, bgr_frames, max_boxes=1024, want_labels=True):
        """Whole path for n+1 consecutive BGR frames [n+1][H][W][3] -> n pairs."""
        fr = np.ascontiguousarray(bgr_frames, np.uint8).reshape(-1, self.H, self.W, 3)
        n = fr.shape[0] - 1
        labels = np.empty((n, self.H, self.W), np.int32) if want_labels else None
        boxes = np.zeros((n, max_boxes), BOX_DTYPE)
        n_boxes = np.zeros(n, np.int32)
        stats = np.zeros(n, STATS_DTYPE)
        self._ck(self.L.dofs3d_process(self.h, _ptr(fr), fr.shape[0], _ptr(labels), _ptr(boxes), _ptr(n_boxes),
                                       max_boxes, _ptr(stats)))
        return {"labels": labels, "boxes": [boxes[i, :n_boxes[i]] for i in range(n)], "n_boxes": n_boxes,
                "stats": stats}

    # ---- compact result formats ------------------------------------------------------------
    def _host_outputs(self, n, label_format, max_boxes, max_runs):
        if label_format == LABELS_RLE:
            labels = np.zeros((n, max_runs), RUN_DTYPE)
        else:
            labels = np.empty((n, self.H, self.W), np.uint16 if label_format == LABELS_U16 else np.int32)
        arr = dict(labels=labels, n_runs=np.zeros(n, np.int32), boxes=np.zeros((n, max_boxes), BOX_DTYPE),
                   n_boxes=np.zeros(n, np.int32), stats=np.zeros(n, STATS_DTYPE))
        o = Outputs(label_format, labels.ctypes.data, arr["n_runs"].ctypes.data, max_runs, arr["boxes"].ctypes.data,
                    arr["n_boxes"].ctypes.data, max_boxes, arr["stats"].ctypes.data)
        return arr, o

    @staticmethod
    def _finish_outputs(arr, n=None):
        n = len(arr["n_boxes"]) if n is None else n
        return {"labels": arr["labels"][:n], "n_runs": arr["n_runs"][:n], "n_boxes": arr["n_boxes"][:n],
                "boxes": [arr["boxes"][i, :arr["n_boxes"][i]] for i in range(n)], "stats": arr["stats"][:n]}

    def process_ex(self, bgr_frames, label_format=LABELS_U16, max_boxes=1024, max_runs=65536):
        """dofs3d_process_ex: whole path with u16 / run-length / int32 labels."""
        fr = np.ascontiguousarray(bgr_frames, np.uint8).reshape(-1, self.H, self.W, 3)
        arr, o = self._host_outputs(fr.shape[0] - 1, label_format, max_boxes, max_runs)
        self._ck(self.L.dofs3d_process_ex(self.h, _ptr(fr), fr.shape[0], C.byref(o)))
        return self._finish_outputs(arr)

    def segment_ex(self, flow, already_blurred=False, label_format=LABELS_U16, max_boxes=1024, max_runs=65536):
        f = np.ascontiguousarray(flow, np.float32).reshape(-1, self.H, self.W, 2)
        arr, o = self._host_outputs(f.shape[0], label_format, max_boxes, max_runs)
        self._ck(self.L.dofs3d_segment_ex(self.h, _ptr(f), 1 if already_blurred else 0, f.shape[0], C.byref(o)))
        return self._finish_outputs(arr)

    # ---- streaming ------------------------------------------------------------------------
    def stream_begin(self):
        self._ck(self.L.dofs3d_stream_begin(self.h))
        self._stream_pending = []

    def stream_submit(self, bgr_frames, label_format=LABELS_U16, max_boxes=1024, max_runs=65536):
        """Asynchronous; the chunk's arrays are kept alive here until stream_collect returns them."""
        fr = np.ascontiguousarray(bgr_frames, np.uint8).reshape(-1, self.H, self.W, 3)
        arr, o = self._host_outputs(self.max_pairs, label_format, max_boxes, max_runs)
        self._ck(self.L.dofs3d_stream_submit(self.h, _ptr(fr), fr.shape[0], C.byref(o)))
        self._stream_pending.append((fr, arr, o))

    def stream_collect(self):
        n = C.c_int(0)
        self._ck(self.L.dofs3d_stream_collect(self.h, C.byref(n)))
        _, arr, _ = self._stream_pending.pop(0)
        return self._finish_outputs(arr, n.value)

    def process_stream(self, bgr_frames, chunk_pairs=None, **kw):
        """A whole clip [n+1][H][W][3] through the streaming entry points in chunks; concatenated results."""
        fr = np.ascontiguousarray(bgr_frames, np.uint8).reshape(-1, self.H, self.W, 3)
        ch = chunk_pairs or self.max_pairs
        self.stream_begin()
        outs, pos, inflight = [], 0, 0
        while pos < fr.shape[0]:
            nf = min(ch + 1 if pos == 0 else ch, fr.shape[0] - pos)
            if inflight == 2:
                outs.append(self.stream_collect())
                inflight -= 1
            self.stream_submit(fr[pos:pos + nf], **kw)
            inflight += 1
            pos += nf
        while inflight:
            outs.append(self.stream_collect())
            inflight -= 1
        return {"labels": np.concatenate([o["labels"] for o in outs]), "n_runs": np.concatenate([o["n_runs"] for o in outs]),
                "n_boxes": np.concatenate([o["n_boxes"] for o in outs]), "boxes": [b for o in outs for b in o["boxes"]],
                "stats": np.concatenate([o["stats"] for o in outs])}

    # ---- Felzenszwalb mode, bird's-eye-view warp ------------------------------------------------
    def segment_fh(self, flow, K=10.0, min_size=100, neighbors=8, flow_dist=5.0, edge_dist=5.0, stage=3):
        """graph.py's segment_graph_flow on ONE flow field [H][W][2]: (labels[H][W] = root id per pixel, n_components)."""
        f = np.ascontiguousarray(flow, np.float32).reshape(self.H, self.W, 2)
        p = FhParams(float(K), int(min_size), int(neighbors), float(flow_dist), float(edge_dist), int(stage))
        labels = np.empty((self.H, self.W), np.int32)
        n = C.c_int32(0)
        self._ck(self.L.dofs3d_segment_fh(self.h, _ptr(f), C.byref(p), _ptr(labels), C.byref(n)))
        return labels, n.value

    def warp_perspective(self, img, mat, out_w, out_h):
        """cv::warpPerspective(img, mat, (out_w, out_h), INTER_CUBIC, BORDER_REPLICATE) on a u8 image [H][W](C)."""
        img = np.ascontiguousarray(img, np.uint8)
        ch = 1 if img.ndim == 2 else img.shape[2]
        m = np.ascontiguousarray(mat, np.float32).reshape(9)
        out = np.empty((out_h, out_w) if img.ndim == 2 else (out_h, out_w, ch), np.uint8)
        self._ck(self.L.dofs3d_warp_perspective(self.h, _ptr(img), img.shape[1], img.shape[0], ch, _ptr(m), out_w, out_h,
                                                _ptr(out)))
        return out

    def bev_transform(self, bgr_frame):
        """The reference's transform(frame, get_mat().first): the 2500 x 14000 bird's-eye-view image."""
        fr = np.ascontiguousarray(bgr_frame, np.uint8).reshape(self.H, self.W, 3)
        out = np.empty((BEV_HEIGHT, BEV_WIDTH, 3), np.uint8)
        self._ck(self.L.dofs3d_bev_transform(self.h, _ptr(fr), _ptr(out)))
        return out

    # ---- per-node state of the last call ----------------------------------------------------
    def node_state(self, pair, node):
        size = C.c_int32(0)
        flow = np.zeros(2, np.float32)
        bbox = np.zeros(4, np.int32)
        self._ck(self.L.dofs3d_node_state(self.h, pair, node, C.byref(size), _ptr(flow), _ptr(bbox)))
        return {"size": size.value, "mean_flow": flow, "bbox": bbox}

    def scored_merges(self, pair):
        n = self._ck(self.L.dofs3d_scored_merges(self.h, pair, 0, None, None, None, None))
        root, time = np.zeros(n, np.int32), np.zeros(n, np.uint32)
        score, kept = np.zeros(n, np.float64), np.zeros(n, np.uint8)
        if n:
            self._ck(self.L.dofs3d_scored_merges(self.h, pair, n, _ptr(root), _ptr(time), _ptr(score), _ptr(kept)))
        return {"root": root, "time": time, "score": score, "kept": kept.astype(bool)}

    def last_scores(self, pair):
        """Forest::get_segment_best_score for every root that has one: {root: score of its latest scored merge}."""
        m = self.scored_merges(pair)
        out, when = {}, {}
        for r, t, s in zip(m["root"].tolist(), m["time"].tolist(), m["score"].tolist()):
            if r not in when or when[r] < t:
                when[r], out[r] = t, s
        return out

    # ---- raw device-pointer entry points (ints are device addresses, e.g. torch.Tensor.data_ptr()) ----
    def synth_frames_dev(self, seed, n_objects, first_frame, n_frames, d_bgr_ptr):
        self._ck(self.L.dofs3d_synth_frames_dev(self.h, seed, n_objects, first_frame, n_frames, C.c_void_p(d_bgr_ptr)))

    def pack_boxes_dev(self, n_pairs, d_out, capacity, d_total):
        self._ck(self.L.dofs3d_pack_boxes_dev(self.h, n_pairs, C.c_void_p(d_out), capacity, C.c_void_p(d_total)))

    def process_ex_dev(self, d_bgr_ptr, n_frames, outputs):
        self._ck(self.L.dofs3d_process_ex_dev(self.h, C.c_void_p(d_bgr_ptr), n_frames, C.byref(outputs)))

    def process_dev(self, d_bgr_ptr, n_frames, d_labels=None, d_boxes=None, d_n_boxes=None, max_boxes=0, d_stats=None):
        v = lambda p: None if p is None else C.c_void_p(p)  # noqa: E731
        self._ck(self.L.dofs3d_process_dev(self.h, v(d_bgr_ptr), n_frames, v(d_labels), v(d_boxes), v(d_n_boxes),
                                           max_boxes, v(d_stats)))

    def segment_dev(self, d_flow, already_blurred, n, d_labels=None, d_boxes=None, d_n_boxes=None, max_boxes=0,
                    d_stats=None):
        v = lambda p: None if p is None else C.c_void_p(p)  # noqa: E731
        self._ck(self.L.dofs3d_segment_dev(self.h, v(d_flow), 1 if already_blurred else 0, n, v(d_labels), v(d_boxes),
                                           v(d_n_boxes), max_boxes, v(d_stats)))

    def flow_dev(self, d_gray0, d_gray1, n, d_flow_out):
        self._ck(self.L.dofs3d_flow_dev(self.h, C.c_void_p(d_gray0), C.c_void_p(d_gray1), n, C.c_void_p(d_flow_out)))

    def gray_dev(self, d_bgr, n_frames, d_gray):
        self._ck(self.L.dofs3d_gray_dev(self.h, C.c_void_p(d_bgr), n_frames, C.c_void_p(d_gray)))

    def sync(self):
        self._ck(self.L.dofs3d_sync(self.h))

    @property
    def stream(self):
        return self.L.dofs3d_stream(self.h)

    @property
    def launch_count(self):
        return self.L.dofs3d_launch_count(self.h)

    @property
    def device_bytes(self):
        return self.L.dofs3d_device_bytes(self.h)

    def set_timing(self, on):
        self._ck(self.L.dofs3d_set_timing(self.h, 1 if on else 0))

    def timing(self):
        """{stage: (total ms, number of timed intervals)} of the last call (set_timing(True) first)."""
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        cnt = (C.c_int * 64)()
        n = self._ck(self.L.dofs3d_get_timing(self.h, names, ms, cnt, 64))
        return {names[i].decode(): (ms[i], cnt[i]) for i in range(n)}


class Forest:
    """The reference's incremental Forest (graph.hpp:72-114) over device state, one call at a time (dofs3d_forest_*)."""

    def __init__(self, ctx, flow):
        self.ctx, self.L = ctx, ctx.L
        f = np.ascontiguousarray(flow, np.float32).reshape(ctx.H, ctx.W, 2)
        self.h = C.c_void_p()
        ctx._ck(self.L.dofs3d_forest_create(ctx.h, _ptr(f), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            self.L.dofs3d_forest_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def find(self, n):
        r = C.c_int32(0)
        self.ctx._ck(self.L.dofs3d_forest_find(self.h, int(n), C.byref(r)))
        return r.value

    def merge(self, a, b):
        r = C.c_int32(0)
        self.ctx._ck(self.L.dofs3d_forest_merge(self.h, int(a), int(b), C.byref(r)))
        return r.value

    def new_merge(self, a, b, score_threshold=0.3, min_size=500):
        self.ctx._ck(self.L.dofs3d_forest_new_merge(self.h, int(a), int(b), float(score_threshold), int(min_size)))

    @property
    def num_sets(self):
        r = C.c_int32(0)
        self.ctx._ck(self.L.dofs3d_forest_num_sets(self.h, C.byref(r)))
        return r.value

    def last_score(self, node):
        r = C.c_double(0)
        self.ctx._ck(self.L.dofs3d_forest_last_score(self.h, int(node), C.byref(r)))
        return r.value

    def bbox(self, node):
        bb = np.zeros(4, np.int32)
        rc = self.ctx._ck(self.L.dofs3d_forest_bbox(self.h, int(node), _ptr(bb)))
        return bb if rc == 1 else None

    def boxes(self, max_boxes=4096):
        out = np.zeros(max_boxes, BOX_DTYPE)
        n = self.ctx._ck(self.L.dofs3d_forest_boxes(self.h, max_boxes, _ptr(out)))
        return out[:n]

    def pixels(self, root, size):
        out = np.zeros(max(size, 1), np.int32)
        n = self.ctx._ck(self.L.dofs3d_forest_pixels(self.h, int(root), int(size), _ptr(out)))
        return np.sort(out[:min(n, size)])


def box_pixel_sets(labels, boxes):
    """Pixel set (sorted pixel ids) of every box of one frame, from the label image and the nesting
    chain: pixels labelled b, plus pixels labelled with any box whose parent_box chain reaches b."""
    lab = labels.reshape(-1)
    nb = len(boxes)
    own = [np.nonzero(lab == b)[0] for b in range(nb)]
    sets = [[own[b]] for b in range(nb)]
    for b in range(nb):
        p = int(boxes[b]["parent_box"])
        hops = 0
        while p >= 0:
            sets[p].append(own[b])
            p = int(boxes[p]["parent_box"])
            hops += 1
            assert hops <= nb
    return [np.sort(np.concatenate(s)).astype(np.int32) for s in sets]
