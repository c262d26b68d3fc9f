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

# every symbol include/dofs3d.h declares
SYMBOLS = [
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

    # ---- raw device-pointer entry points (ints are device addresses, e.g. torch.Tensor.data_ptr()) ----
    def synth_frames_dev(self, seed, n_objects, first_frame, n_frames, d_bgr_ptr):
        self._ck(self.L.dofs3d_synth_frames_dev(self.h, seed, n_objects, first_frame, n_frames, C.c_void_p(d_bgr_ptr)))

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
