"""cv2-vs-cv2 scatter of calcOpticalFlowFarneback on the synthetic 1080p (8 objects) and 4K (60 objects) pairs of the
bench: the default build (SIMD + IPP) against the plain C++ build (cv2.setUseOptimized(False), cv2.ipp.setUseIPP(False)).
CPU only.  Writes profiles/r02_flow_scatter.json — the measured justification of the flow bar in tests/test_gpu_fullsize.py.
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from denseopticalflowsegmentation3d_b200 import synth  # noqa: E402


def flow(g0, g1, optimized):
    cv2.setUseOptimized(optimized)
    cv2.ipp.setUseIPP(optimized)
    return cv2.calcOpticalFlowFarneback(g0, g1, None, 0.5, 3, 15, 3, 5, 1.2, 0)


out = {"cv2": cv2.__version__}
for W, H, n_obj in ((1920, 1080, 8), (3840, 2160, 60)):
    fr = synth.frames(1234, n_obj, 0, 2, W, H)
    g0, g1 = (cv2.cvtColor(x, cv2.COLOR_BGR2GRAY) for x in fr)
    a, b = flow(g0, g1, True), flow(g0, g1, False)
    d = a.astype(np.float64) - b
    e = np.sqrt((d ** 2).sum(-1))
    row = {"max": float(e.max()), "mean": float(e.mean()), "median": float(np.median(e)),
           "p99": float(np.percentile(e, 99)), "p99.9": float(np.percentile(e, 99.9)),
           "frac_gt": {str(t): float((e > t).mean()) for t in (1e-5, 1e-4, 1e-3, 1e-2, 1e-1)},
           "flow_magnitude_max": float(np.sqrt((a.astype(np.float64) ** 2).sum(-1)).max())}
    for band in (16, 64):
        inner = e[band:-band, band:-band]
        m = np.ones_like(e, bool)
        m[band:-band, band:-band] = False
        row[f"interior_{band}px"] = {"max": float(inner.max()), "mean": float(inner.mean())}
        row[f"border_band_{band}px"] = {"max": float(e[m].max()), "mean": float(e[m].mean())}
    if W == 1920:  # a second implementation on the same pair: the numpy oracle with OpenCV's running box sums and with direct sums
        from oracle import farneback_np
        stable = ~(cv2.dilate((e > 1e-4).astype(np.uint8), np.ones((33, 33), np.uint8)) > 0)
        for name, sliding in (("oracle_running_sums_vs_cv2", True), ("oracle_direct_sums_vs_cv2", False)):
            dd = farneback_np.farneback(g0, g1, sliding=sliding).astype(np.float64) - a
            ee = np.sqrt((dd ** 2).sum(-1))
            row[name] = {"stable_fraction": float(stable.mean()), "stable_max": float(ee[stable].max()),
                         "stable_mean": float(ee[stable].mean()), "stable_frac_gt_1e-3": float((ee[stable] > 1e-3).mean()),
                         "all_max": float(ee.max()), "frac_gt_1e-3": float((ee > 1e-3).mean())}
    out[f"{W}x{H}_{n_obj}obj"] = row
    print(W, H, json.dumps(row))
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_flow_scatter.json"), "w"), indent=1)
