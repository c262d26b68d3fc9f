"""oracle/draw_np.py (cv::line LINE_8 thickness 1 = clipLine + LineIterator; cv::addWeighted on uint8) against cv2 live."""
import numpy as np
import pytest


def test_line_pixels_equal_cv2_line():
    cv2 = pytest.importorskip("cv2")
    from oracle import draw_np as D
    rng = np.random.default_rng(0)
    W, H = 97, 61
    for it in range(3000):
        lo, hi = ((-40, 140), (0, 60), (-2000, 2000))[it % 3]   # across the border, inside, far outside
        p = rng.integers(lo, hi, 4)
        p1, p2 = (int(p[0]), int(p[1])), (int(p[2]), int(p[3]))
        img = np.zeros((H, W), np.uint8)
        cv2.line(img, p1, p2, 255, 1)
        mine = np.zeros((H, W), np.uint8)
        for x, y in D.line_pixels(W, H, p1, p2):
            mine[y, x] = 255
        assert np.array_equal(img, mine), (p1, p2)


def test_add_weighted_equals_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import draw_np as D
    a = np.arange(256, dtype=np.uint8)[:, None].repeat(256, 1)
    b = a.T.copy()
    op = 2.0 / 5.0
    assert np.array_equal(cv2.addWeighted(a, 1.0 - op, b, op, 0), D.add_weighted_u8(a, 1.0 - op, b, op))
