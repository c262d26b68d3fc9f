"""TEST INFRASTRUCTURE (oracle) — not part of the product.

Restatement of what the reference's display code takes from OpenCV (cpp/src/draw.cpp:85-160): cv::line(img, p1, p2,
color, 1) — LINE_8, thickness 1 — as OpenCV 4.x draws it (clipLine + the 8-connected LineIterator), the Point2f -> Point
rounding of draw_cube's arguments, and cv::addWeighted on 8-bit images.  Pinned against cv2 4.13 live by
tests/test_draw_oracle.py (bit-exact on thousands of random segments, inside and across the image border).
"""
import numpy as np

INT_MIN = -2 ** 31


def round_point(v):
    """Point_<int>(Point_<float>): saturate_cast<int>(float) = cvRound (half to even); NaN / out of range -> INT_MIN."""
    v = float(v)
    if v != v or not (-2147483648.0 <= v <= 2147483647.0):
        return INT_MIN
    return int(np.rint(np.float32(v)))


def _trunc_div(num, den):
    """(int64)((double)a * b / c) of clipLine: C++ evaluates (double)a * b in double, divides, truncates toward zero."""
    return int(float(num) / float(den))


def clip_line(w, h, x1, y1, x2, y2):
    """cv::clipLine(Size(w, h), pt1, pt2) on 64-bit integers: (visible, x1, y1, x2, y2)."""
    right, bottom = w - 1, h - 1
    if w <= 0 or h <= 0:
        return False, x1, y1, x2, y2
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += _trunc_div(float(a - y1) * (x2 - x1), (y2 - y1))
            y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += _trunc_div(float(a - y2) * (x2 - x1), (y2 - y1))
            y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += _trunc_div(float(a - x1) * (y2 - y1), (x2 - x1))
                x1 = a
                c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += _trunc_div(float(a - x2) * (y2 - y1), (x2 - x1))
                x2 = a
                c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def line_pixels(w, h, p1, p2):
    """Pixels (x, y) cv::line(img, p1, p2, color, 1, LINE_8) sets on a w x h image, in drawing order."""
    x1, y1, x2, y2 = int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1])
    if not (0 <= x1 < w and 0 <= x2 < w and 0 <= y1 < h and 0 <= y2 < h):
        ok, x1, y1, x2, y2 = clip_line(w, h, x1, y1, x2, y2)
        if not ok:
            return []
    dx, dy = x2 - x1, y2 - y1
    step_x, step_y = 1, 1
    if dx < 0:  # leftToRight: start from the left end
        dx, dy = -dx, -dy
        x1, y1 = x2, y2
    if dy < 0:
        dy = -dy
        step_y = -1
    vert = dy > dx
    if vert:
        dx, dy = dy, dx
    err = dx - (dy + dy)
    plus_delta, minus_delta = dx + dx, -(dy + dy)
    out = []
    x, y = x1, y1
    for _ in range(dx + 1):
        out.append((x, y))
        both = err < 0
        err += minus_delta + (plus_delta if both else 0)
        if vert:
            y += step_y
            if both:
                x += step_x
        else:
            x += step_x
            if both:
                y += step_y
    return out


def add_weighted_u8(a, alpha, b, beta):
    """cv::addWeighted(a, alpha, b, beta, 0, dst) on uint8: float arithmetic, round half to even, saturate."""
    al, be = np.float32(alpha), np.float32(beta)
    t = a.astype(np.float32) * al + b.astype(np.float32) * be
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)
