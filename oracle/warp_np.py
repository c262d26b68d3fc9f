"""TEST INFRASTRUCTURE (oracle) — not part of the product.

numpy restatement of cv::warpPerspective(img, result, mat, dsize, INTER_CUBIC, BORDER_REPLICATE) on 8-bit images — what the
reference's `transform` (cpp/src/lifting_3d.cpp:516-522) calls with dsize = 2500 x 14000 to build its bird's-eye-view
image (segment.cpp:143,196).  OpenCV is not part of the reference tree and not version-pinned (cpp/CMakeLists.txt:14);
this follows the published algorithm of OpenCV 4.x modules/imgproc/src/imgwarp.cpp:
  * the matrix is inverted in double (cv::invert of a 3x3: adjugate / determinant);
  * every destination pixel maps to source coordinates in 1/32-pixel fixed point:
        X = cvRound(X0 / W0 * 32) with W0 == 0 -> 0, clamped to the int range; integer part X >> 5, fraction X & 31;
  * bicubic weights (a = -0.75) of the two fractions come from a 32 x 32 table of 4 x 4 products rounded to 15-bit fixed
    point, each table entry adjusted so that its 16 weights sum to 2^15 exactly (initInterTab2D);
  * result = saturate_u8((sum(w * src) + 2^14) >> 15), source coordinates clamped to the image (BORDER_REPLICATE).
Pinned by tests/test_warp_oracle.py against cv2 4.13 live (bit-exact).
"""
import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS
COEF_BITS = 15
COEF_SCALE = 1 << COEF_BITS


def _cubic_coeffs(x):
    a = np.float32(-0.75)
    x = np.float32(x)
    one = np.float32(1)
    c0 = ((a * (x + one) - np.float32(5) * a) * (x + one) + np.float32(8) * a) * (x + one) - np.float32(4) * a
    c1 = ((a + np.float32(2)) * x - (a + np.float32(3))) * x * x + one
    c2 = ((a + np.float32(2)) * (one - x) - (a + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    return np.array([c0, c1, c2, c3], np.float32)


_itab = None


def cubic_table():
    """[32*32][4][4] int16 weights: entry (fy * 32 + fx), rows = y taps."""
    global _itab
    if _itab is not None:
        return _itab
    tab1 = np.stack([_cubic_coeffs(np.float32(i) * np.float32(1.0 / TAB)) for i in range(TAB)])  # [32][4]
    it = np.zeros((TAB * TAB, 4, 4), np.int16)
    for i in range(TAB):
        for j in range(TAB):
            v = np.outer(tab1[i], tab1[j]).astype(np.float32)              # vy * vx in float
            w = np.clip(np.rint(v.astype(np.float64) * COEF_SCALE), -32768, 32767).astype(np.int64)  # saturate_cast<short>(float): cvRound
            isum = int(w.sum())
            if isum != COEF_SCALE:
                diff = isum - COEF_SCALE
                Mk, mk = (2, 2), (2, 2)
                for k1 in (2, 3):
                    for k2 in (2, 3):
                        if w[k1, k2] < w[mk]:
                            mk = (k1, k2)
                        elif w[k1, k2] > w[Mk]:
                            Mk = (k1, k2)
                if diff < 0:
                    w[Mk] -= diff
                else:
                    w[mk] -= diff
            it[i * TAB + j] = w.astype(np.int16)
    _itab = it
    return it


def invert3x3(m):
    m = np.asarray(m, np.float64).reshape(3, 3)
    d = (m[0, 0] * (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) - m[0, 1] * (m[1, 0] * m[2, 2] - m[1, 2] * m[2, 0])
         + m[0, 2] * (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]))
    if d == 0:
        return np.zeros((3, 3))
    d = 1.0 / d
    t = np.empty((3, 3))
    t[0, 0] = (m[1, 1] * m[2, 2] - m[1, 2] * m[2, 1]) * d
    t[0, 1] = (m[0, 2] * m[2, 1] - m[0, 1] * m[2, 2]) * d
    t[0, 2] = (m[0, 1] * m[1, 2] - m[0, 2] * m[1, 1]) * d
    t[1, 0] = (m[1, 2] * m[2, 0] - m[1, 0] * m[2, 2]) * d
    t[1, 1] = (m[0, 0] * m[2, 2] - m[0, 2] * m[2, 0]) * d
    t[1, 2] = (m[0, 2] * m[1, 0] - m[0, 0] * m[1, 2]) * d
    t[2, 0] = (m[1, 0] * m[2, 1] - m[1, 1] * m[2, 0]) * d
    t[2, 1] = (m[0, 1] * m[2, 0] - m[0, 0] * m[2, 1]) * d
    t[2, 2] = (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]) * d
    return t


def warp_perspective_cubic(img, mat, out_w, out_h, rows=None):
    """img [H][W][C] u8 -> [out_h][out_w][C] u8 (rows = optional (y0, y1) band of the output)."""
    img = np.ascontiguousarray(img, np.uint8)
    H, W, C = img.shape
    M = invert3x3(np.asarray(mat, np.float32).astype(np.float64))
    y0, y1 = rows if rows else (0, out_h)
    ys, xs = np.mgrid[y0:y1, 0:out_w].astype(np.float64)
    X0 = M[0, 0] * xs + M[0, 1] * ys + M[0, 2]
    Y0 = M[1, 0] * xs + M[1, 1] * ys + M[1, 2]
    W0 = M[2, 0] * xs + M[2, 1] * ys + M[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        Wi = np.where(W0 != 0, TAB / W0, 0.0)
    fX = np.clip(X0 * Wi, -2147483648.0, 2147483647.0)
    fY = np.clip(Y0 * Wi, -2147483648.0, 2147483647.0)
    X = np.rint(fX).astype(np.int64)
    Y = np.rint(fY).astype(np.int64)
    sx = np.clip(X >> INTER_BITS, -32768, 32767) - 1   # saturate_cast<short>, then the 4x4 window starts one to the left
    sy = np.clip(Y >> INTER_BITS, -32768, 32767) - 1
    alpha = (Y & (TAB - 1)) * TAB + (X & (TAB - 1))
    wts = cubic_table()[alpha].astype(np.int64)        # [h][w][4][4]
    acc = np.zeros((y1 - y0, out_w, C), np.int64)
    for ky in range(4):
        yy = np.clip(sy + ky, 0, H - 1)
        for kx in range(4):
            xx = np.clip(sx + kx, 0, W - 1)
            acc += wts[..., ky, kx, None] * img[yy, xx].astype(np.int64)
    return np.clip((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS, 0, 255).astype(np.uint8)
