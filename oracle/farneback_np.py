"""TEST INFRASTRUCTURE (oracle) — not part of the product.

numpy restatement of the flow stage the reference calls at cpp/src/segment.cpp:97-101,52:

    cv::cvtColor(BGR2GRAY) -> cv::calcOpticalFlowFarneback(g0, g1, flow, 0.5, 3, 15, 3, 5, 1.2, 0)
    -> cv::GaussianBlur(flow, flow, Size(0,0), 3.0)

The arithmetic lives in OpenCV, which is NOT in /root/reference and which the reference does not
pin (cpp/CMakeLists.txt:14 `find_package(OpenCV REQUIRED)`).  This file restates the published
algorithm of OpenCV 4.x (modules/video/src/optflowgf.cpp: FarnebackPrepareGaussian,
FarnebackPolyExp, FarnebackUpdateMatrices, FarnebackUpdateFlow_Blur, FarnebackOpticalFlowImpl::calc;
modules/imgproc: getGaussianKernel, GaussianBlur BORDER_REFLECT_101, resize INTER_LINEAR,
cvtColor's 15-bit fixed-point luma) operation by operation, including OpenCV's sliding-window box
sums.  It is pinned against cv2 4.13 (the OpenCV build in this image) in tests/test_flow_oracle.py;
the reference itself holds no test or golden vector for this stage, so cv2 is
the de-facto oracle and this file documents what it computes.
"""
import numpy as np

F32 = np.float32


def bgr2gray(bgr):
    b, g, r = (bgr[..., i].astype(np.uint32) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def gaussian_kernel(ksize, sigma):
    """cv::getGaussianKernel(ksize, sigma, CV_32F)."""
    if sigma <= 0 and ksize == 3:
        return np.array([0.25, 0.5, 0.25], F32)
    if sigma <= 0:
        sigma = ((ksize - 1) * 0.5 - 1) * 0.3 + 0.8
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-0.5 / (sigma * sigma) * x * x).astype(F32)
    s = 1.0 / k.astype(np.float64).sum()
    return (k.astype(np.float64) * s).astype(F32)


def _reflect101(idx, n):
    idx = np.abs(idx)
    period = 2 * (n - 1) if n > 1 else 1
    idx = idx % period if n > 1 else idx * 0
    return np.where(idx >= n, period - idx, idx)


def gaussian_blur(img, ksize, sigma):
    """Separable float filter, rows (horizontal) first, BORDER_REFLECT_101; float accumulation."""
    k = gaussian_kernel(ksize, sigma)
    r = ksize // 2
    img = img.astype(F32)
    H, W = img.shape[:2]
    xs = _reflect101(np.arange(-r, W + r), W)
    tmp = np.zeros_like(img)
    for i in range(ksize):
        tmp = (tmp + k[i] * img[:, xs[i:i + W]]).astype(F32)
    ys = _reflect101(np.arange(-r, H + r), H)
    out = np.zeros_like(img)
    for i in range(ksize):
        out = (out + k[i] * tmp[ys[i:i + H]]).astype(F32)
    return out


def _linear_coords(dst_n, src_n):
    scale = src_n / dst_n
    f = ((np.arange(dst_n) + 0.5) * scale - 0.5).astype(F32)
    i = np.floor(f).astype(np.int64)
    f = (f - i.astype(F32)).astype(F32)
    lo = i < 0
    i[lo], f[lo] = 0, 0
    hi = i >= src_n - 1
    i[hi], f[hi] = src_n - 1, 0
    return i, f


def resize_linear(img, Wd, Hd):
    """cv::resize(..., INTER_LINEAR) for float data: horizontal pass then vertical pass."""
    H, W = img.shape[:2]
    if (Wd, Hd) == (W, H):
        return img.astype(F32).copy()
    ix, fx = _linear_coords(Wd, W)
    iy, fy = _linear_coords(Hd, H)
    ix1, iy1 = np.minimum(ix + 1, W - 1), np.minimum(iy + 1, H - 1)
    shape = (1, Wd) + (1,) * (img.ndim - 2)
    fxb = fx.reshape(shape)
    rows = (img[:, ix] * (F32(1) - fxb) + img[:, ix1] * fxb).astype(F32)
    fyb = fy.reshape((Hd, 1) + (1,) * (img.ndim - 2))
    return (rows[iy] * (F32(1) - fyb) + rows[iy1] * fyb).astype(F32)


def prepare_gaussian(n, sigma):
    if sigma < np.finfo(F32).eps:
        sigma = n * 0.3
    x = np.arange(-n, n + 1)
    g = np.exp(-x * x / (2 * sigma * sigma)).astype(F32)
    s = 1.0 / g.astype(np.float64).sum()
    g = (g.astype(np.float64) * s).astype(F32)
    xg = (x.astype(F32) * g).astype(F32)
    xxg = ((x * x).astype(F32) * g).astype(F32)
    G = np.zeros((6, 6))
    for yy in range(-n, n + 1):
        for xx in range(-n, n + 1):
            gg = F32(g[yy + n] * g[xx + n])
            G[0, 0] += gg
            gx2 = F32(F32(gg * F32(xx)) * F32(xx))
            G[1, 1] += gx2
            G[3, 3] += F32(F32(gx2 * F32(xx)) * F32(xx))
            G[5, 5] += F32(F32(gx2 * F32(yy)) * F32(yy))
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g[n:], xg[n:], xxg[n:], inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def poly_exp(I, n, sigma):
    """FarnebackPolyExp: vertical sums in float (rows clamped), horizontal in double (columns clamped)."""
    g, xg, xxg, ig11, ig03, ig33, ig55 = prepare_gaussian(n, sigma)
    I = I.astype(F32)
    H, W = I.shape
    ys = np.arange(H)
    r0 = (I * g[0]).astype(F32)
    r1 = np.zeros_like(I)
    r2 = np.zeros_like(I)
    for k in range(1, n + 1):
        a, b = I[np.maximum(ys - k, 0)], I[np.minimum(ys + k, H - 1)]
        p = (a + b).astype(F32)
        r0 = (r0 + (g[k] * p).astype(F32)).astype(F32)
        r1 = (r1 + (xg[k] * (b - a).astype(F32)).astype(F32)).astype(F32)
        r2 = (r2 + (xxg[k] * p).astype(F32)).astype(F32)
    xs = np.arange(W)
    D = np.float64
    b1 = (r0 * g[0]).astype(F32).astype(D)
    b3 = (r1 * g[0]).astype(F32).astype(D)
    b5 = (r2 * g[0]).astype(F32).astype(D)
    b2 = np.zeros((H, W))
    b4 = np.zeros((H, W))
    b6 = np.zeros((H, W))
    for k in range(1, n + 1):
        xp, xm = np.minimum(xs + k, W - 1), np.maximum(xs - k, 0)
        tg = (r0[:, xp] + r0[:, xm]).astype(F32).astype(D)
        b1 = b1 + tg * D(g[k])
        b4 = b4 + tg * D(xxg[k])
        b2 = b2 + ((r0[:, xp] - r0[:, xm]).astype(F32) * xg[k]).astype(F32).astype(D)
        b3 = b3 + ((r1[:, xp] + r1[:, xm]).astype(F32) * g[k]).astype(F32).astype(D)
        b6 = b6 + ((r1[:, xp] - r1[:, xm]).astype(F32) * xg[k]).astype(F32).astype(D)
        b5 = b5 + ((r2[:, xp] + r2[:, xm]).astype(F32) * g[k]).astype(F32).astype(D)
    R = np.empty((H, W, 5), F32)
    R[..., 0] = (b3 * ig11).astype(F32)
    R[..., 1] = (b2 * ig11).astype(F32)
    R[..., 2] = (b1 * ig03 + b5 * ig33).astype(F32)
    R[..., 3] = (b1 * ig03 + b4 * ig33).astype(F32)
    R[..., 4] = (b6 * ig55).astype(F32)
    return R


_BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], F32)


def update_matrices(R0, R1, flow):
    """FarnebackUpdateMatrices, all float."""
    H, W = flow.shape[:2]
    one, half, quarter = F32(1), F32(0.5), F32(0.25)
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    dx, dy = flow[..., 0].astype(F32), flow[..., 1].astype(F32)
    fx = (xx.astype(F32) + dx).astype(F32)
    fy = (yy.astype(F32) + dy).astype(F32)
    x1, y1 = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(F32)).astype(F32)
    fy = (fy - y1.astype(F32)).astype(F32)
    inside = (x1 >= 0) & (x1 < W - 1) & (y1 >= 0) & (y1 < H - 1)
    xc, yc = np.clip(x1, 0, W - 2), np.clip(y1, 0, H - 2)
    a00 = ((one - fx) * (one - fy)).astype(F32)[..., None]
    a01 = (fx * (one - fy)).astype(F32)[..., None]
    a10 = ((one - fx) * fy).astype(F32)[..., None]
    a11 = (fx * fy).astype(F32)[..., None]
    w = ((((a00 * R1[yc, xc]).astype(F32) + (a01 * R1[yc, xc + 1]).astype(F32)).astype(F32)
          + (a10 * R1[yc + 1, xc]).astype(F32)).astype(F32) + (a11 * R1[yc + 1, xc + 1]).astype(F32)).astype(F32)
    r2 = np.where(inside, w[..., 0], F32(0))
    r3 = np.where(inside, w[..., 1], F32(0))
    r4 = np.where(inside, ((R0[..., 2] + w[..., 2]).astype(F32) * half).astype(F32), R0[..., 2])
    r5 = np.where(inside, ((R0[..., 3] + w[..., 3]).astype(F32) * half).astype(F32), R0[..., 3])
    r6 = np.where(inside, ((R0[..., 4] + w[..., 4]).astype(F32) * quarter).astype(F32), (R0[..., 4] * half).astype(F32))
    r2 = ((R0[..., 0] - r2).astype(F32) * half).astype(F32)
    r3 = ((R0[..., 1] - r3).astype(F32) * half).astype(F32)
    r2 = (r2 + ((r4 * dy).astype(F32) + (r6 * dx).astype(F32)).astype(F32)).astype(F32)
    r3 = (r3 + ((r6 * dy).astype(F32) + (r5 * dx).astype(F32)).astype(F32)).astype(F32)
    sx = np.ones(W, F32)
    sy = np.ones(H, F32)
    for i in range(min(5, W)):
        sx[i] = (sx[i] * _BORDER[i]).astype(F32)
    for i in range(min(5, W)):
        sx[W - 1 - i] = F32(sx[W - 1 - i] * _BORDER[i])
    for i in range(min(5, H)):
        sy[i] = F32(sy[i] * _BORDER[i])
    for i in range(min(5, H)):
        sy[H - 1 - i] = F32(sy[H - 1 - i] * _BORDER[i])
    scale = (sx[None, :] * sy[:, None]).astype(F32)
    r2, r3, r4, r5, r6 = ((v * scale).astype(F32) for v in (r2, r3, r4, r5, r6))
    M = np.empty((H, W, 5), F32)
    M[..., 0] = ((r4 * r4).astype(F32) + (r6 * r6).astype(F32)).astype(F32)
    M[..., 1] = ((r4 + r5).astype(F32) * r6).astype(F32)
    M[..., 2] = ((r5 * r5).astype(F32) + (r6 * r6).astype(F32)).astype(F32)
    M[..., 3] = ((r4 * r2).astype(F32) + (r6 * r3).astype(F32)).astype(F32)
    M[..., 4] = ((r6 * r2).astype(F32) + (r5 * r3).astype(F32)).astype(F32)
    return M


def update_flow_blur(M, block_size, sliding=True):
    """FarnebackUpdateFlow_Blur: box mean of M with replicated borders and the 2x2 solve in double.
    sliding=True follows OpenCV's running sums (float row differences accumulated in double);
    sliding=False sums every window directly in double (what the CUDA kernel does)."""
    H, W = M.shape[:2]
    m = block_size // 2
    D = np.float64
    ys, xs = np.arange(H), np.arange(W)
    if sliding:
        v0 = M[0].astype(D) * (m + 2)
        for y in range(1, m):
            v0 = v0 + M[min(y, H - 1)].astype(D)
        diff = (M[np.minimum(ys + m, H - 1)] - M[np.maximum(ys - m - 1, 0)]).astype(F32).astype(D)
        vs = v0[None] + np.cumsum(diff, axis=0)                      # [H][W][5], sequential double adds
        pad = np.concatenate([np.repeat(vs[:, :1], m + 1, 1), vs, np.repeat(vs[:, -1:], m + 1, 1)], 1)  # x in [-m-1, W+m]
        h0 = pad[:, m + 1] * (m + 2)
        for x in range(1, m):
            h0 = h0 + pad[:, m + 1 + x]
        hd = pad[:, xs + m + 1 + m] - pad[:, xs + m + 1 - m - 1]
        hs = h0[:, None] + np.cumsum(hd, axis=1)
    else:
        vs = np.zeros((H, W, 5))
        for j in range(-m, m + 1):
            vs = vs + M[np.clip(ys + j, 0, H - 1)].astype(D)
        hs = np.zeros((H, W, 5))
        for i in range(-m, m + 1):
            hs = hs + vs[:, np.clip(xs + i, 0, W - 1)]
    scale = 1.0 / (block_size * block_size)
    g11, g12, g22, h1, h2 = (hs[..., c] * scale for c in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((H, W, 2), F32)
    flow[..., 0] = ((g11 * h2 - g12 * h1) * idet).astype(F32)
    flow[..., 1] = ((g22 * h1 - g12 * h2) * idet).astype(F32)
    return flow


def farneback(prev, nxt, pyr_scale=0.5, levels=3, winsize=15, iters=3, poly_n=5, poly_sigma=1.2, sliding=True,
              pyramid=None):
    """calcOpticalFlowFarneback(prev, next, None, pyr_scale, levels, winsize, iters, poly_n, poly_sigma, 0).
    pyramid: optional callable (img_u8, ksize, sigma, Wk, Hk) -> float image, to swap the blur+resize."""
    H, W = prev.shape
    k = 0
    scale = 1.0
    while k < levels:
        scale *= pyr_scale
        if W * scale < 32 or H * scale < 32:
            break
        k += 1
    levels = k
    flow = None
    for k in range(levels, -1, -1):
        scale = pyr_scale ** k if k else 1.0
        scale = 1.0
        for _ in range(k):
            scale *= pyr_scale
        sigma = (1.0 / scale - 1) * 0.5
        smooth = max(int(np.rint(sigma * 5)) | 1, 3)
        Wk, Hk = int(np.rint(W * scale)), int(np.rint(H * scale))
        if flow is None:
            flow = np.zeros((Hk, Wk, 2), F32)
        else:
            flow = (resize_linear(flow, Wk, Hk).astype(np.float64) * (1.0 / pyr_scale)).astype(F32)
        R = []
        for img in (prev, nxt):
            if pyramid is not None:
                I = pyramid(img, smooth, sigma, Wk, Hk)
            else:
                I = resize_linear(gaussian_blur(img.astype(F32), smooth, sigma), Wk, Hk)
            R.append(poly_exp(I, poly_n, poly_sigma))
        M = update_matrices(R[0], R[1], flow)
        for it in range(iters):
            flow = update_flow_blur(M, winsize, sliding)
            if it < iters - 1:
                M = update_matrices(R[0], R[1], flow)
    return flow


def flow_blur(flow, sigma=3.0):
    """cv::GaussianBlur(flow, flow, Size(0,0), sigma) on CV_32FC2 (segment.cpp:52)."""
    ksize = int(np.rint(sigma * 8 + 1)) | 1
    return gaussian_blur(flow, ksize, sigma)
