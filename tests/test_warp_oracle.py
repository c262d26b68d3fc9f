"""oracle/warp_np.py (cv::warpPerspective, INTER_CUBIC, BORDER_REPLICATE on 8-bit images) against cv2 live: bit-exact."""
import numpy as np
import pytest


def test_warp_oracle_equals_cv2():
    cv2 = pytest.importorskip("cv2")
    from oracle import warp_np
    rng = np.random.default_rng(0)
    img = cv2.GaussianBlur(rng.integers(0, 256, (90, 160, 3), dtype=np.uint8), (0, 0), 1.5)
    src = np.float32([[50, 70], [20, 30], [75, 30], [150, 70]])
    dst = np.float32([[20, 300], [20, 40], [180, 40], [180, 300]])
    M = cv2.getPerspectiveTransform(src, dst).astype(np.float32)
    ref = cv2.warpPerspective(img, M, (220, 340), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    assert np.array_equal(warp_np.warp_perspective_cubic(img, M, 220, 340), ref)
    gray = np.ascontiguousarray(img[..., 0])
    ref1 = cv2.warpPerspective(gray, M, (101, 77), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    assert np.array_equal(warp_np.warp_perspective_cubic(gray[..., None], M, 101, 77)[..., 0], ref1)


def test_warp_oracle_bev_geometry():
    """The reference's own homography (get_mat) and BEV size, a few bands of the 2500 x 14000 output."""
    cv2 = pytest.importorskip("cv2")
    from oracle import cpu, warp_np
    rng = np.random.default_rng(1)
    img = cv2.GaussianBlur(rng.integers(0, 256, (360, 640, 3), dtype=np.uint8), (0, 0), 2.0)
    persp, _, _ = cpu.port().get_mats()
    ref = cv2.warpPerspective(img, persp, (2500, 14000), flags=cv2.INTER_CUBIC, borderMode=cv2.BORDER_REPLICATE)
    for band in ((0, 32), (6000, 6040), (13968, 14000)):
        assert np.array_equal(warp_np.warp_perspective_cubic(img, persp, 2500, 14000, rows=band), ref[band[0]:band[1]])
