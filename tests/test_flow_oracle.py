"""CPU tests: pin the numpy restatement of the flow stage (oracle/farneback_np.py) against cv2 4.13 —
the OpenCV the reference would call (segment.cpp:97-101,52) — live, and against the committed outputs of
cv2 in the authoring container (tests/golden)."""
import numpy as np
import pytest

from oracle import farneback_np as fn

cv2 = pytest.importorskip("cv2")


def epe(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return np.sqrt((d ** 2).sum(-1))


def test_gray_bit_exact(golden_pair):
    assert np.array_equal(fn.bgr2gray(golden_pair["bgr0_head"]), golden_pair["gray0_head"])
    rng = np.random.default_rng(0)
    bgr = rng.integers(0, 256, size=(33, 47, 3), dtype=np.uint8)
    assert np.array_equal(fn.bgr2gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


@pytest.mark.parametrize("name", ["golden_pair", "golden_synth"])
def test_farneback_restatement_matches_cv2_golden(name, request):
    g = request.getfixturevalue(name)
    f = fn.farneback(g["gray0"], g["gray1"])
    e = epe(f, g["flow"])
    assert e.max() <= 2e-4 and e.mean() <= 3e-6, (e.max(), e.mean())


def test_flow_blur_restatement(golden_pair):
    assert np.abs(fn.flow_blur(golden_pair["flow"]) - golden_pair["flow_blurred"]).max() <= 1e-5


def test_pyramid_pieces_match_cv2_live():
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(90, 160)).astype(np.float32)
    for ks, sg in ((3, 0.0), (3, 0.5), (9, 1.5), (19, 3.5)):
        assert np.abs(fn.gaussian_blur(img, ks, sg) - cv2.GaussianBlur(img, (ks, ks), sg, sigmaY=sg)).max() <= 1e-4
    # power-of-two ratios (every level of the 640x360, 1080p and 4K pyramids) are exact averages; for other
    # ratios cv2's IPP path places the taps ~5e-6 apart from the published formula
    for Wd, Hd, tol in ((80, 45, 1e-4), (40, 45, 1e-4), (320, 180, 1e-4), (40, 22, 2e-3), (20, 11, 2e-3)):
        assert np.abs(fn.resize_linear(img, Wd, Hd) - cv2.resize(img, (Wd, Hd), interpolation=cv2.INTER_LINEAR)).max() <= tol
    f2 = rng.normal(size=(45, 80, 2)).astype(np.float32)
    assert np.array_equal(fn.resize_linear(f2, 160, 90), cv2.resize(f2, (160, 90), interpolation=cv2.INTER_LINEAR))
