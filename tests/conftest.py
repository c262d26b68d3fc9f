import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """The CPU restatement (oracle/dofs3d_oracle.cpp)."""
    from oracle import cpu
    return cpu.port()


@pytest.fixture(scope="session")
def ref():
    """The unchanged reference sources (oracle/_ref), when that library was built."""
    from oracle import cpu
    if not cpu.ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return cpu.ref()


@pytest.fixture(scope="session")
def golden_pair():
    return np.load(os.path.join(GOLDEN, "pair_1052_1053.npz"))


@pytest.fixture(scope="session")
def golden_synth():
    return np.load(os.path.join(GOLDEN, "synth_320x180.npz"))


@pytest.fixture(scope="session")
def golden_lift():
    return np.load(os.path.join(GOLDEN, "lift_random.npz"))


def random_flow(seed, W, H, smooth=2.0, scale=2.0, flat=True):
    """Seeded smooth random flow field with exactly-flat patches (ties) — numpy only."""
    rng = np.random.default_rng(seed)
    f = rng.normal(size=(H, W, 2)).astype(np.float32)
    # cheap separable box smoothing, a few passes
    k = max(int(smooth), 1)
    for _ in range(3):
        for ax in (0, 1):
            acc = np.zeros_like(f)
            for s in range(-k, k + 1):
                acc += np.roll(f, s, axis=ax)
            f = (acc / np.float32(2 * k + 1)).astype(np.float32)
    f *= np.float32(scale / max(float(np.abs(f).max()), 1e-6))
    if flat:
        f[H // 4:H // 2, W // 4:W // 2] = 0.0           # exact zero-weight ties
        f[H // 2:H // 2 + H // 6, W // 3:W // 3 + W // 5] = np.float32(1.25)  # a rigidly moving block
    return np.ascontiguousarray(f)
