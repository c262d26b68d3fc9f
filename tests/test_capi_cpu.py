"""CPU tests of the boundary: libdofs3d.so loads without a GPU, exports every symbol include/dofs3d.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import denseopticalflowsegmentation3d_b200 as dofs
from denseopticalflowsegmentation3d_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dofs3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dofs3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = dofs.load_library()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert sorted(capi.SYMBOLS) == syms


def test_struct_layouts_match_header():
    assert ctypes.sizeof(capi.Box) == capi.BOX_DTYPE.itemsize == 216
    assert ctypes.sizeof(capi.Stats) == capi.STATS_DTYPE.itemsize == 40
    for name, _ in capi.Box._fields_:
        if name in capi.BOX_DTYPE.names:
            assert getattr(capi.Box, name).offset == capi.BOX_DTYPE.fields[name][1], name


def test_default_params_are_the_reference_constants(port):
    p = dofs.default_params()
    persp, inv, up = port.get_mats()  # get_mat / get_mat_upper (lifting_3d.cpp:441-514)
    assert np.array_equal(np.array(p.persp, np.float32).reshape(3, 3), persp)
    assert np.array_equal(np.array(p.inv, np.float32).reshape(3, 3), inv)
    assert np.array_equal(np.array(p.inv_upper, np.float32).reshape(3, 3, 3), up)
    assert (p.pyr_scale, p.levels, p.winsize, p.iters, p.poly_n, p.poly_sigma) == (0.5, 3, 15, 3, 5, 1.2)  # segment.cpp:101
    assert (p.blur_sigma, p.neighbors, p.min_size, p.score_threshold) == (3.0, 8, 500, 0.3)
    assert [list(r) for r in p.cls_size] == [[258, 84], [349, 165], [370, 180]]  # lifting_3d.cpp:255-259
    assert list(p.cls_min_convexity) == [3.0 / 4.0, 1.0 / 2.0, 20.0 / 29.0]  # graph.cpp:328-339


def test_params_for_size_rescales_the_calibration():
    """dofs3d_params_for_size: at 640x360 it is the reference's calibration bit for bit; at another size the image-side
    coordinates scale and the bird's-eye-view side stays."""
    from denseopticalflowsegmentation3d_b200 import capi
    ref = dofs.default_params()
    same = capi.params_for_size(640, 360)
    assert bytes(same) == bytes(ref)
    big = capi.params_for_size(1920, 1080)

    def apply(m, x, y):
        v = np.array(m, np.float64).reshape(3, 3) @ np.array([x, y, 1.0])
        return v[:2] / v[2]

    for x, y in [(215, 265), (90, 121), (400, 200), (625, 265)]:
        assert np.allclose(apply(big.persp, 3 * x, 3 * y), apply(ref.persp, x, y), rtol=0, atol=0.5)  # same BEV point
    for bx, by in [(100, 13000), (800, 6000), (450, 9000)]:
        assert np.allclose(apply(big.inv, bx, by), 3 * apply(ref.inv, bx, by), rtol=0, atol=2e-2)
        for c in range(3):
            assert np.allclose(apply(big.inv_upper[c], bx, by), 3 * apply(ref.inv_upper[c], bx, by), rtol=0, atol=2e-2)
    assert (big.min_size, big.winsize, big.neighbors) == (ref.min_size, ref.winsize, ref.neighbors)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dofs.DofsError) as ei:
        dofs.Context(64, 48)
    assert ei.value.status == -2  # DOFS3D_ERR_CUDA


def test_bad_arguments_rejected():
    L = dofs.load_library()
    h = ctypes.c_void_p()
    assert L.dofs3d_create(ctypes.byref(h), 0, 1, 1, 1, None) == -1  # DOFS3D_ERR_ARG
    assert L.dofs3d_create(None, 0, 64, 64, 1, None) == -1
    assert L.dofs3d_sync(None) == -1


def test_host_shim_builds_and_fails_loudly_without_gpu(tmp_path):
    """The C++ drop-in (reference cpp/inc signatures) compiles against the C ABI; without a GPU it throws."""
    import subprocess
    import torch
    from denseopticalflowsegmentation3d_b200 import build as b
    b.build_host()
    exe = b.build_host_program(os.path.join(ROOT, "tests", "cpp", "test_host_shim.cpp"), str(tmp_path / "shim_test"))
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
