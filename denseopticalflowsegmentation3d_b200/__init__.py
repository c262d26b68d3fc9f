"""dofs3d-b200: sm_100a implementation of the flow -> flow-graph clustering -> 3D lifting path of
DmitriyZhuravlev/DenseOpticalFlowSegmentation3D behind a C ABI (include/dofs3d.h).

The compute lives in csrc/*.cu(h) (libdofs3d.so, built in-tree by build.py).  This package is only
the Python binding used by tests/ and bench.py; there is no CPU implementation here — every call
fails loudly when the CUDA library or a GPU is missing.
"""
from .capi import Context, DofsError, Box, Stats, Params, lib_path, load_library, default_params  # noqa: F401
