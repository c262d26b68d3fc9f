"""Builds csrc/dofs3d.cu into denseopticalflowsegmentation3d_b200/libdofs3d.so for sm_100a (in-tree, so the
library travels with the source snapshot)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "dofs3d.cu")
OUT = os.path.join(HERE, "libdofs3d.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    deps.append(os.path.join(HERE, "..", "include", "dofs3d.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [NVCC] + FLAGS + ["-o", OUT, SRC]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return OUT


HOST_SRC = os.path.join(HERE, "host", "src", "dofs3d_host.cpp")
HOST_OUT = os.path.join(HERE, "libdofs3d_host.so")
HOST_INC = os.path.join(HERE, "host", "inc")


def build_host(force=False):
    """The C++ host shim (reference cpp/inc entry points over the C ABI): g++, links against libdofs3d.so."""
    deps = [HOST_SRC, OUT] + [os.path.join(HOST_INC, f) for f in os.listdir(HOST_INC)]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_OUT) for d in deps):
        return HOST_OUT
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", "-o", HOST_OUT, HOST_SRC,
                           "-L" + HERE, "-ldofs3d", "-Wl,-rpath,$ORIGIN"])
    return HOST_OUT


def build_host_program(src, out):
    """Compiles a C++ program against the shim headers and libraries (used by tests/ and the demo driver)."""
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + HOST_INC, src, "-o", out, "-L" + HERE,
                           "-ldofs3d_host", "-ldofs3d", "-Wl,-rpath," + HERE])
    return out


def build_multi_gpu_driver(out=None):
    """examples/multi_gpu_driver.cpp: the C++ multi-GPU streaming driver (links the C ABI, the CUDA runtime and NCCL)."""
    root = os.path.dirname(HERE)
    out = out or os.path.join(root, "examples", "multi_gpu_driver")
    cuda = os.path.dirname(os.path.dirname(NVCC))
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(root, "include"),
                           "-I" + os.path.join(cuda, "include"), os.path.join(root, "examples", "multi_gpu_driver.cpp"),
                           "-o", out, "-L" + HERE, "-ldofs3d", "-L" + os.path.join(cuda, "lib64"), "-lcudart", "-lnccl",
                           "-lpthread", "-Wl,-rpath," + HERE])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    build_host(force="--force" in sys.argv)
    if "--examples" in sys.argv:
        print(build_multi_gpu_driver())
    print(OUT)
    print(HOST_OUT)
