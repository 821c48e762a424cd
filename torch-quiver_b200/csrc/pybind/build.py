"""Build the pybind11 adapter (torch_quiver_pybind.cpp: the reference's `torch_quiver` plugin surface over the C ABI).

    python torch-quiver_b200/csrc/pybind/build.py                 -> torch-quiver_b200/torch_quiver_pybind/torch_quiver_pb*.so
    python torch-quiver_b200/csrc/pybind/build.py --name torch_quiver   -> the drop-in module name the reference's Python
                                                                           package imports (srcs/python/quiver/*.py)
The in-repo artefact is called torch_quiver_pb so that it can be imported next to the ctypes adapter package
`torch_quiver` in one process (tests/test_gpu_pybind_adapter.py).  g++ against this image's torch headers; links
libquiver_b200.so through an $ORIGIN-relative rpath, so the pair stays relocatable."""
import argparse
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(os.path.dirname(HERE))
ROOT = os.path.dirname(PKG)
OUT_DIR = os.path.join(PKG, "torch_quiver_pybind")
LIB_DIR = os.path.join(PKG, "torch_quiver")


def ext_path(name="torch_quiver_pb"):
    return os.path.join(OUT_DIR, name + sysconfig.get_config_var("EXT_SUFFIX"))


def build(name="torch_quiver_pb"):
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(HERE, "torch_quiver_pybind.cpp")
    out = ext_path(name)
    deps = [src, os.path.join(ROOT, "include", "quiver_b200.h"), os.path.join(LIB_DIR, "libquiver_b200.so"), __file__]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = [os.path.join(ROOT, "include")] + cpp_extension.include_paths() + [sysconfig.get_paths()["include"],
                                                                             os.path.join(cuda_home, "include")]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-w", f"-DTORCH_EXTENSION_NAME={name}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    cmd += [f"-I{p}" for p in inc] + [src, "-o", out, f"-L{LIB_DIR}", "-l:libquiver_b200.so", "-Wl,-rpath,$ORIGIN/../torch_quiver",
                                      f"-L{tlib}", f"-Wl,-rpath,{tlib}", "-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu",
                                      "-ltorch_cuda", "-ltorch_python"]
    print("[pybind adapter] g++", os.path.basename(src), "->", os.path.relpath(out, ROOT), flush=True)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--name", default="torch_quiver_pb")
    print(build(ap.parse_args().name))
