"""Build the reference's OWN CPU implementation of the hot path into oracle/_ref/ (TEST INFRASTRUCTURE).

The reference CPU sampler/reindex (srcs/cpp/src/quiver/quiver.cpp:21-129, srcs/cpp/include/quiver/quiver.cpu.hpp)
compiles from three of its own source files with plain g++ against this image's torch headers.  Nothing is copied:
the sources are compiled where they lie under /root/reference and only the resulting .so lands in oracle/_ref/
(git-ignored, but shipped to the GPU box by gpurun).  Two flavours:

  torch_quiver_ref      as shipped by the reference's setup.py:40-46 (no -fopenmp => at::parallel_for is serial)
  torch_quiver_ref_omp  same sources + -fopenmp, so at::parallel_for uses every host core

Used by tests/ (structural parity, golden-vector generation) and bench.py (--impl reference, cpu_baseline).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QUIVER_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SRCS = ["srcs/cpp/src/quiver/quiver.cpp", "srcs/cpp/src/quiver/cpu/tensor.cpp", "srcs/cpp/src/quiver/torch/module.cpp"]


def ext_path(name):
    return os.path.join(OUT, name + sysconfig.get_config_var("EXT_SUFFIX"))


def build_one(name, openmp):
    import torch
    from torch.utils import cpp_extension

    out = ext_path(name)
    srcs = [os.path.join(REF, s) for s in SRCS]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs):
        return out
    os.makedirs(OUT, exist_ok=True)
    inc = [os.path.join(REF, "srcs/cpp/include")] + cpp_extension.include_paths() + [sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-std=c++17", "-O3", "-fPIC", "-shared", "-w", f"-DTORCH_EXTENSION_NAME={name}",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    if openmp:
        cmd += ["-fopenmp"]
    cmd += [f"-I{p}" for p in inc] + srcs + ["-o", out, f"-L{libdir}", f"-Wl,-rpath,{libdir}",
                                             "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    if openmp:
        cmd += ["-l:libgomp.so.1"]
    print("[build_ref]", name, "...", flush=True)
    subprocess.check_call(cmd)
    return out


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} not present: keeping any prebuilt files in {OUT}")
        return 0
    build_one("torch_quiver_ref", openmp=False)
    try:
        build_one("torch_quiver_ref_omp", openmp=True)
    except subprocess.CalledProcessError as e:  # OpenMP flavour is optional
        print("[build_ref] openmp flavour failed:", e)
    return 0


if __name__ == "__main__":
    sys.exit(main())
