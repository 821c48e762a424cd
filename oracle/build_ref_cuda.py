"""Build the reference's OWN CUDA extension for sm_100a into oracle/_ref/ (TEST INFRASTRUCTURE, never shipped).

The reference's three CUDA translation units (srcs/cpp/src/quiver/cuda/quiver_{sample,feature,comm}.cu) plus its CPU
files and pybind module compile UNMODIFIED with nvcc 12.9 for sm_100a, provided the thrust algorithm headers its code
uses without including (sort / sequence / unique / scan / ...: older CUDA toolkits pulled them in transitively) are
pre-included -- `-include oracle/ref_cuda_shim.h` (ours; it only #includes NVIDIA headers).  Sources are compiled where
they lie under /root/reference; only objects and the .so land in oracle/_ref/ (git-ignored, shipped by gpurun).

The result, oracle/_ref/torch_quiver_ref_cuda*.so, is the reference's `torch_quiver` module under another name:
  * tests/test_gpu_vs_reference_cuda.py -- live L1 oracle: the reference kernels EXECUTING on the same B200 vs ours;
  * bench.py `ref_gpu_baseline` -- the reference kernels' sampled edges/s and gather GB/s on the same batches
    ("the reference kernels recompiled for sm_100a" is the bar).
nvcc cross-compiles here without a GPU (~2-4 min per TU, the three TUs in parallel).
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QUIVER_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
NAME = "torch_quiver_ref_cuda"
CU = ["srcs/cpp/src/quiver/cuda/quiver_sample.cu", "srcs/cpp/src/quiver/cuda/quiver_feature.cu",
      "srcs/cpp/src/quiver/cuda/quiver_comm.cu"]
CPP = ["srcs/cpp/src/quiver/quiver.cpp", "srcs/cpp/src/quiver/cpu/tensor.cpp", "srcs/cpp/src/quiver/torch/module.cpp"]


def ext_path():
    return os.path.join(OUT, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref_cuda] {REF} not present: keeping any prebuilt files in {OUT}")
        return 0
    import torch
    from torch.utils import cpp_extension
    import nvidia.nccl

    out = ext_path()
    srcs = [os.path.join(REF, s) for s in CU + CPP]
    shim = os.path.join(HERE, "ref_cuda_shim.h")
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(s) for s in srcs + [shim, __file__]):
        return 0
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    nccl_root = list(nvidia.nccl.__path__)[0]
    inc = [os.path.join(REF, "srcs/cpp/include")] + cpp_extension.include_paths("cuda") + \
          [sysconfig.get_paths()["include"], os.path.join(nccl_root, "include")]
    defs = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DHAVE_CUDA",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs, objs = [], []
    for s in CU:
        obj = os.path.join(OUT, "obj", os.path.basename(s) + ".o")
        objs.append(obj)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "--expt-extended-lambda",
               "--expt-relaxed-constexpr", "-w", "-Xcompiler", "-fPIC", "-include", shim] + defs + \
              [f"-I{p}" for p in inc] + ["-c", os.path.join(REF, s), "-o", obj]
        print("[build_ref_cuda] nvcc", os.path.basename(s), flush=True)
        procs.append(subprocess.Popen(cmd))
    for s in CPP:
        obj = os.path.join(OUT, "obj", os.path.basename(s) + ".o")
        objs.append(obj)
        cmd = ["g++", "-std=c++17", "-O3", "-fPIC", "-w"] + defs + [f"-I{p}" for p in inc] + \
              ["-c", os.path.join(REF, s), "-o", obj]
        procs.append(subprocess.Popen(cmd))
    if any(p.wait() != 0 for p in procs):
        print("[build_ref_cuda] compilation failed")
        return 1
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    nlib = os.path.join(nccl_root, "lib")
    cmd = ["g++", "-shared"] + objs + ["-o", out, f"-L{tlib}", f"-Wl,-rpath,{tlib}", f"-L{nlib}", f"-Wl,-rpath,{nlib}",
                                       "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64",
                                       "-lc10", "-lc10_cuda", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-ltorch_python",
                                       "-l:libnccl.so.2", "-lcudart", "-lcurand"]
    subprocess.check_call(cmd)
    print("[build_ref_cuda] built", out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
