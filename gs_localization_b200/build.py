"""In-tree build of the sm_100a kernels + C ABI into gs_localization_b200/libgsr_b200.so.

Plain nvcc, no torch headers: the library is a C-ABI shared object
(include/gsr_b200.h).  nvcc cross-compiles without a GPU, so this runs on the CPU box;
the built .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgsr_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--cudart", "shared",
    "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "gsr_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in sources():
        obj = os.path.join(HERE, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [NVCC, *FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    rpaths = ["/usr/local/cuda/lib64"]
    try:
        import nvidia.cuda_runtime  # the libcudart torch itself loads
        rpaths.insert(0, os.path.join(list(nvidia.cuda_runtime.__path__)[0], "lib"))
    except Exception:
        pass
    link = [NVCC, "-shared", "--cudart", "shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs,
            "-ccbin", FLAGS[-1]]
    for r in rpaths:
        link += ["-Xlinker", "-rpath", "-Xlinker", r]
    subprocess.check_call(link)
    return LIB


BINDING_SRC = os.path.join(HERE, "binding", "torch_binding.cpp")
BINDING = os.path.join(HERE, "_gsr_torch.so")


def build_binding(force: bool = False) -> str:
    """The compiled torch-facing glue (binding/torch_binding.cpp -> _gsr_torch.so): host C++ only, linked against
    libgsr_b200.so (found next to it through $ORIGIN) and the torch libraries of the running interpreter."""
    build(force=False)
    deps = [BINDING_SRC, os.path.join(HERE, "..", "include", "gsr_b200.h"), __file__]
    if not force and os.path.exists(BINDING) and all(os.path.getmtime(d) <= os.path.getmtime(BINDING) for d in deps):
        return BINDING
    import sysconfig

    import torch
    tdir = os.path.dirname(torch.__file__)
    cxx = FLAGS[-1]
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-w",
           "-DTORCH_EXTENSION_NAME=_gsr_torch", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-isystem", os.path.join(tdir, "include"), "-isystem", os.path.join(tdir, "include", "torch", "csrc", "api", "include"),
           "-isystem", sysconfig.get_paths()["include"], "-isystem", "/usr/local/cuda/include",
           BINDING_SRC, "-o", BINDING, "-L" + HERE, "-l:libgsr_b200.so", "-L" + os.path.join(tdir, "lib"),
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + os.path.join(tdir, "lib")]
    subprocess.check_call(cmd)
    return BINDING


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_binding(force="--force" in sys.argv))
