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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
