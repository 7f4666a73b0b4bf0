#!/usr/bin/env python
"""build.py — compiles the engine for sm_100a with nvcc (no CMake needed; CMakeLists.txt does the same for users).

Outputs (in-tree, git-ignored, shipped to the GPU box by gpurun):
  stringsearchlib_b200/lib/libssb_core.so          engine core (CUDA kernels + C API of include/ssb_engine.h)
  stringsearchlib_b200/lib/libStringSearchLib.so   README guid-keyed surface  (include/StringSearchLib.h)
  stringsearchlib_b200/lib/libStringSearchLib_h.so as-shipped handle surface  (include/StringSearchLib_handle.h)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(ROOT, "stringsearchlib_b200", "csrc")
LIB = os.path.join(ROOT, "stringsearchlib_b200", "lib")
OBJ = os.path.join(ROOT, "build", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CUFLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"] + ARCH + os.environ.get("SSB_NVCC_FLAGS", "").split()
CU = ["primitives.cu", "build.cu", "sieve.cu", "search.cu", "merge.cu", "digest.cu", "engine.cu"]


def newer(src_list, out):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in src_list)


def run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return p.stdout


def build(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith((".h", ".cuh"))]
    headers += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    jobs = []
    objs = []
    for f in CU:
        src = os.path.join(SRC, f)
        obj = os.path.join(OBJ, f.replace(".cu", ".o"))
        objs.append(obj)
        if force or newer([src] + headers, obj):
            jobs.append(([NVCC] + CUFLAGS + ["-c", src, "-o", obj], obj + ".log"))
    with ThreadPoolExecutor(max_workers=len(CU)) as ex:
        list(ex.map(lambda j: run(*j), jobs))
    core = os.path.join(LIB, "libssb_core.so")
    if force or jobs or not os.path.exists(core):
        run([NVCC, "-shared", "-cudart", "static", "-o", core] + objs + ARCH)
    for name, src in (("libStringSearchLib.so", "capi_guid.cpp"), ("libStringSearchLib_h.so", "capi_handle.cpp")):
        out = os.path.join(LIB, name)
        srcp = os.path.join(SRC, src)
        if force or jobs or newer([srcp] + headers, out):
            run(["g++", "-std=c++17", "-O2", "-fPIC", "-fvisibility=hidden", "-shared", "-pthread", srcp, "-o", out,
                 "-L" + LIB, "-lssb_core", "-Wl,-rpath,$ORIGIN", "-lstdc++"])
    # micro-benchmark behind DESIGN.md §5 (random sector reads out of HBM; not part of the product)
    micro = os.path.join(ROOT, "scripts", "micro", "random_sectors.cu")
    mout = os.path.join(LIB, "random_sectors")
    if force or newer([micro], mout):
        run([NVCC, "-O3", "-lineinfo"] + ARCH + ["-o", mout, micro])
    if verbose:
        for f in CU:
            log = os.path.join(OBJ, f.replace(".cu", ".o.log"))
            if os.path.exists(log):
                print(open(log).read())
    return core


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB)
