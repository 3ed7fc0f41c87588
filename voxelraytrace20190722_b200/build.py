"""In-tree build of libvrt.so (sm_100a) with nvcc.

``python -m voxelraytrace20190722_b200.build`` or :func:`build_native`.
The shared object is written next to this file so that it travels to the GPU
box with the repo snapshot; it is git-ignored.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libvrt%s.so" % os.environ.get("VRT_LIB_SUFFIX", ""))
SOURCES = ["vrt_api.cu", "vrt_build.cu", "vrt_trace.cu", "vrt_gi.cu"]
HEADERS = ["vrt_exact.cuh", "vrt_internal.h", "vrt_prims.cuh", "vrt_gi.cuh", os.path.join("..", "..", "include", "vrt.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    # bit-exactness with the FMA-free x86-64 reference (DESIGN.md "Arithmetic"):
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libvrt.so cannot be built (there is no CPU fallback)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_native(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(PKG, "build", s.replace(".cu", "%s.o" % os.environ.get("VRT_LIB_SUFFIX", "")))
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("VRT_EXTRA_NVCC", "").split(), "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    with open(os.path.join(PKG, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    if verbose:
        sys.stderr.write("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
