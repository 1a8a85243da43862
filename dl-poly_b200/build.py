"""Builds libdlpgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdlpgpu.so")
SOURCES = ["ctx.cu", "cells.cu", "forces.cu", "halo.cu", "spme.cu", "hostio.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--std=c++17",
         "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v"]   # host pass un-fused too (aarch64 hosts would contract h_dcell / h_geometry)
# -fmad=false where a floating-point expression decides an integer (cell index, list membership, halo / migration
# thresholds): those follow the reference's un-fused IEEE arithmetic bit for bit.  forces.cu contracts to FMA (the pair
# terms only need the 1e-9 / 1e-10 bars); its cutoff tests use explicit _rn intrinsics.
EXTRA = {"ctx.cu": ["-fmad=false"], "cells.cu": ["-fmad=false"], "halo.cu": ["-fmad=false"], "forces.cu": [], "spme.cu": [], "hostio.cu": []}


HOST = os.path.join(HERE, "host")
HOST_OUT = os.path.join(HOST, "_build")
HOST_LIB = os.path.join(HOST_OUT, "libdlpoly_host.so")
HOST_CHECK = os.path.join(HOST_OUT, "dlpoly_check")
CXX = os.environ.get("CXX", "g++")
# -ffp-contract=off: the host-side table generators and decisions follow the reference's un-fused IEEE arithmetic
HOST_FLAGS = ["-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-Wall", "-I", os.path.join(HERE, "..", "include")]


def build_host(force=False):
    """The C++ host side above the C ABI (host/dlpoly_host.cpp -> libdlpoly_host.so) and its check driver (dlpoly_check)."""
    srcs = [os.path.join(HOST, f) for f in ("dlpoly_host.cpp", "dlpoly_host.hpp", "dlpoly_check.cpp")] + [OUT]
    if not force and os.path.exists(HOST_LIB) and os.path.exists(HOST_CHECK) and \
            all(os.path.getmtime(s) <= min(os.path.getmtime(HOST_LIB), os.path.getmtime(HOST_CHECK)) for s in srcs):
        return HOST_CHECK
    os.makedirs(HOST_OUT, exist_ok=True)
    link = ["-L", HERE, "-ldlpgpu"]
    cmds = [[CXX] + HOST_FLAGS + ["-shared", "-o", HOST_LIB, os.path.join(HOST, "dlpoly_host.cpp")] + link + ["-Wl,-rpath,$ORIGIN/../.."],
            [CXX] + HOST_FLAGS + ["-o", HOST_CHECK, os.path.join(HOST, "dlpoly_check.cpp"), "-L", HOST_OUT, "-ldlpoly_host"] + link +
            ["-Wl,-rpath,$ORIGIN:$ORIGIN/../.."]]
    for cmd in cmds:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("host build failed: %s" % " ".join(cmd))
    return HOST_CHECK


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dlpgpu.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_host()
        return OUT
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)

    def cc(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + EXTRA[src] + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, obj, r.returncode, r.stdout

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(cc, SOURCES))
    log = []
    for src, obj, rc, out in res:
        log.append("== %s\n%s" % (src, out))
        if rc != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [NVCC, "-shared", "-o", OUT] + [r[1] for r in res] + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
    build_host(force=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
