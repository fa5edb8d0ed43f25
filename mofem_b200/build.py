"""Builds mofem_b200/libmofem_b200.so in-tree with nvcc for sm_100a.

    python -m mofem_b200.build [--force] [--verbose]

The integration kernels are compiled with -fmad=false (every FMA there is an explicit fma();
see csrc/fem_device.cuh); the HBM-bound CSR / PCG kernels use the default contraction.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmofem_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
SOURCES = {
    "integrate.cu": ["-fmad=false"],
    "sample.cu": ["-fmad=false"],
    "rho.cu": ["-fmad=false"],
    "overlay.cu": ["-fmad=false"],
    "csr.cu": [],
    "scan.cu": [],
    "pcg.cu": [],
    "api.cu": [],
}


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(os.path.dirname(HERE), "include", "mofem_b200.h"))
    return hdr


HOST_LIB = os.path.join(HERE, "liboverlay_host.so")


def build_host(force=False, verbose=False):
    """Host build of the overlay core (csrc/overlay_host.cpp): preprocessing without a device + the CPU parity check."""
    src = os.path.join(CSRC, "overlay_host.cpp")
    hdrs = [os.path.join(CSRC, "overlay_core.h"), os.path.join(CSRC, "overlay_general.h")]
    if force or not os.path.exists(HOST_LIB) or os.path.getmtime(HOST_LIB) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs]):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", src, "-o", HOST_LIB]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return HOST_LIB


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    build_host(force, verbose)
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    newest_hdr = max(os.path.getmtime(h) for h in _deps())
    objs, relink = [], force or not os.path.exists(LIB)
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(path), newest_hdr):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            relink = True
    if relink:
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
