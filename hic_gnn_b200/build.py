"""Build libhicgat_sm100.so in-tree with nvcc for sm_100a (no torch headers involved).

    python -m hic_gnn_b200.build [--force] [--trace]

``--trace`` additionally builds ``libhicgat_trace.so`` with ``-DHICGAT_TRACE`` (per-CTA %globaltimer stamps in the
pair-loss kernels, read by ``scripts/trace_pairloss.py``); never loaded by the package itself.

The shared library is self-contained (static cudart) and exposes only the C ABI declared in
``include/hicgat.h``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhicgat_sm100.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden",
    "-cudart", "static",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    deps.append(os.path.join(ROOT, "include", "hicgat.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    if trace:
        return _build(os.path.join(PKG, "libhicgat_trace.so"), os.path.join(PKG, "build_trace"), ["-DHICGAT_TRACE"], verbose)
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(PKG, "build"), [], verbose)


def _build(lib: str, objdir: str, extra: list, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    procs = []
    os.makedirs(objdir, exist_ok=True)
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out)
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libhicgat_sm100.so")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", lib + ".tmp", *objs]
    subprocess.run(link, check=True)
    os.replace(lib + ".tmp", lib)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
