"""In-tree native builds (nvcc / g++), no JIT cache.

`python -m solidstateamfoam_b200._build` or `__graft_entry__.build()` compiles:
  * lib/libssam_gpu.so      hand-written sm_100a CUDA kernels + the C-ABI (include/ssam_gpu.h)
  * lib/libssam_meshgen.so  synthetic mesh generators (blockMesh/snappy/decomposePar stand-ins)
  * lib/ssam_case_driver    C++ host-side mirror of the reference's RTS interfaces + solver call replay
The oracle (test infrastructure) is built by oracle/Makefile, see build_oracle().
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "lib")
CSRC = os.path.join(PKG, "csrc")

NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def _walk(d: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; the CUDA path cannot be built")
    return p


def build_cuda(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libssam_gpu.so")
    cu = _walk(CSRC, (".cu",))
    deps = cu + _walk(CSRC, (".cuh", ".h", ".inc")) + _walk(os.path.join(ROOT, "include"), (".h",))
    if not force and _newer(target, deps):
        return target
    cmd = [nvcc_path(), *NVCC_ARCH, "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
           "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
           "-I", os.path.join(ROOT, "include"), "-I", CSRC, *(extra or []), "-o", target, *cu, "-ldl"]
    _run(cmd, verbose)
    return target


def build_meshgen(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB, exist_ok=True)
    target = os.path.join(LIB, "libssam_meshgen.so")
    src = os.path.join(PKG, "meshgen", "meshgen.cpp")
    if not force and _newer(target, [src, os.path.join(PKG, "meshgen", "foam_io.inc")]):
        return target
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", target, src], verbose)
    return target


def build_host(force: bool = False, verbose: bool = False) -> str | None:
    """C++ host-side mirror of the reference interfaces + the call-sequence driver."""
    hostdir = os.path.join(PKG, "host")
    srcs = _walk(hostdir, (".cpp",))
    if not srcs:
        return None
    target = os.path.join(LIB, "ssam_case_driver")
    deps = srcs + _walk(hostdir, (".H", ".h", ".hpp")) + _walk(os.path.join(ROOT, "include"), (".h",))
    if not force and _newer(target, deps):
        return target
    _run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", hostdir, "-o", target, *srcs,
          "-L", LIB, "-lssam_gpu", "-lssam_meshgen", "-Wl,-rpath,$ORIGIN", "-ldl", "-lpthread"], verbose)
    return target


def build_oracle(force: bool = False, verbose: bool = False) -> str:
    odir = os.path.join(ROOT, "oracle")
    cmd = ["make", "-C", odir] + (["-B"] if force else [])
    _run(cmd, verbose)
    return os.path.join(odir, "liboracle.so")


def build_oracle_ref(verbose: bool = False) -> str | None:
    """oracle/_ref/libssam_ref.so: the reference's own model sources against oracle/ofstub (test infrastructure).
    Built only where the reference tree is mounted; elsewhere a prebuilt copy (it travels with the snapshot) is kept."""
    odir = os.path.join(ROOT, "oracle")
    ref = os.environ.get("SSAM_REFERENCE_TREE", "/root/reference")
    so = os.path.join(odir, "_ref", "libssam_ref.so")
    if os.path.isdir(ref):
        _run(["make", "-C", odir, "ref", f"REF={ref}"], verbose)
        if os.path.exists(os.path.join(LIB, "libssam_gpu.so")):
            # the drop-in demonstration: reference mixture / friction sources + integration/openfoam GPU models
            _run(["make", "-C", odir, "shim", f"REF={ref}"], verbose)
            # ... and with the mixture's own translation unit GPU-backed as well (device-resident fields)
            _run(["make", "-C", odir, "shim2", f"REF={ref}"], verbose)
    return so if os.path.exists(so) else None


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_meshgen(force, verbose)
    build_cuda(force, verbose)
    build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    build_oracle(force="--force" in sys.argv, verbose=True)
