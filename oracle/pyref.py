"""ctypes wrapper of oracle/_ref/libssam_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the reference's own model sources (viscoPlasticModels, frictionModels,
incompressibleTwoPhaseThermalMixture), compiled unmodified from /root/reference against the stand-in OpenFOAM
headers of oracle/ofstub/ (oracle/Makefile, target `ref`).  It exists only where the reference tree exists or
where a prebuilt copy travelled with the repository snapshot; `available()` says which.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from solidstateamfoam_b200 import params as P

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libssam_ref.so")
_SO_GPU = os.path.join(_HERE, "_ref", "libssam_ref_gpu.so")  # same harness, GPU-backed viscosity models (integration/openfoam)
_SO_GPU2 = os.path.join(_HERE, "_ref", "libssam_ref_gpu2.so")  # ... and the GPU-backed mixture translation unit
REFERENCE_TREE = os.environ.get("SSAM_REFERENCE_TREE", "/root/reference")
_lib = None
_lib_gpu = None
_lib_gpu2 = None


def build_shim() -> bool:
    """oracle/_ref/libssam_ref_gpu.so (Makefile target `shim`); needs libssam_gpu.so to be built already."""
    if os.path.isdir(REFERENCE_TREE) and not os.environ.get("SSAM_REF_NOBUILD"):
        r = subprocess.run(["make", "-C", _HERE, "shim", f"REF={REFERENCE_TREE}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the OpenFOAM-shim demo failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return os.path.exists(_SO_GPU)


def build_shim2() -> bool:
    """oracle/_ref/libssam_ref_gpu2.so (Makefile target `shim2`): the mixture's own translation unit GPU-backed too"""
    if os.path.isdir(REFERENCE_TREE) and not os.environ.get("SSAM_REF_NOBUILD"):
        r = subprocess.run(["make", "-C", _HERE, "shim2", f"REF={REFERENCE_TREE}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the GPU-mixture shim failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return os.path.exists(_SO_GPU2)


def build() -> bool:
    """(Re)build from the reference tree when it is present; keep a prebuilt library otherwise."""
    if os.path.isdir(REFERENCE_TREE) and not os.environ.get("SSAM_REF_NOBUILD"):
        r = subprocess.run(["make", "-C", _HERE, "ref", f"REF={REFERENCE_TREE}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building oracle/_ref failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return os.path.exists(_SO)


def available() -> bool:
    try:
        return build()
    except RuntimeError:
        return False


def lib(gpu=False):
    global _lib, _lib_gpu, _lib_gpu2
    if gpu == 2:
        if _lib_gpu2 is None:
            if not build_shim2():
                raise RuntimeError("oracle/_ref/libssam_ref_gpu2.so is absent and there is no reference tree to build it from")
            _lib_gpu2 = _bind(C.CDLL(_SO_GPU2))
        return _lib_gpu2
    if gpu:
        if _lib_gpu is None:
            if not build_shim():
                raise RuntimeError("oracle/_ref/libssam_ref_gpu.so is absent and there is no reference tree to build it from")
            _lib_gpu = _bind(C.CDLL(_SO_GPU))
        return _lib_gpu
    if _lib is None:
        if not build():
            raise RuntimeError("oracle/_ref/libssam_ref.so is absent and there is no reference tree to build it from")
        _lib = _bind(C.CDLL(_SO))
    return _lib


def _bind(L):
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    L.ref_create.restype = vp
    L.ref_create.argtypes = [C.POINTER(P.MeshDesc), C.POINTER(C.c_char_p), C.c_char_p]
    L.ref_destroy.argtypes = [vp]
    L.ref_last_error.restype = C.c_char_p
    L.ref_last_error.argtypes = [vp]
    L.ref_info.restype = C.c_char_p
    L.ref_info.argtypes = [vp]
    L.ref_set_scalar.argtypes = [vp, C.c_char_p, dp, C.c_int]
    L.ref_construct.argtypes = [vp, C.c_int]
    L.ref_thermo_correct.argtypes = [vp]
    L.ref_friction_correct.argtypes = [vp]
    L.ref_eval.argtypes = [vp, C.c_char_p, dp]
    L.ref_bc_rotational_slip.argtypes = [vp, C.c_char_p, dp]
    L.ref_calculate_strain.argtypes = [vp, C.c_double, dp, dp, dp, dp, dp, dp]
    L.ref_set_U.argtypes = [vp, dp, C.POINTER(C.c_int32)]
    if hasattr(L, "ref_gpu_traffic"):
        L.ref_gpu_traffic.argtypes = [vp, dp, C.c_int]
        L.ref_gpu_contexts.restype = C.c_int
    if hasattr(L, "ref_teqn_update"):
        L.ref_teqn_update.argtypes = [vp, C.c_int]
    return L


class ReferenceError_(RuntimeError):
    pass


class ReferenceCase:
    """The reference's mixture + friction model on one mesh, reading `<case_dir>/constant/{transport,friction}Properties`."""

    def __init__(self, mesh, case_dir: str, gpu: bool = False):
        self.mesh = mesh
        self.L = lib(gpu)
        self._desc, self._keep = P.mesh_desc(mesh)
        names = (C.c_char_p * mesh.nPatches)(*[n.encode() for n in mesh.patchName])
        self.h = C.c_void_p(self.L.ref_create(C.byref(self._desc), names, case_dir.encode()))
        self.nTot = mesh.nCells + mesh.nBoundaryFaces

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != 0:
            raise ReferenceError_(self.L.ref_last_error(self.h).decode())

    def set(self, name: str, values, as_file: bool = False):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.nTot,), (name, v.shape)
        self._chk(self.L.ref_set_scalar(self.h, name.encode(), v.ctypes.data_as(C.POINTER(C.c_double)), 1 if as_file else 0))

    def construct(self, with_friction: bool = True):
        self._chk(self.L.ref_construct(self.h, 1 if with_friction else 0))

    def thermo_correct(self):
        self._chk(self.L.ref_thermo_correct(self.h))

    def friction_correct(self):
        self._chk(self.L.ref_friction_correct(self.h))

    def eval(self, what: str) -> np.ndarray:
        n = self.mesh.nFaces if what in ("muf", "nuf") else self.nTot
        out = np.zeros(n)
        self._chk(self.L.ref_eval(self.h, what.encode(), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def bc_rotational_slip(self, patch_name: str) -> np.ndarray:
        """updateCoeffs() of the *RotationalSlip patch field described in <case_dir>/0/U -> U_b [n, 3]"""
        p = self.mesh.patch_id(patch_name)
        out = np.zeros((int(self.mesh.patchSize[p]), 3))
        self._chk(self.L.ref_bc_rotational_slip(self.h, patch_name.encode(), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def calculate_strain(self, deltaT: float, gradU, p_rgh, epsilonEq):
        """fluid/calculateStrain.H verbatim (its fvm solve is a no-op) -> dict(Z, sigmaf, epsilonEq, sigmavm)"""
        dp = C.POINTER(C.c_double)
        g = np.ascontiguousarray(gradU, dtype=np.float64).reshape(-1)
        assert g.size == 9 * self.nTot
        pr = np.ascontiguousarray(p_rgh, dtype=np.float64)
        eq = np.array(epsilonEq, dtype=np.float64)
        Z, sf, vm = np.zeros(self.nTot), np.zeros(self.nTot), np.zeros(self.nTot)
        self._chk(self.L.ref_calculate_strain(self.h, deltaT, g.ctypes.data_as(dp), pr.ctypes.data_as(dp),
                                             eq.ctypes.data_as(dp), Z.ctypes.data_as(dp), sf.ctypes.data_as(dp),
                                             vm.ctypes.data_as(dp)))
        return {"Z": Z, "sigmaf": sf, "epsilonEq": eq, "sigmavm": vm}

    def set_U(self, U, patch_kinds=None):
        """U incl. boundary values [nTot, 3] and the snGrad kind of each patch (GPU-backed models only)"""
        u = np.ascontiguousarray(U, dtype=np.float64).reshape(-1)
        assert u.size == 3 * self.nTot
        k = None
        if patch_kinds is not None:
            self._kinds = np.ascontiguousarray(patch_kinds, dtype=np.int32)
            k = self._kinds.ctypes.data_as(C.POINTER(C.c_int32))
        self._chk(self.L.ref_set_U(self.h, u.ctypes.data_as(C.POINTER(C.c_double)), k))

    def info(self) -> str:
        return self.L.ref_info(self.h).decode()

    def teqn_update(self, fused: int = 0):
        """fluid/TEqn.H:3-15 through integration/openfoam/ssamTEqnUpdate.H (GPU-mixture shim only)"""
        self._chk(self.L.ref_teqn_update(self.h, fused))

    def gpu_traffic(self, reset: bool = False):
        """(bytes host->device, bytes device->host) the shim moved for this mesh since the last reset"""
        out = np.zeros(2)
        self.L.ref_gpu_traffic(self.h, out.ctypes.data_as(C.POINTER(C.c_double)), 1 if reset else 0)
        return int(out[0]), int(out[1])

    def gpu_contexts(self) -> int:
        return int(self.L.ref_gpu_contexts())
