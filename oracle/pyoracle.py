"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package.  See oracle/oracle.cpp for what the oracle is (a CPU restatement,
parity unpinned by the reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from solidstateamfoam_b200 import params as P

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

FAITHFUL, FUSED = 0, 1


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(_HERE, "liboracle.so")
        src = os.path.join(_HERE, "oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(so)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(P.MeshDesc), ip]
        L.orc_destroy.argtypes = [vp]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [vp]
        L.orc_field.restype = dp
        L.orc_field.argtypes = [vp, C.c_int, C.c_int]
        L.orc_ntot.argtypes = [vp]
        L.orc_bc_rotational_slip.argtypes = [vp, C.POINTER(P.RotSlipParams)]
        L.orc_grad_u.argtypes = [vp, C.c_int]
        L.orc_strain_rate.argtypes = [vp, C.c_double, C.c_int, C.c_int]
        L.orc_strain_from_grad.argtypes = [vp, C.c_double, C.c_int, C.c_int]
        L.orc_viscosity_correct.argtypes = [vp, C.c_int, C.POINTER(P.ViscosityParams), C.c_int]
        for n in ("calc_nu", "mu", "rho", "cp", "k"):
            getattr(L, "orc_mixture_" + n).argtypes = [vp, C.POINTER(P.MixtureParams), C.c_int]
        L.orc_solver_rho_rhocp.argtypes = [vp, C.POINTER(P.MixtureParams), C.c_int]
        L.orc_stress_strain.argtypes = [vp, C.POINTER(P.StressStrainParams), C.c_int]
        L.orc_face_interpolate.argtypes = [vp, C.c_int, dp]
        L.orc_mixture_face_visc.argtypes = [vp, C.POINTER(P.MixtureParams), C.c_int]
        L.orc_face_field.restype = dp
        L.orc_face_field.argtypes = [vp, C.c_int]
        L.orc_field_min_max.argtypes = [C.POINTER(vp), C.c_int, C.c_int, dp, dp]
        L.orc_div_dev_rho_reff_explicit.argtypes = [vp]
        L.orc_interface_correct.argtypes = [vp, C.c_double, C.c_int]
        L.orc_interface_nhatf.restype = dp
        L.orc_interface_nhatf.argtypes = [vp]
        L.orc_set_phi.argtypes = [vp, dp]
        L.orc_alpha_compression_coeff.argtypes = [vp, C.c_double, C.c_double, C.c_double]
        L.orc_phic.restype = dp
        L.orc_phic.argtypes = [vp]
        L.orc_exchange_cell_centres.argtypes = [C.POINTER(vp), C.c_int]
        L.orc_interface_delta_n.restype = C.c_double
        L.orc_interface_delta_n.argtypes = [C.POINTER(vp), C.c_int]
        L.orc_friction_patch_sums.argtypes = [vp, C.c_int, dp]
        L.orc_friction_correct.argtypes = [vp, C.POINTER(P.StickParams), dp, C.c_int]
        L.orc_friction_patch_centre.argtypes = [vp, dp, dp]
        L.orc_update.argtypes = [vp, C.POINTER(P.UpdateParams), dp, C.c_int]
        L.orc_addressing_size.restype = C.c_int64
        L.orc_addressing_size.argtypes = [vp, C.c_int]
        L.orc_addressing_get.argtypes = [vp, C.c_int, ip]
        L.orc_halo_exchange.argtypes = [C.POINTER(vp), C.c_int, C.c_int]
        L.orc_update_multi.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(P.UpdateParams), C.c_int]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"oracle status {status}: {msg}")
        self.status = status


class OracleCase:
    """One mesh (partition) + its fields on the CPU."""

    def __init__(self, mesh, patch_u_bc=None):
        self.mesh = mesh
        self._desc, self._keep = P.mesh_desc(mesh)
        bc = None
        if patch_u_bc is not None:
            self._bc = np.ascontiguousarray(patch_u_bc, dtype=np.int32)
            bc = self._bc.ctypes.data_as(C.POINTER(C.c_int32))
        self.h = C.c_void_p(lib().orc_create(C.byref(self._desc), bc))
        self.nTot = mesh.nCells + mesh.nBoundaryFaces

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st != 0:
            raise OracleError(st, lib().orc_last_error(self.h).decode())

    def field(self, f, mark=False) -> np.ndarray:
        """numpy view [nTot, ncomp] (or [nTot]) into the oracle's storage: internal then boundary."""
        nc = P.ncomp(f)
        ptr = lib().orc_field(self.h, f, 1 if mark else 0)
        a = np.ctypeslib.as_array(ptr, shape=(self.nTot * nc,))
        return a.reshape(self.nTot, nc) if nc > 1 else a

    def set(self, f, internal, boundary=None):
        a = self.field(f, mark=True)
        n = self.mesh.nCells
        a[:n] = internal
        if boundary is not None:
            a[n:] = boundary

    def internal(self, f):
        return self.field(f)[: self.mesh.nCells].copy()

    def boundary(self, f):
        return self.field(f)[self.mesh.nCells:].copy()

    # staged calls -------------------------------------------------------------------------------
    def bc_rotational_slip(self, p):
        self._chk(lib().orc_bc_rotational_slip(self.h, C.byref(p)))

    def grad_u(self, form=FAITHFUL):
        self._chk(lib().orc_grad_u(self.h, form))

    def strain_rate(self, factor, out_field, form=FAITHFUL):
        self._chk(lib().orc_strain_rate(self.h, factor, out_field, form))

    def strain_from_grad(self, factor, out_field, form=FAITHFUL):
        self._chk(lib().orc_strain_from_grad(self.h, factor, out_field, form))

    def viscosity_correct(self, phase, p, form=FAITHFUL):
        self._chk(lib().orc_viscosity_correct(self.h, phase, C.byref(p), form))

    def mixture(self, what, p, form=FAITHFUL):
        self._chk(getattr(lib(), "orc_mixture_" + what)(self.h, C.byref(p), form))

    def solver_rho_rhocp(self, p, form=FAITHFUL):
        self._chk(lib().orc_solver_rho_rhocp(self.h, C.byref(p), form))

    def stress_strain(self, p, form=FAITHFUL):
        self._chk(lib().orc_stress_strain(self.h, C.byref(p), form))

    def div_dev_rho_reff_explicit(self):
        """-fvc::div((rho*nuEff)*dev2(T(grad U))) from GRADU, RHO_SOLVER, NU"""
        self._chk(lib().orc_div_dev_rho_reff_explicit(self.h))

    def interface_correct(self, deltaN, stage=3):
        """interfaceProperties::correct(); stage 1 = cell gradient, 2 = the rest (halo steps go in between)."""
        self._chk(lib().orc_interface_correct(self.h, deltaN, stage))

    def alpha_compression_coeff(self, phi, cAlpha, icAlpha, scAlpha):
        """multiRegionVoF/alphaEqn.H:58-87: phic from the face flux phi [nFaces], U and GRADU -> [nFaces]"""
        ph = np.ascontiguousarray(phi, dtype=np.float64)
        assert ph.shape == (self.mesh.nFaces,)
        lib().orc_set_phi(self.h, ph.ctypes.data_as(C.POINTER(C.c_double)))
        self._chk(lib().orc_alpha_compression_coeff(self.h, cAlpha, icAlpha, scAlpha))
        return np.ctypeslib.as_array(lib().orc_phic(self.h), shape=(self.mesh.nFaces,)).copy()

    def nhatf(self):
        return np.ctypeslib.as_array(lib().orc_interface_nhatf(self.h), shape=(self.mesh.nFaces,)).copy()

    def face_interpolate(self, field):
        """fvc::interpolate(field), linear scheme -> [nFaces]"""
        out = np.zeros(self.mesh.nFaces)
        self._chk(lib().orc_face_interpolate(self.h, field, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def mixture_face_visc(self, p, nuf=False):
        """muf() / nuf(): returns the surfaceScalarField (internal faces then boundary faces)."""
        self._chk(lib().orc_mixture_face_visc(self.h, C.byref(p), 1 if nuf else 0))
        which = P.FF_NUF if nuf else P.FF_MUF
        return np.ctypeslib.as_array(lib().orc_face_field(self.h, which), shape=(self.mesh.nFaces,)).copy()

    def friction_patch_sums(self, patch):
        out = np.zeros(4)
        lib().orc_friction_patch_sums(self.h, patch, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def friction_correct(self, p, sums=None, form=FAITHFUL):
        s = None
        if sums is not None:
            s = np.ascontiguousarray(sums, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        self._chk(lib().orc_friction_correct(self.h, C.byref(p), s, form))

    def friction_patch_centre(self):
        c = np.zeros(3)
        a = C.c_double()
        lib().orc_friction_patch_centre(self.h, c.ctypes.data_as(C.POINTER(C.c_double)), C.byref(a))
        return c, a.value

    def update(self, p, sums=None, form=FAITHFUL):
        s = None
        if sums is not None:
            s = np.ascontiguousarray(sums, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double))
        self._chk(lib().orc_update(self.h, C.byref(p), s, form))

    def addressing(self, which) -> np.ndarray:
        n = lib().orc_addressing_size(self.h, which)
        out = np.zeros(max(n, 0), dtype=np.int32)
        lib().orc_addressing_get(self.h, which, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out


def halo_exchange(cases, field):
    arr = (C.c_void_p * len(cases))(*[c.h for c in cases])
    st = lib().orc_halo_exchange(arr, len(cases), field)
    if st:
        raise OracleError(st, "halo exchange: processor patches do not match")


def exchange_cell_centres(cases):
    """processorPolyPatch::neighbFaceCellCentres() for every rank (needed by mesh.delta() on processor faces)"""
    arr = (C.c_void_p * len(cases))(*[c.h for c in cases])
    st = lib().orc_exchange_cell_centres(arr, len(cases))
    if st:
        raise OracleError(st, "cell-centre exchange: processor patches do not match")


def field_min_max(cases, field):
    arr = (C.c_void_p * len(cases))(*[c.h for c in cases])
    mn, mx = C.c_double(), C.c_double()
    lib().orc_field_min_max(arr, len(cases), field, C.byref(mn), C.byref(mx))
    return mn.value, mx.value


def interface_delta_n(cases):
    arr = (C.c_void_p * len(cases))(*[c.h for c in cases])
    return lib().orc_interface_delta_n(arr, len(cases))


def interface_correct_multi(cases, deltaN):
    """Decomposed interfaceProperties::correct(): processor-patch evaluations between the stages."""
    halo_exchange(cases, P.F_ALPHA1)
    for c in cases:
        c.interface_correct(deltaN, 1)
    halo_exchange(cases, P.F_GRADALPHA)
    for c in cases:
        c.interface_correct(deltaN, 2)
    halo_exchange(cases, P.F_CURVATURE)


def update_multi(cases, p, form=FAITHFUL):
    """All ranks' update, one thread per rank, in-memory halo copies (the timed CPU baseline)."""
    arr = (C.c_void_p * len(cases))(*[c.h for c in cases])
    st = lib().orc_update_multi(arr, len(cases), C.byref(p), form)
    if st:
        raise OracleError(st, "; ".join(lib().orc_last_error(c.h).decode() for c in cases))
