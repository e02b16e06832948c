"""ctypes mirrors of the POD structs and enums in include/ssam_gpu.h (no native code loaded here)."""
from __future__ import annotations

import ctypes as C

# status codes
OK, ERR_INVALID, ERR_CUDA, ERR_UNKNOWN_MODEL, ERR_MISSING_FIELD, ERR_COMM, ERR_UNITS = range(7)

PATCH_GENERIC, PATCH_PROCESSOR, PATCH_EMPTY = 0, 1, 2
BC_VALUE, BC_ZERO_GRADIENT = 0, 1

(F_U, F_T, F_ALPHA1, F_P, F_EPSDOT, F_NU1, F_NU2, F_SIGMAF, F_SIGMAY, F_NU, F_MUEFF, F_RHO, F_CP, F_K, F_QVISC,
 F_QFRIC, F_GRADU, F_STRAINRATE, F_RHO_SOLVER, F_RHOCP) = range(20)
F_PRGH, F_Z, F_SIGMAF_PP, F_DEPSEQ, F_EPSEQ, F_SIGMAVM = range(20, 26)
F_GRADALPHA, F_CURVATURE, F_DIVDEVRHOREFF = 26, 27, 28
F_K1, F_K2 = 29, 30
F_COUNT = 31
FF_NHATF, FF_MUF, FF_NUF, FF_KF, FF_PHI, FF_PHIC = range(6)
FIELD_NAMES = ["U", "temp", "alpha1", "p", "epsilonDotEq", "nu1", "nu2", "sigmaf", "sigmay", "nu", "muEff", "rho",
               "cp", "k", "qvisc", "qfric", "gradU", "strainRate", "rhoSolver", "rhoCp", "p_rgh", "Z", "sigmafPP",
               "dEpsilonEq", "epsilonEq", "sigmavm", "gradAlpha", "interfaceProperties:K", "divDevRhoReffExplicit", "k1", "k2"]


def ncomp(field: int) -> int:
    return 3 if field in (F_U, F_GRADALPHA, F_DIVDEVRHOREFF) else (9 if field == F_GRADU else 1)


VISC_NEWTONIAN, VISC_SHEPPARD_WRIGHT, VISC_FKS, VISC_NORTON_HOFF = range(4)
K_CONSTANT, K_CUBIC = 0, 1
UNITS_KELVIN, UNITS_CELSIUS = 0, 1
STICK_CONSTANT, STICK_EXPONENTIAL = 0, 1
ROTSLIP_CONSTANT, ROTSLIP_EXPONENTIAL = 0, 1

(ADDR_OWNER_START, ADDR_EXTRA_START, ADDR_EXTRA_FACE, ADDR_EXTRA_OTHER, ADDR_CELL_FACES_START, ADDR_CELL_FACES,
 ADDR_HALO_SEND_CELLS, ADDR_HALO_PATCH_OFFSETS, ADDR_FRIC_CELLS, ADDR_FACE_COLOUR) = range(10)
ADDR_CELL_ORDER, ADDR_FACE_ORDER = 10, 11
LAYOUT_CALLER, LAYOUT_AUTO, LAYOUT_COMPACT = range(3)
ADDR_NAMES = ["ownerStart", "extraStart", "extraFace", "extraOther", "cellFacesStart", "cellFaces", "haloSendCells",
              "haloPatchOffsets", "fricCells", "faceColour", "cellOrder", "faceOrder"]

UNIQUE_ID_BYTES = 128

# OpenFOAM constants
SMALL, VSMALL, ROOTVSMALL, ROOTVGREAT = 1e-15, 1e-300, 1e-150, 1e150

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class MeshDesc(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nInternalFaces", C.c_int32), ("nFaces", C.c_int32), ("nPatches", C.c_int32),
                ("owner", _ip), ("neighbour", _ip), ("patchStart", _ip), ("patchSize", _ip), ("patchKind", _ip),
                ("patchNbrRank", _ip), ("Sf", _dp), ("weights", _dp), ("V", _dp), ("C", _dp), ("Cf_b", _dp),
                ("magSf_b", _dp), ("deltaCoeffs_b", _dp)]


class ViscosityParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("_pad", C.c_int32), ("nu", C.c_double),
                ("sigmaR", C.c_double), ("Q", C.c_double), ("R", C.c_double), ("A", C.c_double), ("n", C.c_double),
                ("a1", C.c_double), ("a2", C.c_double), ("b1", C.c_double), ("b2", C.c_double), ("b3", C.c_double),
                ("e0", C.c_double), ("c1", C.c_double), ("c2", C.c_double), ("Tm", C.c_double), ("Ta", C.c_double),
                ("mu0", C.c_double), ("m", C.c_double), ("strainRateEffMin", C.c_double),
                ("rho", C.c_double), ("nuMin", C.c_double), ("nuMax", C.c_double), ("eDotMin", C.c_double)]


class Conductivity(C.Structure):
    _fields_ = [("kModel", C.c_int32), ("units", C.c_int32), ("k", C.c_double), ("k0", C.c_double),
                ("k1", C.c_double), ("k2", C.c_double), ("k3", C.c_double), ("Tmin", C.c_double), ("Tmax", C.c_double)]


class MixtureParams(C.Structure):
    _fields_ = [("rho1", C.c_double), ("rho2", C.c_double), ("cp1", C.c_double), ("cp2", C.c_double),
                ("k1", Conductivity), ("k2", Conductivity)]


class StickParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("fricPatch", C.c_int32), ("betaTQ", C.c_double), ("omega", C.c_double * 3),
                ("delta", C.c_double), ("muf", C.c_double), ("omega0", C.c_double), ("delta0", C.c_double),
                ("Rs", C.c_double), ("mu0", C.c_double), ("lambda_", C.c_double)]


class RotSlipParams(C.Structure):
    _fields_ = [("model", C.c_int32), ("patch", C.c_int32), ("omega", C.c_double * 3), ("Vt", C.c_double * 3),
                ("delta", C.c_double), ("omega0", C.c_double), ("delta0", C.c_double), ("R0", C.c_double)]


class StressStrainParams(C.Structure):
    _fields_ = [("Q", C.c_double), ("R", C.c_double), ("sigmaR", C.c_double), ("A", C.c_double), ("n", C.c_double),
                ("deltaT", C.c_double)]


class UpdateParams(C.Structure):
    _fields_ = [("visc1", ViscosityParams), ("visc2", ViscosityParams), ("mix", MixtureParams), ("stick", StickParams)]


class HostIO(C.Structure):
    _fields_ = [("U_i", _dp), ("U_b", _dp), ("T_i", _dp), ("T_b", _dp), ("alpha1_i", _dp), ("alpha1_b", _dp),
                ("p_i", _dp), ("p_b", _dp), ("nOut", C.c_int32), ("outFields", _ip),
                ("out_i", C.POINTER(_dp)), ("out_b", C.POINTER(_dp))]


# ---- constructors with the reference's dictionary defaults ---------------------------------------
def newtonian(nu: float, rho: float) -> ViscosityParams:
    return ViscosityParams(model=VISC_NEWTONIAN, nu=nu, rho=rho, nuMin=ROOTVSMALL, nuMax=ROOTVGREAT,
                           eDotMin=ROOTVSMALL)


def sheppard_wright(sigmaR, Q, R, A, n, rho, nuMin=ROOTVSMALL, nuMax=ROOTVGREAT, eDotMin=ROOTVSMALL):
    """SheppardWrightCoeffs; limiter defaults from SheppardWright.C:104-106."""
    return ViscosityParams(model=VISC_SHEPPARD_WRIGHT, sigmaR=sigmaR, Q=Q, R=R, A=A, n=n, rho=rho, nuMin=nuMin,
                           nuMax=nuMax, eDotMin=eDotMin)


def fks(a1, a2, b1, b2, b3, e0, c1, c2, Tm, Ta, rho, nuMin=ROOTVSMALL, nuMax=ROOTVGREAT, eDotMin=ROOTVSMALL):
    """FKSCoeffs (FKS.C:199-213); limiter defaults FKS.C:165-167."""
    return ViscosityParams(model=VISC_FKS, a1=a1, a2=a2, b1=b1, b2=b2, b3=b3, e0=e0, c1=c1, c2=c2, Tm=Tm, Ta=Ta,
                           rho=rho, nuMin=nuMin, nuMax=nuMax, eDotMin=eDotMin)


def norton_hoff(mu0, m, rho, nuMin, nuMax, strainRateEffMin):
    """NortonHoffCoeffs: all required (NortonHoff.C:82-87)."""
    return ViscosityParams(model=VISC_NORTON_HOFF, mu0=mu0, m=m, rho=rho, nuMin=nuMin, nuMax=nuMax,
                           strainRateEffMin=strainRateEffMin, eDotMin=ROOTVSMALL)


def k_constant(k: float) -> Conductivity:
    return Conductivity(kModel=K_CONSTANT, units=UNITS_KELVIN, k=k)


def k_cubic(k0, k1, k2, k3, Tmin, Tmax, units="Kelvin") -> Conductivity:
    u = {"Kelvin": UNITS_KELVIN, "Celsius": UNITS_CELSIUS}.get(units, 99)
    return Conductivity(kModel=K_CUBIC, units=u, k0=k0, k1=k1, k2=k2, k3=k3, Tmin=Tmin, Tmax=Tmax)


def mixture(rho1, rho2, cp1, cp2, k1: Conductivity, k2: Conductivity) -> MixtureParams:
    return MixtureParams(rho1=rho1, rho2=rho2, cp1=cp1, cp2=cp2, k1=k1, k2=k2)


def constant_slip_stick(fricPatch, betaTQ, delta, omega, muf) -> StickParams:
    return StickParams(model=STICK_CONSTANT, fricPatch=fricPatch, betaTQ=betaTQ, omega=(C.c_double * 3)(*omega),
                       delta=delta, muf=muf)


def exponential_slip_stick(fricPatch, betaTQ, omega, omega0, delta0, Rs, mu0, lambda_) -> StickParams:
    return StickParams(model=STICK_EXPONENTIAL, fricPatch=fricPatch, betaTQ=betaTQ, omega=(C.c_double * 3)(*omega),
                       omega0=omega0, delta0=delta0, Rs=Rs, mu0=mu0, lambda_=lambda_)


def no_friction_patch(betaTQ) -> StickParams:
    """qvisc only (NortonHoff registers no sigmay)."""
    return StickParams(model=STICK_CONSTANT, fricPatch=-1, betaTQ=betaTQ, omega=(C.c_double * 3)(0, 0, 0))


def constant_rotational_slip(patch, omega, delta, Vt) -> RotSlipParams:
    return RotSlipParams(model=ROTSLIP_CONSTANT, patch=patch, omega=(C.c_double * 3)(*omega),
                         Vt=(C.c_double * 3)(*Vt), delta=delta)


def exponential_rotational_slip(patch, omega, omega0, delta0, R0, Vt) -> RotSlipParams:
    return RotSlipParams(model=ROTSLIP_EXPONENTIAL, patch=patch, omega=(C.c_double * 3)(*omega),
                         Vt=(C.c_double * 3)(*Vt), omega0=omega0, delta0=delta0, R0=R0)


def mesh_desc(m) -> tuple[MeshDesc, list]:
    """MeshDesc for a mesh.Mesh; returns (desc, keepalive list of the numpy arrays it points into)."""
    import numpy as np

    nif = m.nInternalFaces
    arrs = dict(
        owner=np.ascontiguousarray(m.owner, dtype=np.int32),
        neighbour=np.ascontiguousarray(m.neighbour, dtype=np.int32),
        patchStart=np.ascontiguousarray(m.patchStart, dtype=np.int32),
        patchSize=np.ascontiguousarray(m.patchSize, dtype=np.int32),
        patchKind=np.ascontiguousarray(m.patchKind, dtype=np.int32),
        patchNbrRank=np.ascontiguousarray(m.patchNbrRank, dtype=np.int32),
        Sf=np.ascontiguousarray(m.Sf, dtype=np.float64),
        weights=np.ascontiguousarray(m.weights, dtype=np.float64),
        V=np.ascontiguousarray(m.V, dtype=np.float64),
        C=np.ascontiguousarray(m.C, dtype=np.float64),
        Cf_b=np.ascontiguousarray(m.Cf[nif:], dtype=np.float64),
        magSf_b=np.ascontiguousarray(m.magSf[nif:], dtype=np.float64),
        deltaCoeffs_b=np.ascontiguousarray(m.deltaCoeffs[nif:], dtype=np.float64),
    )
    d = MeshDesc(nCells=m.nCells, nInternalFaces=nif, nFaces=m.nFaces, nPatches=m.nPatches)
    for k, a in arrs.items():
        setattr(d, k, a.ctypes.data_as(_ip if a.dtype == np.int32 else _dp))
    return d, list(arrs.values())
