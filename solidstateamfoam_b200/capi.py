"""ctypes binding of lib/libssam_gpu.so (include/ssam_gpu.h) -- what an FFI host would bind.

There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
construction fails loudly (SsamError).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _build
from . import params as P

_lib = None

# every exported symbol of include/ssam_gpu.h (checked by tests/test_abi_symbols.py)
SYMBOLS = [
    "ssam_create", "ssam_destroy", "ssam_last_error", "ssam_sync", "ssam_set_stream", "ssam_set_blocking",
    "ssam_abi_version", "ssam_launch_count", "ssam_set_cell_kernel", "ssam_set_cell_layout", "ssam_cell_layout_is_compact", "ssam_compact_cell_order", "ssam_profile_enable", "ssam_profile_get", "ssam_mesh_set", "ssam_patch_u_bc_set", "ssam_addressing_size",
    "ssam_addressing_get", "ssam_field_upload", "ssam_field_download", "ssam_field_device_ptr", "ssam_field_is_set",
    "ssam_bc_rotational_slip", "ssam_grad_u", "ssam_strain_rate", "ssam_viscosity_correct", "ssam_mixture_calc_nu",
    "ssam_thermo_correct", "ssam_mixture_mu", "ssam_mixture_rho", "ssam_mixture_cp", "ssam_mixture_k", "ssam_mixture_k_phase",
    "ssam_solver_rho_rhocp", "ssam_stress_strain", "ssam_div_dev_rho_reff_explicit", "ssam_interface_correct",
    "ssam_interface_nhatf_download", "ssam_interface_nhatf_device_ptr", "ssam_interface_delta_n", "ssam_face_field_download", "ssam_face_field_upload", "ssam_alpha_compression_coeff",
    "ssam_face_field_device_ptr", "ssam_face_interpolate", "ssam_mixture_muf", "ssam_mixture_nuf", "ssam_field_min_max", "ssam_friction_correct", "ssam_friction_patch_centre", "ssam_update_fused", "ssam_prepare_fused",
    "ssam_update_fused_host", "ssam_set_host_chunks", "ssam_debug_math", "ssam_measure_fp64_peak", "ssam_comm_unique_id", "ssam_comm_init", "ssam_comm_init_local", "ssam_halo_exchange", "ssam_halo_transport", "ssam_halo_bytes_last_update",
]


class SsamError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"ssam status {status}: {msg}")
        self.status = status
        self.msg = msg


def lib_path() -> str:
    return os.path.join(_build.LIB, "libssam_gpu.so")


def load():
    """Load libssam_gpu.so (building it in-tree first if the .so is absent and nvcc is available)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        _build.build_cuda()
    if not os.path.exists(path):
        raise SsamError(P.ERR_CUDA, f"{path} is missing: the CUDA extension is required, there is no fallback")
    L = C.CDLL(path)
    vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    L.ssam_create.argtypes = [C.POINTER(vp), C.c_int]
    L.ssam_destroy.argtypes = [vp]
    L.ssam_last_error.restype = C.c_char_p
    L.ssam_last_error.argtypes = [vp]
    L.ssam_sync.argtypes = [vp]
    L.ssam_set_stream.argtypes = [vp, vp]
    L.ssam_set_blocking.argtypes = [vp, C.c_int]
    L.ssam_launch_count.restype = C.c_int64
    L.ssam_launch_count.argtypes = [vp]
    L.ssam_set_cell_kernel.argtypes = [vp, C.c_int]
    L.ssam_set_cell_layout.argtypes = [vp, C.c_int]
    L.ssam_cell_layout_is_compact.argtypes = [vp]
    L.ssam_compact_cell_order.argtypes = [C.POINTER(P.MeshDesc), ip, dp, dp, C.POINTER(C.c_int)]
    L.ssam_profile_enable.argtypes = [vp, C.c_int]
    L.ssam_profile_get.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.ssam_mesh_set.argtypes = [vp, C.POINTER(P.MeshDesc)]
    L.ssam_patch_u_bc_set.argtypes = [vp, ip]
    L.ssam_addressing_size.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.ssam_addressing_get.argtypes = [vp, C.c_int, ip, C.c_int64]
    L.ssam_field_upload.argtypes = [vp, C.c_int, dp, dp]
    L.ssam_field_download.argtypes = [vp, C.c_int, dp, dp]
    L.ssam_field_device_ptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
    L.ssam_field_is_set.argtypes = [vp, C.c_int]
    L.ssam_bc_rotational_slip.argtypes = [vp, C.POINTER(P.RotSlipParams)]
    L.ssam_grad_u.argtypes = [vp]
    L.ssam_strain_rate.argtypes = [vp, C.c_double, C.c_int]
    L.ssam_viscosity_correct.argtypes = [vp, C.c_int, C.POINTER(P.ViscosityParams)]
    L.ssam_thermo_correct.argtypes = [vp, C.POINTER(P.ViscosityParams), C.POINTER(P.ViscosityParams),
                                      C.POINTER(P.MixtureParams)]
    for n in ("calc_nu", "mu", "rho", "cp", "k"):
        getattr(L, "ssam_mixture_" + n).argtypes = [vp, C.POINTER(P.MixtureParams)]
    L.ssam_solver_rho_rhocp.argtypes = [vp, C.POINTER(P.MixtureParams)]
    L.ssam_stress_strain.argtypes = [vp, C.POINTER(P.StressStrainParams)]
    L.ssam_div_dev_rho_reff_explicit.argtypes = [vp]
    L.ssam_interface_correct.argtypes = [vp, C.c_double, C.c_uint32]
    L.ssam_interface_nhatf_download.argtypes = [vp, dp, dp]
    L.ssam_interface_nhatf_device_ptr.restype = C.c_void_p
    L.ssam_interface_nhatf_device_ptr.argtypes = [vp]
    L.ssam_interface_delta_n.argtypes = [vp, dp]
    L.ssam_face_field_download.argtypes = [vp, C.c_int, dp, dp]
    L.ssam_face_field_device_ptr.restype = C.c_void_p
    L.ssam_face_field_device_ptr.argtypes = [vp, C.c_int]
    L.ssam_face_interpolate.argtypes = [vp, C.c_int, C.c_int]
    L.ssam_mixture_muf.argtypes = [vp, C.POINTER(P.MixtureParams)]
    L.ssam_mixture_nuf.argtypes = [vp, C.POINTER(P.MixtureParams)]
    L.ssam_field_min_max.argtypes = [vp, C.c_int, dp, dp]
    L.ssam_mixture_k_phase.argtypes = [vp, C.POINTER(P.MixtureParams), C.c_int]
    L.ssam_face_field_upload.argtypes = [vp, C.c_int, dp, dp]
    L.ssam_alpha_compression_coeff.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.ssam_friction_correct.argtypes = [vp, C.POINTER(P.StickParams)]
    L.ssam_friction_patch_centre.argtypes = [vp, dp, dp]
    L.ssam_update_fused.argtypes = [vp, C.POINTER(P.UpdateParams), C.c_int]
    L.ssam_prepare_fused.argtypes = [vp, C.POINTER(P.UpdateParams), C.c_int]
    L.ssam_update_fused_host.argtypes = [vp, C.POINTER(P.UpdateParams), C.POINTER(P.HostIO)]
    L.ssam_set_host_chunks.argtypes = [vp, C.c_int]
    L.ssam_debug_math.argtypes = [vp, C.c_int, dp, dp, dp, C.c_int64]
    L.ssam_measure_fp64_peak.argtypes = [vp, dp]
    L.ssam_comm_unique_id.argtypes = [vp]
    L.ssam_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
    L.ssam_comm_init_local.argtypes = [C.POINTER(vp), C.c_int]
    L.ssam_halo_exchange.argtypes = [vp, C.c_uint32]
    L.ssam_halo_transport.argtypes = [vp]
    L.ssam_halo_bytes_last_update.argtypes = [vp]
    L.ssam_halo_bytes_last_update.restype = C.c_int64
    _lib = L
    return L


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


FUSED_WRITE_GRAD = 1
FUSED_HALO_INPUTS = 2


def compact_cell_order(mesh):
    """host-only: (order, remote neighbours per cell before, after, would SSAM_LAYOUT_AUTO use it)"""
    L = load()
    desc, keep = P.mesh_desc(mesh)
    order = np.zeros(mesh.nCells, dtype=np.int32)
    a, b, use = C.c_double(), C.c_double(), C.c_int()
    st = L.ssam_compact_cell_order(C.byref(desc), order.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(a), C.byref(b),
                                   C.byref(use))
    if st != 0:
        raise SsamError(st, L.ssam_last_error(None).decode())
    return order, a.value, b.value, bool(use.value)


class Context:
    """One rank / one GPU.  Thin, 1:1 with the C entry points."""

    def __init__(self, device: int = 0):
        L = load()
        self.L = L
        h = C.c_void_p()
        st = L.ssam_create(C.byref(h), device)
        if st != 0:
            raise SsamError(st, L.ssam_last_error(None).decode())
        self.h = h
        self.mesh = None
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.ssam_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _chk(self, st):
        if st != 0:
            raise SsamError(st, self.L.ssam_last_error(self.h).decode())

    # lifecycle ------------------------------------------------------------------------------------
    def sync(self):
        self._chk(self.L.ssam_sync(self.h))

    def set_stream(self, cuda_stream_ptr: int | None):
        self._chk(self.L.ssam_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def set_blocking(self, blocking: bool):
        self._chk(self.L.ssam_set_blocking(self.h, 1 if blocking else 0))

    def launch_count(self) -> int:
        return int(self.L.ssam_launch_count(self.h))

    def halo_transport(self) -> str:
        return {0: "none", 1: "peer-memory", 2: "nccl"}[self.L.ssam_halo_transport(self.h)]

    def halo_bytes_per_update(self) -> int:
        """bytes this rank sent to its neighbours during the last fused update"""
        return int(self.L.ssam_halo_bytes_last_update(self.h))

    def set_cell_layout(self, mode: int):
        """0 = caller's cell order, 1 = tile-compact when it pays (default), 2 = always; before mesh_set"""
        self._chk(self.L.ssam_set_cell_layout(self.h, mode))

    def layout_is_compact(self) -> bool:
        return bool(self.L.ssam_cell_layout_is_compact(self.h))

    def set_cell_kernel(self, variant: int):
        self._chk(self.L.ssam_set_cell_kernel(self.h, variant))

    def profile_enable(self, on: bool):
        self._chk(self.L.ssam_profile_enable(self.h, 1 if on else 0))

    def profile_get(self):
        ms = C.c_double()
        n = C.c_int64()
        self._chk(self.L.ssam_profile_get(self.h, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    # mesh -----------------------------------------------------------------------------------------
    def mesh_set(self, mesh, patch_u_bc=None):
        desc, keep = P.mesh_desc(mesh)
        self._chk(self.L.ssam_mesh_set(self.h, C.byref(desc)))
        self.mesh = mesh
        self.nCells, self.nBnd = mesh.nCells, mesh.nBoundaryFaces
        self.nTot = self.nCells + self.nBnd
        if patch_u_bc is not None:
            bc = np.ascontiguousarray(patch_u_bc, dtype=np.int32)
            self._chk(self.L.ssam_patch_u_bc_set(self.h, bc.ctypes.data_as(C.POINTER(C.c_int32))))

    def addressing(self, which: int) -> np.ndarray:
        n = C.c_int64()
        self._chk(self.L.ssam_addressing_size(self.h, which, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        self._chk(self.L.ssam_addressing_get(self.h, which, out.ctypes.data_as(C.POINTER(C.c_int32)), n.value))
        return out

    # fields ---------------------------------------------------------------------------------------
    def upload(self, field: int, internal=None, boundary=None):
        i, b = _f64(internal), _f64(boundary)
        nc = P.ncomp(field)
        if i is not None:
            assert i.size == self.nCells * nc, (i.shape, self.nCells, nc)
        if b is not None:
            assert b.size == self.nBnd * nc, (b.shape, self.nBnd, nc)
        self._chk(self.L.ssam_field_upload(self.h, field, _dptr(i), _dptr(b)))

    def download(self, field: int):
        nc = P.ncomp(field)
        shape_i = (self.nCells, nc) if nc > 1 else (self.nCells,)
        shape_b = (self.nBnd, nc) if nc > 1 else (self.nBnd,)
        i = np.empty(shape_i, dtype=np.float64)
        b = np.empty(shape_b, dtype=np.float64)
        self._chk(self.L.ssam_field_download(self.h, field, _dptr(i), _dptr(b)))
        return i, b

    def device_ptr(self, field: int, comp: int = 0) -> tuple[int, int]:
        p = C.c_void_p()
        n = C.c_int64()
        self._chk(self.L.ssam_field_device_ptr(self.h, field, comp, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def is_set(self, field: int) -> bool:
        return bool(self.L.ssam_field_is_set(self.h, field))

    # compute --------------------------------------------------------------------------------------
    def bc_rotational_slip(self, p: P.RotSlipParams):
        self._chk(self.L.ssam_bc_rotational_slip(self.h, C.byref(p)))

    def grad_u(self):
        self._chk(self.L.ssam_grad_u(self.h))

    def strain_rate(self, factor: float, out_field: int):
        self._chk(self.L.ssam_strain_rate(self.h, factor, out_field))

    def viscosity_correct(self, phase: int, p: P.ViscosityParams):
        self._chk(self.L.ssam_viscosity_correct(self.h, phase, C.byref(p)))

    def thermo_correct(self, v1, v2, mix):
        self._chk(self.L.ssam_thermo_correct(self.h, C.byref(v1), C.byref(v2), C.byref(mix)))

    def mixture(self, what: str, p: P.MixtureParams):
        self._chk(getattr(self.L, "ssam_mixture_" + what)(self.h, C.byref(p)))

    def mixture_k_phase(self, p: P.MixtureParams, phase: int):
        """k1() / k2() -> F_K1 / F_K2"""
        self._chk(self.L.ssam_mixture_k_phase(self.h, C.byref(p), phase))

    def solver_rho_rhocp(self, p: P.MixtureParams):
        self._chk(self.L.ssam_solver_rho_rhocp(self.h, C.byref(p)))

    def stress_strain(self, p: P.StressStrainParams):
        self._chk(self.L.ssam_stress_strain(self.h, C.byref(p)))

    def div_dev_rho_reff_explicit(self):
        """-fvc::div((rho*nuEff)*dev2(T(grad U))) -> F_DIVDEVRHOREFF"""
        self._chk(self.L.ssam_div_dev_rho_reff_explicit(self.h))

    def interface_correct(self, deltaN: float, flags: int = 0):
        """interfaceProperties::correct(): GRADALPHA, nHatf, CURVATURE from ALPHA1."""
        self._chk(self.L.ssam_interface_correct(self.h, deltaN, flags))

    def nhatf(self, n_internal_faces: int, n_boundary_faces: int) -> np.ndarray:
        out = np.zeros(n_internal_faces + n_boundary_faces)
        bi = out[n_internal_faces:]
        self._chk(self.L.ssam_interface_nhatf_download(self.h, _dptr(out), _dptr(bi) if n_boundary_faces else None))
        return out

    def face_field(self, which: int, n_internal_faces: int, n_boundary_faces: int) -> np.ndarray:
        out = np.zeros(n_internal_faces + n_boundary_faces)
        bi = out[n_internal_faces:]
        self._chk(self.L.ssam_face_field_download(self.h, which, _dptr(out), _dptr(bi) if n_boundary_faces else None))
        return out

    def face_field_upload(self, which: int, values):
        """surfaceScalarField [nFaces] (internal faces, then boundary faces) -> device slot"""
        v = np.ascontiguousarray(values, dtype=np.float64)
        nif = self.mesh.nInternalFaces
        assert v.shape == (self.mesh.nFaces,)
        self._chk(self.L.ssam_face_field_upload(self.h, which, _dptr(v), _dptr(v[nif:]) if self.nBnd else None))

    def alpha_compression_coeff(self, cAlpha: float, icAlpha: float, scAlpha: float):
        """multiRegionVoF/alphaEqn.H:58-87 -> FF_PHIC"""
        self._chk(self.L.ssam_alpha_compression_coeff(self.h, cAlpha, icAlpha, scAlpha))

    def face_interpolate(self, field: int, face_field: int):
        """fvc::interpolate(field) (linear) -> face field slot"""
        self._chk(self.L.ssam_face_interpolate(self.h, field, face_field))

    def mixture_face_visc(self, p: P.MixtureParams, nuf: bool = False):
        """muf() / nuf() -> FF_MUF / FF_NUF"""
        self._chk((self.L.ssam_mixture_nuf if nuf else self.L.ssam_mixture_muf)(self.h, C.byref(p)))

    def field_min_max(self, field: int):
        lo, hi = C.c_double(), C.c_double()
        self._chk(self.L.ssam_field_min_max(self.h, field, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def interface_delta_n(self) -> float:
        d = C.c_double()
        self._chk(self.L.ssam_interface_delta_n(self.h, C.byref(d)))
        return d.value

    def friction_correct(self, p: P.StickParams):
        self._chk(self.L.ssam_friction_correct(self.h, C.byref(p)))

    def friction_patch_centre(self):
        c = np.zeros(3)
        a = C.c_double()
        self._chk(self.L.ssam_friction_patch_centre(self.h, _dptr(c), C.byref(a)))
        return c, a.value

    def update_fused(self, p: P.UpdateParams, flags: int = 0):
        self._chk(self.L.ssam_update_fused(self.h, C.byref(p), flags))

    def prepare_fused(self, p: P.UpdateParams, flags: int = 0):
        self._chk(self.L.ssam_prepare_fused(self.h, C.byref(p), flags))

    def set_host_chunks(self, n: int):
        self._chk(self.L.ssam_set_host_chunks(self.h, n))

    def update_fused_host(self, p: P.UpdateParams, io: P.HostIO):
        self._chk(self.L.ssam_update_fused_host(self.h, C.byref(p), C.byref(io)))

    def fp64_peak(self) -> float:
        """measured FP64 issue ceiling, thread-level FP64 instructions per second (DFMA-only micro-kernel)"""
        v = C.c_double()
        self._chk(self.L.ssam_measure_fp64_peak(self.h, C.byref(v)))
        return v.value

    def debug_math(self, fn: int, x, y=None) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        yy = None if y is None else np.ascontiguousarray(np.broadcast_to(y, x.shape), dtype=np.float64)
        out = np.empty_like(x)
        self._chk(self.L.ssam_debug_math(self.h, fn, _dptr(x), _dptr(yy), _dptr(out), x.size))
        return out

    # comm -----------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        L = load()
        buf = C.create_string_buffer(P.UNIQUE_ID_BYTES)
        st = L.ssam_comm_unique_id(buf)
        if st != 0:
            raise SsamError(st, L.ssam_last_error(None).decode())
        return buf.raw

    def comm_init(self, uid: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(uid, P.UNIQUE_ID_BYTES)
        self._keep.append(buf)
        self._chk(self.L.ssam_comm_init(self.h, buf, rank, n_ranks))

    @staticmethod
    def comm_init_local(ctxs):
        """link contexts of this process into one rank group (context r = rank r)"""
        L = load()
        arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
        st = L.ssam_comm_init_local(arr, len(ctxs))
        if st != 0:
            raise SsamError(st, L.ssam_last_error(None).decode())

    def halo_exchange(self, fields):
        mask = 0
        for f in fields:
            mask |= 1 << f
        self._chk(self.L.ssam_halo_exchange(self.h, mask))

    def host_update(self, case, out_fields, U_b=None):
        """ssam_update_fused_host on numpy buffers: returns {field: (internal, boundary)}"""
        n, nb = case.mesh.nCells, case.mesh.nBoundaryFaces
        dp = C.POINTER(C.c_double)
        keep = []

        def ptr(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            keep.append(a)
            return a.ctypes.data_as(dp)

        io = P.HostIO()
        io.U_i, io.U_b = ptr(case.U_i), ptr(case.U_b if U_b is None else U_b)
        io.T_i, io.T_b = ptr(case.T_i), ptr(case.T_b)
        io.alpha1_i, io.alpha1_b = ptr(case.alpha1_i), ptr(case.alpha1_b)
        io.p_i, io.p_b = ptr(case.p_i), ptr(case.p_b)
        outs = list(out_fields)
        oi = [np.empty(n) for _ in outs]
        ob = [np.empty(max(nb, 1)) for _ in outs]
        io.nOut = len(outs)
        io.outFields = (C.c_int32 * len(outs))(*outs)
        io.out_i = (dp * len(outs))(*[a.ctypes.data_as(dp) for a in oi])
        io.out_b = (dp * len(outs))(*[a.ctypes.data_as(dp) for a in ob])
        self.update_fused_host(case.params, io)
        return {f: (oi[k], ob[k][:nb]) for k, f in enumerate(outs)}

    # convenience ----------------------------------------------------------------------------------
    def load_case(self, case, apply_bc=True):
        """mesh + inputs of a cases.Case; the tool-patch U comes from the rotational-slip BC kernel."""
        self.mesh_set(case.mesh, case.patch_u_bc)
        self.upload(P.F_U, case.U_i, case.U_b)
        self.upload(P.F_T, case.T_i, case.T_b)
        self.upload(P.F_ALPHA1, case.alpha1_i, case.alpha1_b)
        self.upload(P.F_P, case.p_i, case.p_b)
        if apply_bc and case.rotslip is not None:
            self.bc_rotational_slip(case.rotslip)
