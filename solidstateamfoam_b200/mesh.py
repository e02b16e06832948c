"""Host-side mesh container (OpenFOAM conventions) + synthetic generators.

The arrays are exactly what `ssam_mesh_set` (include/ssam_gpu.h) takes: geometry is an
*input* of the GPU path, never recomputed.  Generators live in
meshgen/meshgen.cpp (C++), this module only wraps them in numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _build

PATCH_GENERIC, PATCH_PROCESSOR, PATCH_EMPTY = 0, 1, 2

_lib = None


def _meshgen():
    global _lib
    if _lib is None:
        path = os.path.join(_build.LIB, "libssam_meshgen.so")
        if not os.path.exists(path):
            _build.build_meshgen()
        lib = C.CDLL(path)
        P = C.c_void_p
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
        lib.smg_hex_block.restype = P
        lib.smg_hex_block.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, ip, C.c_int]
        lib.smg_poly_refined.restype = P
        lib.smg_poly_refined.argtypes = [C.c_int, C.c_int, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double,
                                         C.c_double, dp]
        lib.smg_decompose.restype = C.c_int
        lib.smg_decompose.argtypes = [P, ip, C.c_int, C.POINTER(P)]
        lib.smg_free.argtypes = [P]
        lib.smg_sizes.argtypes = [P, ip]
        for n in ("owner", "neighbour", "patch_start", "patch_size", "patch_kind", "patch_nbr_rank", "cell_global",
                  "face_global", "face_flip"):
            getattr(lib, "smg_" + n).restype = ip
            getattr(lib, "smg_" + n).argtypes = [P]
        for n in ("Sf", "Cf", "magSf", "weights", "delta_coeffs", "V", "C"):
            getattr(lib, "smg_" + n).restype = dp
            getattr(lib, "smg_" + n).argtypes = [P]
        lib.smg_patch_name.restype = C.c_char_p
        lib.smg_patch_name.argtypes = [P, C.c_int]
        lib.smg_has_global_maps.restype = C.c_int32
        lib.smg_has_global_maps.argtypes = [P]
        lib.smg_points.restype = dp
        lib.smg_points.argtypes = [P]
        lib.smg_n_points.restype = C.c_int64
        lib.smg_n_points.argtypes = [P]
        lib.smg_write_polymesh.argtypes = [P, C.c_char_p, C.c_int]
        lib.smg_read_polymesh.restype = P
        lib.smg_read_polymesh.argtypes = [C.c_char_p]
        lib.smg_couple_processor_patches.argtypes = [C.POINTER(P), C.c_int]
        lib.smg_write_field.argtypes = [P, C.c_char_p, C.c_char_p, C.c_int, dp, dp, ip, C.c_int]
        lib.smg_read_field.argtypes = [P, C.c_char_p, C.c_int, dp, dp, ip]
        lib.smg_io_error.restype = C.c_char_p
        _lib = lib
    return _lib


@dataclass
class Mesh:
    nCells: int
    nInternalFaces: int
    nFaces: int
    owner: np.ndarray          # int32 [nFaces]
    neighbour: np.ndarray      # int32 [nInternalFaces]
    patchStart: np.ndarray     # int32 [nPatches]
    patchSize: np.ndarray
    patchKind: np.ndarray
    patchNbrRank: np.ndarray
    patchName: list
    Sf: np.ndarray             # float64 [nFaces,3]
    Cf: np.ndarray             # float64 [nFaces,3]
    magSf: np.ndarray          # float64 [nFaces]
    weights: np.ndarray        # float64 [nFaces] (internal: linear weights; boundary: patch weights)
    deltaCoeffs: np.ndarray    # float64 [nFaces]
    V: np.ndarray              # float64 [nCells]
    C: np.ndarray              # float64 [nCells,3]
    cellGlobal: np.ndarray | None = None
    faceGlobal: np.ndarray | None = None
    faceFlip: np.ndarray | None = None
    meta: dict = field(default_factory=dict)

    @property
    def nBoundaryFaces(self) -> int:
        return self.nFaces - self.nInternalFaces

    @property
    def nPatches(self) -> int:
        return len(self.patchStart)

    def patch_id(self, name: str) -> int:
        """polyBoundaryMesh::findPatchID: -1 when absent (constantSlipStick.C:82)."""
        try:
            return self.patchName.index(name)
        except ValueError:
            return -1

    def patch_slice_b(self, p: int) -> slice:
        """slice of patch p in boundary-face numbering (0 = first boundary face)."""
        s = int(self.patchStart[p]) - self.nInternalFaces
        return slice(s, s + int(self.patchSize[p]))

    def face_cells(self, p: int) -> np.ndarray:
        s = int(self.patchStart[p])
        return self.owner[s:s + int(self.patchSize[p])]


def _arr(ptr, n, dtype, shape=None):
    if n == 0:
        a = np.zeros(0, dtype=dtype)
    else:
        a = np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)
    return a.reshape(shape) if shape is not None else a


def _from_handle(h, free=True) -> Mesh:
    lib = _meshgen()
    sz = (C.c_int32 * 4)()
    lib.smg_sizes(h, sz)
    nc, nif, nf, npch = (int(v) for v in sz)
    m = Mesh(
        nCells=nc, nInternalFaces=nif, nFaces=nf,
        owner=_arr(lib.smg_owner(h), nf, np.int32),
        neighbour=_arr(lib.smg_neighbour(h), nif, np.int32),
        patchStart=_arr(lib.smg_patch_start(h), npch, np.int32),
        patchSize=_arr(lib.smg_patch_size(h), npch, np.int32),
        patchKind=_arr(lib.smg_patch_kind(h), npch, np.int32),
        patchNbrRank=_arr(lib.smg_patch_nbr_rank(h), npch, np.int32),
        patchName=[lib.smg_patch_name(h, p).decode() for p in range(npch)],
        Sf=_arr(lib.smg_Sf(h), 3 * nf, np.float64, (nf, 3)),
        Cf=_arr(lib.smg_Cf(h), 3 * nf, np.float64, (nf, 3)),
        magSf=_arr(lib.smg_magSf(h), nf, np.float64),
        weights=_arr(lib.smg_weights(h), nf, np.float64),
        deltaCoeffs=_arr(lib.smg_delta_coeffs(h), nf, np.float64),
        V=_arr(lib.smg_V(h), nc, np.float64),
        C=_arr(lib.smg_C(h), 3 * nc, np.float64, (nc, 3)),
    )
    if lib.smg_has_global_maps(h):
        m.cellGlobal = _arr(lib.smg_cell_global(h), nc, np.int32)
        m.faceGlobal = _arr(lib.smg_face_global(h), nf, np.int32)
        m.faceFlip = _arr(lib.smg_face_flip(h), nf, np.int32)
    if free:
        lib.smg_free(h)
    return m


def _d(a, n):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    assert a.size == n
    return a


IDENTITY = (1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0)


def hex_block(nx, ny, nz, origin=(0.0, 0.0, 0.0), length=(1.0, 1.0, 1.0), grading=(1.0, 1.0, 1.0),
              affine=IDENTITY, side_nbr_rank=(-1,) * 6, my_rank=0, _keep=False):
    """blockMesh-ordered hex block; patches xmin,xmax,ymin,ymax,zmin,zmax (+ processor sides)."""
    lib = _meshgen()
    o, l, g, a = _d(origin, 3), _d(length, 3), _d(grading, 3), _d(affine, 9)
    s = np.ascontiguousarray(np.asarray(side_nbr_rank, dtype=np.int32))
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    h = lib.smg_hex_block(nx, ny, nz, o.ctypes.data_as(dp), l.ctypes.data_as(dp), g.ctypes.data_as(dp),
                          a.ctypes.data_as(dp), s.ctypes.data_as(ip), my_rank)
    if _keep:
        return h
    m = _from_handle(h)
    m.meta.update(kind="hex", dims=(nx, ny, nz), origin=tuple(o), length=tuple(l))
    return m


def poly_refined(nbx, nby, nbz, origin=(0.0, 0.0, 0.0), length=(1.0, 1.0, 1.0), r1=0.3, d1=0.4, r2=0.15, d2=0.2,
                 affine=IDENTITY, _keep=False):
    """Base hex grid with two 2:1 octree refinement levels in cylinders under the zmax patch."""
    lib = _meshgen()
    o, l, a = _d(origin, 3), _d(length, 3), _d(affine, 9)
    dp = C.POINTER(C.c_double)
    h = lib.smg_poly_refined(nbx, nby, nbz, o.ctypes.data_as(dp), l.ctypes.data_as(dp), r1, d1, r2, d2,
                             a.ctypes.data_as(dp))
    if _keep:
        return h
    m = _from_handle(h)
    m.meta.update(kind="poly", dims=(nbx, nby, nbz), origin=tuple(o), length=tuple(l))
    return m


def decompose(handle_or_builder, cell_proc, n_procs):
    """decomposePar stand-in: `handle_or_builder` is a callable returning a kept handle."""
    lib = _meshgen()
    h = handle_or_builder() if callable(handle_or_builder) else handle_or_builder
    cp = np.ascontiguousarray(np.asarray(cell_proc, dtype=np.int32))
    out = (C.c_void_p * n_procs)()
    lib.smg_decompose(h, cp.ctypes.data_as(C.POINTER(C.c_int32)), n_procs, out)
    parts = [_from_handle(out[p]) for p in range(n_procs)]
    g = _from_handle(h)
    return g, parts


def slab_partition(nx, ny, nz_per_rank, rank, n_ranks, origin=(0.0, 0.0, 0.0), cell=(1.0, 1.0, 1.0)):
    """Rank `rank`'s z-slab of an nx x ny x (nz_per_rank*n_ranks) hex box, generated directly
    with processor patches on the shared z faces (weak-scaling bench; no global mesh is built)."""
    side = [-1] * 6
    if rank > 0:
        side[4] = rank - 1
    if rank < n_ranks - 1:
        side[5] = rank + 1
    o = (origin[0], origin[1], origin[2] + rank * nz_per_rank * cell[2])
    m = hex_block(nx, ny, nz_per_rank, o, (nx * cell[0], ny * cell[1], nz_per_rank * cell[2]),
                  side_nbr_rank=side, my_rank=rank)
    m.meta.update(rank=rank, n_ranks=n_ranks)
    return m


# ---- OpenFOAM case I/O ------------------------------------------------------------
class FoamIOError(RuntimeError):
    pass


class MeshHandle:
    """A mesh kept alive on the C++ side (needed for polyMesh / field I/O, which uses points and faces)."""

    def __init__(self, h):
        if not h:
            raise FoamIOError(_meshgen().smg_io_error().decode())
        self.h = C.c_void_p(h) if not isinstance(h, C.c_void_p) else h

    def mesh(self) -> Mesh:
        return _from_handle(self.h, free=False)

    def close(self):
        if self.h:
            _meshgen().smg_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def write_polymesh(self, polymesh_dir: str, binary: bool = False):
        if _meshgen().smg_write_polymesh(self.h, polymesh_dir.encode(), 1 if binary else 0):
            raise FoamIOError(_meshgen().smg_io_error().decode())

    def write_field(self, path: str, name: str, internal, boundary, patch_bc=None, binary: bool = False):
        i = np.ascontiguousarray(internal, dtype=np.float64)
        b = np.ascontiguousarray(boundary, dtype=np.float64)
        nc = 1 if i.ndim == 1 else i.shape[1]
        bc = None
        if patch_bc is not None:
            bca = np.ascontiguousarray(patch_bc, dtype=np.int32)
            bc = bca.ctypes.data_as(C.POINTER(C.c_int32))
        dp = C.POINTER(C.c_double)
        if _meshgen().smg_write_field(self.h, path.encode(), name.encode(), nc, i.ctypes.data_as(dp),
                                      b.ctypes.data_as(dp), bc, 1 if binary else 0):
            raise FoamIOError(_meshgen().smg_io_error().decode())

    def read_field(self, path: str, ncomp: int):
        """(internal, boundary, patch_bc) of a volScalarField / volVectorField file; zeroGradient patches are
        evaluated (value = adjacent cell), patch_bc: 0 value-type, 1 zeroGradient."""
        lib = _meshgen()
        sz = (C.c_int32 * 4)()
        lib.smg_sizes(self.h, sz)
        nc, nif, nf, npch = (int(v) for v in sz)
        shape_i = (nc, ncomp) if ncomp > 1 else (nc,)
        shape_b = (nf - nif, ncomp) if ncomp > 1 else (nf - nif,)
        i = np.zeros(shape_i)
        b = np.zeros(shape_b)
        bc = np.zeros(npch, dtype=np.int32)
        dp = C.POINTER(C.c_double)
        if lib.smg_read_field(self.h, path.encode(), ncomp, i.ctypes.data_as(dp), b.ctypes.data_as(dp),
                              bc.ctypes.data_as(C.POINTER(C.c_int32))):
            raise FoamIOError(lib.smg_io_error().decode())
        return i, b, bc


def hex_block_handle(*args, **kw) -> MeshHandle:
    return MeshHandle(hex_block(*args, _keep=True, **kw))


def read_polymesh(polymesh_dir: str) -> MeshHandle:
    """constant/polyMesh of a serial case or of one processorN directory (ascii or binary)."""
    return MeshHandle(_meshgen().smg_read_polymesh(polymesh_dir.encode()))


def read_decomposed_case(case_dir: str, n_procs: int):
    """processor0..N-1/constant/polyMesh, with the processor-patch weights / deltaCoeffs completed from the
    neighbour side (what processorFvPatch::makeWeights gets through its geometry exchange)."""
    hs = [read_polymesh(os.path.join(case_dir, f"processor{r}", "constant", "polyMesh")) for r in range(n_procs)]
    arr = (C.c_void_p * n_procs)(*[h.h for h in hs])
    if _meshgen().smg_couple_processor_patches(arr, n_procs):
        raise FoamIOError(_meshgen().smg_io_error().decode())
    return hs
