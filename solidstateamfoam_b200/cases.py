"""Synthetic cases for BASELINE.json's five configs (SURVEY.md §8d "Concrete synthetic inputs").

The reference ships no case files, so all material constants below are *our* choices (literature /
placeholder magnitudes, flagged in SURVEY.md §8d); meshes come from mesh.py, fields are seeded
(Philox, seed = 20261019 + cfg index).  Every config takes a `scale` so the same case runs at a size
the CPU oracle finishes in seconds (parity tests) and at BASELINE.json's full size (bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import mesh as M
from . import params as P

SEED0 = 20261019
OMEGA = (0.0, 0.0, 41.9)
VT = (2e-3, 0.0, 0.0)


@dataclass
class Case:
    name: str
    mesh: M.Mesh
    U_i: np.ndarray
    U_b: np.ndarray
    T_i: np.ndarray
    T_b: np.ndarray
    alpha1_i: np.ndarray
    alpha1_b: np.ndarray
    p_i: np.ndarray
    p_b: np.ndarray
    params: P.UpdateParams
    rotslip: P.RotSlipParams | None
    patch_u_bc: np.ndarray
    outputs: list = field(default_factory=list)   # SSAM_F_* the fused update produces for this case
    bytes_per_update: int = 0                     # algorithmic bytes, SURVEY §8d formula
    meta: dict = field(default_factory=dict)


def algorithmic_bytes(mesh: M.Mesh, n_fric: int, norton_hoff: bool) -> int:
    """B = 40*N_if + 52*N_bf + 128*N_c + 40*N_fric  (112/cell for NortonHoff) -- SURVEY.md §8d."""
    per_cell = 112 if norton_hoff else 128
    return 40 * mesh.nInternalFaces + 52 * mesh.nBoundaryFaces + per_cell * mesh.nCells + 40 * n_fric


def _fields(mesh: M.Mesh, seed: int, alpha_mode: str, z0: float | None = None, step: int = 0, dt: float = 1e-3,
            box=None):
    """U, T, alpha1, p at cell centres and boundary-face centres (boundary = 'calculated' values of the
    same analytic fields; tool-patch U is overwritten by the rotational-slip BC afterwards)."""
    rng = np.random.Generator(np.random.Philox(seed))
    nc, nif = mesh.nCells, mesh.nInternalFaces
    X = np.concatenate([mesh.C, mesh.Cf[nif:]], axis=0)
    if box is None:
        lo = X.min(axis=0)
        hi = X.max(axis=0)
    else:  # global box of a decomposed case, so the analytic fields are continuous across ranks
        lo, hi = np.asarray(box[0], dtype=float), np.asarray(box[1], dtype=float)
    L = hi - lo
    xc = 0.5 * (lo + hi)
    ztop = hi[2]
    om = np.array(OMEGA)
    # rigid rotation about the vertical axis through the top-patch centre, decaying with depth,
    # rotated in phase by omega*dt*step, + 1e-3 noise  (near-rigid rotation = the AFSD stick zone,
    # the worst case for summation-order sensitivity)
    rel = X - np.array([xc[0], xc[1], 0.0])
    ph = om[2] * dt * step
    c, s = math.cos(ph), math.sin(ph)
    rel = np.stack([c * rel[:, 0] - s * rel[:, 1], s * rel[:, 0] + c * rel[:, 1], rel[:, 2]], axis=1)
    decay = np.exp(-(ztop - X[:, 2]) / (0.125 * L[2]))
    U = np.cross(np.broadcast_to(om, rel.shape), rel) * decay[:, None]
    U += 1e-3 * rng.uniform(-1.0, 1.0, size=U.shape) * (np.abs(om[2]) * 0.5 * L[0])
    r2 = (X[:, 0] - xc[0]) ** 2 + (X[:, 1] - xc[1]) ** 2 + (X[:, 2] - ztop) ** 2
    T = 300.0 + 500.0 * np.exp(-r2 / (0.3 * L[0]) ** 2) + 100.0 * step * dt
    if z0 is None:
        z0 = lo[2] + 0.6 * L[2]
    # substrate below z0 plus a column of deposited material under the tool, so the friction patch
    # (zmax) touches material (alpha1 = 1) inside the tool radius and air outside it
    rxy = np.sqrt((X[:, 0] - xc[0]) ** 2 + (X[:, 1] - xc[1]) ** 2)
    rtool = 0.3 * L[0]
    if alpha_mode == "step":
        alpha = np.where((X[:, 2] < z0) | (rxy < rtool), 1.0, 0.0)
    else:  # smooth tanh interfaces overshooting [0,1] by 1e-3 to exercise the clamp
        az = 0.5 * (1.0 - np.tanh((X[:, 2] - z0) / (0.02 * L[2])))
        ar = 0.5 * (1.0 - np.tanh((rxy - rtool) / (0.02 * L[0])))
        alpha = np.maximum(az, ar) * (1.0 + 2e-3) - 1e-3
    p = 1e5 * (1.0 + 0.1 * rng.uniform(-1.0, 1.0, size=X.shape[0]))
    return (np.ascontiguousarray(U[:nc]), np.ascontiguousarray(U[nc:]), T[:nc].copy(), T[nc:].copy(),
            alpha[:nc].copy(), alpha[nc:].copy(), p[:nc].copy(), p[nc:].copy())


AIR = dict(nu=1.5e-5, rho=1.2, cp=1005.0, k=0.026)


def sheppard_wright_aa6061():
    # AA6061 literature constants (not in the reference)
    return P.sheppard_wright(sigmaR=2.222e7, Q=1.45e5, R=8.314, A=2.409e8, n=3.55, rho=2700.0, nuMin=1e-6,
                             nuMax=1e4, eDotMin=1e-6)


def fks_placeholder():
    return P.fks(a1=3.24e8, a2=1.14e8, b1=0.01, b2=1.0, b3=1.0, e0=1.0, c1=6.0, c2=1.0, Tm=855.0, Ta=293.0,
                 rho=2700.0, nuMin=1e-6, nuMax=1e4, eDotMin=1e-6)


def fks_general_exponents():
    """FKS with b2, c2 != 1: no pow(x, 1) shortcut applies (two extra pow per cell); same values as the golden case
    cfg2_fks_general_exponents (tests/golden/gen_reference_golden.py)"""
    return P.fks(a1=2.1e8, a2=0.6e8, b1=0.05, b2=0.7, b3=1.3, e0=0.01, c1=3.0, c2=2.5, Tm=900.0, Ta=300.0, rho=2650.0,
                 nuMin=1e-8, nuMax=1e6, eDotMin=1e-9)


def norton_hoff_placeholder():
    return P.norton_hoff(mu0=1e8, m=0.2, rho=2700.0, nuMin=1e-6, nuMax=1e4, strainRateEffMin=1e-6)


def _finish(name, mesh, flds, visc1, mix, stick, rotslip, seed, meta=None) -> Case:
    top = mesh.patch_id("zmax")
    bc = np.full(mesh.nPatches, P.BC_VALUE, dtype=np.int32)
    # side walls zeroGradient, bottom + top fixedValue: exercises both snGrad kinds
    for nm in ("xmin", "xmax", "ymin", "ymax"):
        pid = mesh.patch_id(nm)
        if pid >= 0:
            bc[pid] = P.BC_ZERO_GRADIENT
    up = P.UpdateParams(visc1=visc1, visc2=P.newtonian(AIR["nu"], AIR["rho"]), mix=mix, stick=stick)
    nh = visc1.model == P.VISC_NORTON_HOFF
    outs = [P.F_EPSDOT, P.F_NU1, P.F_NU2, P.F_NU, P.F_MUEFF, P.F_RHO, P.F_CP, P.F_K, P.F_QVISC]
    if not nh:
        outs += [P.F_SIGMAF, P.F_SIGMAY]
    n_fric = 0
    if stick.fricPatch >= 0:
        outs.append(P.F_QFRIC)
        n_fric = int(mesh.patchSize[stick.fricPatch])
    U_i, U_b, T_i, T_b, a_i, a_b, p_i, p_b = flds
    # zeroGradient patches carry U_b = U_P
    for pid in range(mesh.nPatches):
        if bc[pid] == P.BC_ZERO_GRADIENT and mesh.patchKind[pid] == P.PATCH_GENERIC:
            U_b[mesh.patch_slice_b(pid)] = U_i[mesh.face_cells(pid)]
    c = Case(name=name, mesh=mesh, U_i=U_i, U_b=U_b, T_i=T_i, T_b=T_b, alpha1_i=a_i, alpha1_b=a_b, p_i=p_i, p_b=p_b,
             params=up, rotslip=rotslip, patch_u_bc=bc, outputs=outs,
             bytes_per_update=algorithmic_bytes(mesh, n_fric, nh), meta=dict(meta or {}, seed=seed, top=top))
    return c


def _hex(dims, cell, side=(-1,) * 6, rank=0, origin=(0.0, 0.0, 0.0), affine=M.IDENTITY, grading=(1.0, 1.0, 1.0)):
    nx, ny, nz = dims
    return M.hex_block(nx, ny, nz, origin, (nx * cell, ny * cell, nz * cell), grading, affine, side, rank)


def cfg1(dims=(64, 64, 32), step=0, affine=M.IDENTITY, grading=(1.0, 1.0, 1.0), visc1=None) -> Case:
    """hex 64x64x32, alpha1 step, SheppardWright AA6061, constantRotationalSlip + constantSlipStick."""
    seed = SEED0 + 1
    m = _hex(dims, 0.5e-3, affine=affine, grading=grading)
    top = m.patch_id("zmax")
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], P.k_constant(167.0), P.k_constant(AIR["k"]))
    stick = P.constant_slip_stick(top, betaTQ=0.9, delta=0.5, omega=OMEGA, muf=0.3)
    rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
    return _finish("cfg1", m, _fields(m, seed, "step", step=step), visc1 or sheppard_wright_aa6061(), mix, stick, rot,
                   seed)


def cfg2(dims=(256, 256, 128), units="Kelvin", side=(-1,) * 6, rank=0, origin=(0.0, 0.0, 0.0), z0=None,
         box=None, visc1=None, k_coeffs=(100.0, 0.2, -1e-4, 1e-8)) -> Case:
    """hex 256x256x128, FKS + cubic k(T) for phase 1 (Kelvin or Celsius fit)."""
    seed = SEED0 + 2 + rank
    m = _hex(dims, 0.25e-3, side=side, rank=rank, origin=origin)
    top = m.patch_id("zmax")
    k1 = P.k_cubic(*k_coeffs, 293.0 if units == "Kelvin" else 20.0,
                   855.0 if units == "Kelvin" else 582.0, units)
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], k1, P.k_constant(AIR["k"]))
    if top >= 0:
        stick = P.constant_slip_stick(top, betaTQ=0.9, delta=0.5, omega=OMEGA, muf=0.3)
        rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
    else:  # interior slab of a decomposed box: no tool patch on this rank
        stick = P.no_friction_patch(0.9)
        rot = None
    return _finish("cfg2", m, _fields(m, seed, "tanh", z0=z0, box=box), visc1 or fks_placeholder(), mix, stick, rot,
                   seed, meta=dict(rank=rank))


def cfg3(base=(128, 128, 128)) -> Case:
    """~4M-cell polyhedral substrate+air mesh (2 levels of 2:1 refinement under the tool),
    exponentialRotationalSlip BC + exponentialSlipStick friction heating, SheppardWright."""
    seed = SEED0 + 3
    nb = base
    h = 0.4e-3 * 128 / nb[0]
    L = (nb[0] * h, nb[1] * h, nb[2] * h)
    m = M.poly_refined(nb[0], nb[1], nb[2], (0.0, 0.0, 0.0), L, r1=0.33 * L[0], d1=0.30 * L[2], r2=0.2 * L[0],
                       d2=0.14 * L[2])
    top = m.patch_id("zmax")
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], P.k_constant(167.0), P.k_constant(AIR["k"]))
    stick = P.exponential_slip_stick(top, betaTQ=0.9, omega=OMEGA, omega0=41.9, delta0=0.3, Rs=19e-3, mu0=0.4,
                                     lambda_=0.5)
    rot = P.exponential_rotational_slip(top, OMEGA, 41.9, 0.3, 19e-3, VT)
    return _finish("cfg3", m, _fields(m, seed, "tanh"), sheppard_wright_aa6061(), mix, stick, rot, seed)


def cfg4(dims=(256, 256, 256), visc1=None) -> Case:
    """fluid region of the two-region case: hex 256^3, NortonHoff, qfric skipped (SURVEY D-1)."""
    seed = SEED0 + 4
    m = _hex(dims, 0.25e-3)
    top = m.patch_id("zmax")
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], P.k_constant(167.0), P.k_constant(AIR["k"]))
    rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
    return _finish("cfg4", m, _fields(m, seed, "tanh"), visc1 or norton_hoff_placeholder(), mix,
                   P.no_friction_patch(0.9), rot, seed)


def cfg5_slab(rank, n_ranks, dims_per_rank=(512, 512, 32)) -> Case:
    """rank's z-slab of the AFSD box (512x512x256 at 8 ranks), SheppardWright, processor patches."""
    seed = SEED0 + 5
    nx, ny, nzr = dims_per_rank
    cell = 0.25e-3
    side = [-1] * 6
    if rank > 0:
        side[4] = rank - 1
    if rank < n_ranks - 1:
        side[5] = rank + 1
    m = _hex((nx, ny, nzr), cell, side=tuple(side), rank=rank, origin=(0.0, 0.0, rank * nzr * cell))
    top = m.patch_id("zmax")
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], P.k_constant(167.0), P.k_constant(AIR["k"]))
    if top >= 0:
        stick = P.constant_slip_stick(top, betaTQ=0.9, delta=0.5, omega=OMEGA, muf=0.3)
        rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
    else:
        stick = P.no_friction_patch(0.9)
        rot = None
    z0 = 0.6 * n_ranks * nzr * cell
    box = ((0.0, 0.0, 0.0), (nx * cell, ny * cell, n_ranks * nzr * cell))
    return _finish("cfg5", m, _fields(m, seed + rank, "tanh", z0=z0, box=box), sheppard_wright_aa6061(), mix, stick,
                   rot, seed,
                   meta=dict(rank=rank, n_ranks=n_ranks))


def block_grid(n_ranks: int, dims) -> tuple:
    """(px, py, pz) with px*py*pz = n_ranks and equal blocks, minimising the processor faces of the busiest rank --
    the structured partition a graph partitioner (scotch) finds for a box; ties go to the split with the fewest
    cuts normal to x, then y.  512x512x256 -> 1x2x1, 2x2x1, 2x2x2 (131 072 processor faces per rank at 8 ranks,
    SURVEY 8e)."""
    best = None
    for px in range(1, n_ranks + 1):
        if n_ranks % px or dims[0] % px:
            continue
        for py in range(1, n_ranks // px + 1):
            if (n_ranks // px) % py or dims[1] % py:
                continue
            pz = n_ranks // (px * py)
            if dims[2] % pz:
                continue
            p = (px, py, pz)
            loc = [dims[a] // p[a] for a in range(3)]
            cut = sum(min(2, p[a] - 1) * loc[(a + 1) % 3] * loc[(a + 2) % 3] for a in range(3))
            key = (cut, px, py)
            if best is None or key < best[0]:
                best = (key, p)
    if best is None:
        raise ValueError(f"cannot split {dims} into {n_ranks} equal blocks")
    return best[1]


def _block_sides(idx, grid) -> tuple:
    """neighbour rank across each of the six sides (xmin, xmax, ymin, ymax, zmin, zmax) of block idx of a structured
    grid of blocks numbered x-fastest; -1 = physical boundary"""
    side = [-1] * 6
    for a in range(3):
        for s, d in ((0, -1), (1, +1)):
            j = list(idx)
            j[a] += d
            if 0 <= j[a] < grid[a]:
                side[2 * a + s] = j[0] + grid[0] * (j[1] + grid[1] * j[2])
    return tuple(side)


def cfg5_part(rank: int, n_ranks: int, decomp: str = "block", dims=(512, 512, 256)) -> Case:
    """rank's partition of the cfg5 AFSD box (BASELINE.json configs[4]: 64M cells, SheppardWright), generated
    directly with its processor patches (no global mesh is ever built).  decomp = "block": structured blocks with up
    to three neighbours per rank (block_grid); "slab": z-slabs (two neighbours)."""
    seed = SEED0 + 5
    cell = 0.25e-3
    grid = block_grid(n_ranks, dims) if decomp == "block" else (1, 1, n_ranks)
    if any(dims[a] % grid[a] for a in range(3)):
        raise ValueError(f"cannot split {dims} {decomp}-wise into {n_ranks}")
    idx = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
    loc = tuple(dims[a] // grid[a] for a in range(3))

    origin = tuple(idx[a] * loc[a] * cell for a in range(3))
    m = _hex(loc, cell, side=_block_sides(idx, grid), rank=rank, origin=origin)
    top = m.patch_id("zmax")
    mix = P.mixture(2700.0, AIR["rho"], 896.0, AIR["cp"], P.k_constant(167.0), P.k_constant(AIR["k"]))
    # Every rank names the friction patch, whether or not it holds any of its faces: decomposePar keeps each physical
    # patch on every rank (possibly empty) and every rank takes part in constantSlipStick's reduce() of the patch
    # centroid (constantSlipStick.C:133-135).  A rank that passed "no friction patch" instead would stay out of that
    # all-reduce and stall the others.
    if top >= 0:
        stick = P.constant_slip_stick(top, betaTQ=0.9, delta=0.5, omega=OMEGA, muf=0.3)
        rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
    else:
        stick = P.no_friction_patch(0.9)
        rot = None
    box = ((0.0, 0.0, 0.0), tuple(dims[a] * cell for a in range(3)))
    return _finish("cfg5", m, _fields(m, seed + rank, "tanh", z0=0.6 * dims[2] * cell, box=box),
                   sheppard_wright_aa6061(), mix, stick, rot, seed,
                   meta=dict(rank=rank, n_ranks=n_ranks, decomp=decomp, grid=grid))


def parity_parts(n_ranks: int, decomp: str):
    """A small sheared box in n_ranks partitions for the multi-rank parity checks (tests/multigpu_parity.py and the
    parity record of bench.py): "slab" = z-slabs generated directly, "block" = structured blocks generated directly
    (12^3 box, block_grid), "random" = a ragged blocky cellProc list (random owner per 3x3x2 brick, every pair of
    ranks touches) through the decomposePar stand-in."""
    dims = (12, 10, 4 * max(n_ranks, 2))
    cell = 0.5e-3
    box = ((0.0, 0.0, 0.0), (dims[0] * cell, dims[1] * cell, dims[2] * cell))
    if n_ranks == 1:
        meshes = [M.hex_block(*dims, length=box[1], affine=(1.0, 0.2, 0.1, 0.0, 1.0, 0.15, 0.0, 0.0, 1.0))]
    elif decomp == "slab":
        nz = dims[2] // n_ranks
        meshes = [M.hex_block(dims[0], dims[1], nz, (0, 0, r * nz * cell), (dims[0] * cell, dims[1] * cell, nz * cell),
                              side_nbr_rank=tuple([-1] * 4 + [r - 1 if r > 0 else -1, r + 1 if r < n_ranks - 1 else -1]),
                              my_rank=r) for r in range(n_ranks)]
    elif decomp == "block":
        # the structured blocks bench.py times at full size (block_grid / cfg5_part), up to three neighbours per rank
        dims = (12, 12, 12)
        box = ((0.0, 0.0, 0.0), (dims[0] * cell, dims[1] * cell, dims[2] * cell))
        grid = block_grid(n_ranks, dims)
        loc = tuple(dims[a] // grid[a] for a in range(3))
        meshes = []
        for r in range(n_ranks):
            idx = (r % grid[0], (r // grid[0]) % grid[1], r // (grid[0] * grid[1]))
            meshes.append(_hex(loc, cell, side=_block_sides(idx, grid), rank=r,
                               origin=tuple(idx[a] * loc[a] * cell for a in range(3))))
    else:
        h = M.hex_block(*dims, length=box[1], affine=(1.0, 0.2, 0.1, 0.0, 1.0, 0.15, 0.0, 0.0, 1.0), _keep=True)
        rng = np.random.default_rng(11)
        ii, jj, kk = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
        brick = (ii // 3) + 7 * (jj // 3) + 53 * (kk // 2)
        owner_of_brick = rng.integers(0, n_ranks, size=brick.max() + 1)
        owner_of_brick[:n_ranks] = np.arange(n_ranks)  # no empty rank
        cp = owner_of_brick[brick].transpose(2, 1, 0).reshape(-1)
        _, meshes = M.decompose(h, cp, n_ranks)
    parts = []
    for r, m in enumerate(meshes):
        top = m.patch_id("zmax")
        mix = P.mixture(2700.0, 1.2, 896.0, 1005.0, P.k_cubic(100.0, 0.2, -1e-4, 1e-8, 293.0, 855.0), P.k_constant(0.026))
        stick = P.constant_slip_stick(top, 0.9, 0.5, OMEGA, 0.3)
        rot = P.constant_rotational_slip(top, OMEGA, 0.5, VT)
        parts.append(_finish("mg", m, _fields(m, 777 + r, "tanh", box=box), sheppard_wright_aa6061(), mix, stick, rot, 777))
    return parts


def advance(case: Case, step: int) -> None:
    """cfg1's '10 updates with U,T advanced analytically per step' (SURVEY §8d)."""
    flds = _fields(case.mesh, case.meta["seed"], "step" if case.name == "cfg1" else "tanh", step=step)
    case.U_i, case.U_b, case.T_i, case.T_b = flds[0], flds[1], flds[2], flds[3]
    for pid in range(case.mesh.nPatches):
        if case.patch_u_bc[pid] == P.BC_ZERO_GRADIENT and case.mesh.patchKind[pid] == P.PATCH_GENERIC:
            case.U_b[case.mesh.patch_slice_b(pid)] = case.U_i[case.mesh.face_cells(pid)]


# ---- OpenFOAM-style case directory for the C++ host mirror (host/case_driver.cpp) ------------------------
def _visc_dict(v: P.ViscosityParams, rho, cp, kc: P.Conductivity) -> str:
    def coeffs(name, keys, extra=()):
        body = "\n".join(f"        {k} {getattr(v, k)!r};" for k in keys)
        body += "".join(f"\n        {k} {val!r};" for k, val in extra)
        return f"    transportModel {name};\n    {name}Coeffs\n    {{\n{body}\n    }}\n"

    lim = [(k, getattr(v, k)) for k in ("nuMin", "nuMax", "eDotMin")]
    if v.model == P.VISC_SHEPPARD_WRIGHT:
        s = coeffs("SheppardWright", ["sigmaR", "Q", "R", "A", "n", "rho"], lim)
    elif v.model == P.VISC_FKS:
        s = coeffs("FKS", ["a1", "a2", "b1", "b2", "b3", "e0", "c1", "c2", "Tm", "Ta", "rho"], lim)
    elif v.model == P.VISC_NORTON_HOFF:
        s = coeffs("NortonHoff", ["mu0", "m", "rho", "nuMin", "nuMax", "strainRateEffMin"])
    else:
        s = f"    transportModel Newtonian;\n    nu {v.nu!r};\n"
    s += f"    rho {rho!r};\n    cp {cp!r};\n"
    if kc.kModel == P.K_CUBIC:
        units = "Kelvin" if kc.units == P.UNITS_KELVIN else "Celsius"
        s += (f"    kModel cubic;\n    k0 {kc.k0!r};\n    k1 {kc.k1!r};\n    k2 {kc.k2!r};\n    k3 {kc.k3!r};\n"
              f"    Tmin {kc.Tmin!r};\n    Tmax {kc.Tmax!r};\n    units {units};\n")
    else:
        s += f"    kModel constant;\n    k {kc.k!r};\n"
    return s


def write_case_dir(path: str, case: Case, eps0: float = 1.0, mu0: float = 1.0) -> None:
    """mesh.bin / fields.bin / patches.txt + constant/transportProperties, constant/frictionProperties, 0/U."""
    import os

    m = case.mesh
    os.makedirs(os.path.join(path, "constant"), exist_ok=True)
    os.makedirs(os.path.join(path, "0"), exist_ok=True)
    nif = m.nInternalFaces
    with open(os.path.join(path, "mesh.bin"), "wb") as fh:
        np.array([m.nCells, nif, m.nFaces, m.nPatches], dtype=np.int32).tofile(fh)
        for a in (m.owner, m.neighbour, m.patchStart, m.patchSize, m.patchKind, m.patchNbrRank, case.patch_u_bc):
            np.ascontiguousarray(a, dtype=np.int32).tofile(fh)
        for a in (m.Sf, m.weights, m.V, m.C, m.Cf[nif:], m.magSf[nif:], m.deltaCoeffs[nif:]):
            np.ascontiguousarray(a, dtype=np.float64).tofile(fh)
    with open(os.path.join(path, "patches.txt"), "w") as fh:
        fh.write("\n".join(m.patchName) + "\n")
    nb = m.nBoundaryFaces
    with open(os.path.join(path, "fields.bin"), "wb") as fh:
        for i, b in ((case.U_i, case.U_b), (case.T_i, case.T_b), (case.alpha1_i, case.alpha1_b), (case.p_i, case.p_b),
                     (np.full(m.nCells, eps0), np.full(nb, eps0)), (np.full(m.nCells, mu0), np.full(nb, mu0))):
            np.ascontiguousarray(i, dtype=np.float64).tofile(fh)
            np.ascontiguousarray(b, dtype=np.float64).tofile(fh)
    _write_dicts(path, case)


def write_foam_case(path: str, case: Case, handle, binary: bool = False) -> None:
    """The same case as a real OpenFOAM case directory: constant/polyMesh, 0/{U,temp,alpha.metal,p} and the
    dictionaries (`handle` = mesh.MeshHandle of the case's hex block, which carries points and faces)."""
    import os

    handle.write_polymesh(os.path.join(path, "constant", "polyMesh"), binary=binary)
    _write_dicts(path, case)
    u_path = os.path.join(path, "0", "U")
    bc_text = open(u_path).read()  # the rotational-slip patch entry written by _write_dicts
    handle.write_field(u_path, "U", case.U_i, case.U_b, case.patch_u_bc, binary)
    # splice the BC entry of the tool patch into the field file (type + parameters; value stays)
    if case.rotslip is not None:
        pname = case.mesh.patchName[case.rotslip.patch]
        entry = bc_text[bc_text.index(pname):]
        entry = entry[entry.index("{") + 1:entry.index("}")]
        params = "\n".join(l for l in entry.splitlines() if l.strip() and not l.strip().startswith("value"))
        txt = open(u_path, "rb").read()
        marker = (f"    {pname}\n    {{\n        type            fixedValue;\n").encode()
        assert marker in txt
        txt = txt.replace(marker, (f"    {pname}\n    {{\n" + params + "\n").encode())
        open(u_path, "wb").write(txt)
    handle.write_field(os.path.join(path, "0", "temp"), "temp", case.T_i, case.T_b, None, binary)
    handle.write_field(os.path.join(path, "0", "alpha.metal"), "alpha.metal", case.alpha1_i, case.alpha1_b, None, binary)
    handle.write_field(os.path.join(path, "0", "p"), "p", case.p_i, case.p_b, None, binary)


def _write_dicts(path: str, case: Case) -> None:
    import os

    m = case.mesh
    os.makedirs(os.path.join(path, "constant"), exist_ok=True)
    os.makedirs(os.path.join(path, "0"), exist_ok=True)
    p = case.params
    with open(os.path.join(path, "constant", "transportProperties"), "w") as fh:
        fh.write("// written by solidstateamfoam_b200.cases.write_case_dir\nphases (metal air);\n\nmetal\n{\n")
        fh.write(_visc_dict(p.visc1, p.mix.rho1, p.mix.cp1, p.mix.k1))
        fh.write("}\n\nair\n{\n")
        fh.write(_visc_dict(p.visc2, p.mix.rho2, p.mix.cp2, p.mix.k2))
        fh.write("}\n\nsigma 0.07;\n")
    s = p.stick
    patch = m.patchName[s.fricPatch] if s.fricPatch >= 0 else "zmax"
    om = "(" + " ".join(repr(x) for x in s.omega) + ")"
    with open(os.path.join(path, "constant", "frictionProperties"), "w") as fh:
        name = "constantSlipStick" if s.model == P.STICK_CONSTANT else "exponentialSlipStick"
        fh.write(f"stickModel {name};\nalphaName alpha.metal;\npName p;\nfricPatch {patch};\n\n{name}Coeffs\n{{\n")
        fh.write(f"    alphaName alpha.metal;\n    pName p;\n    fricPatch {patch};\n")
        if s.model == P.STICK_CONSTANT:
            fh.write(f"    betaTQ {s.betaTQ!r};\n    delta {s.delta!r};\n    omega {om};\n    muf {s.muf!r};\n")
        else:
            fh.write(f"    betaTQ {s.betaTQ!r};\n    omega {om};\n    omega0 {s.omega0!r};\n    delta0 {s.delta0!r};\n"
                     f"    Rs {s.Rs!r};\n    mu0 {s.mu0!r};\n    lambda {s.lambda_!r};\n")
        fh.write("}\n")
    with open(os.path.join(path, "0", "U"), "w") as fh:
        fh.write("dimensions [0 1 -1 0 0 0 0];\ninternalField uniform (0 0 0);\nboundaryField\n{\n")
        r = case.rotslip
        if r is not None:
            omr = "(" + " ".join(repr(x) for x in r.omega) + ")"
            vt = "(" + " ".join(repr(x) for x in r.Vt) + ")"
            fh.write(f"    {m.patchName[r.patch]}\n    {{\n")
            if r.model == P.ROTSLIP_CONSTANT:
                fh.write(f"        type constantRotationalSlip;\n        omega {omr};\n        delta {r.delta!r};\n")
            else:
                fh.write(f"        type exponentialRotationalSlip;\n        omega {omr};\n        omega0 {r.omega0!r};\n"
                         f"        delta0 {r.delta0!r};\n        R0 {r.R0!r};\n")
            fh.write(f"        Vt {vt};\n        value uniform (0 0 0);\n    }}\n")
        fh.write("}\n")
