"""GPU parity tests proper: the CUDA path, called through the C-ABI, against the oracle.

Bars (BASELINE.json north_star): integer tables bit-exact; grad U / strain rates bit-exact (same
summation order, no FMA); every other FP64 field <= 1e-11 relative per cell and boundary face.
"""
import math

import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

from helpers import EPS_FACTOR, SR_FACTOR, check_field, compare_ctx_oracle, oracle_case, oracle_update

pytestmark = pytest.mark.gpu

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)


def small_cases():
    return {
        "cfg1_hex": lambda: cases.cfg1(dims=(20, 16, 12)),
        "cfg1_sheared_graded": lambda: cases.cfg1(dims=(16, 12, 10), affine=SHEAR, grading=(2.0, 0.5, 3.0)),
        "cfg2_fks_kelvin": lambda: cases.cfg2(dims=(16, 16, 8)),
        "cfg2_fks_celsius": lambda: cases.cfg2(dims=(12, 10, 8), units="Celsius"),
        "cfg3_poly": lambda: cases.cfg3(base=(12, 12, 12)),
        "cfg4_nortonhoff": lambda: cases.cfg4(dims=(12, 12, 12)),
    }


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(small_cases()))
def test_integer_addressing_bit_exact(ctx, name):
    case = small_cases()[name]()
    ctx.load_case(case)
    oc = oracle_case(case)
    for which in (P.ADDR_OWNER_START, P.ADDR_EXTRA_START, P.ADDR_EXTRA_FACE, P.ADDR_EXTRA_OTHER,
                  P.ADDR_CELL_FACES_START, P.ADDR_CELL_FACES, P.ADDR_HALO_SEND_CELLS, P.ADDR_HALO_PATCH_OFFSETS,
                  P.ADDR_FACE_COLOUR):
        got, ref = ctx.addressing(which), oc.addressing(which)
        assert got.dtype == ref.dtype and got.shape == ref.shape, P.ADDR_NAMES[which]
        assert got.tobytes() == ref.tobytes(), f"{P.ADDR_NAMES[which]} differs"


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["direct", "tiled", "pipe"])
@pytest.mark.parametrize("name", list(small_cases()))
def test_fused_update_matches_oracle(ctx, name, variant):
    """both cell-kernel variants (one thread per cell with direct loads / tiled with bulk-async staging)"""
    case = small_cases()[name]()
    ctx.load_case(case)
    ctx.set_cell_kernel(variant)
    try:
        ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
    finally:
        ctx.set_cell_kernel(1)
    oc = oracle_update(case, O.FAITHFUL, ctx=ctx)
    report = {}
    compare_ctx_oracle(ctx, oc, case.outputs + [P.F_GRADU, P.F_U], report)
    print(name, {k: (f"{v[0]:.1e}", v[1]) for k, v in report.items()})


def test_tiled_kernel_capacity_fallback(ctx):
    """a polyhedral tile whose owned faces exceed the staged capacity must fall back to global loads;
    a mesh much larger than one tile exercises partial last tiles and cross-tile gathers"""
    case = cases.cfg3(base=(20, 20, 20))
    ctx.load_case(case)
    res = {}
    for variant in (0, 1, 2):
        ctx.set_cell_kernel(variant)
        ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
        res[variant] = {f: ctx.download(f) for f in case.outputs + [P.F_GRADU]}
    ctx.set_cell_kernel(1)
    for f in res[0]:
        for v in (1, 2):
            assert np.array_equal(res[0][f][0], res[v][f][0], equal_nan=True), (v, P.FIELD_NAMES[f])
            assert np.array_equal(res[0][f][1], res[v][f][1], equal_nan=True), (v, P.FIELD_NAMES[f])


@pytest.mark.parametrize("name", ["cfg1_sheared_graded", "cfg2_fks_kelvin", "cfg3_poly", "cfg4_nortonhoff"])
def test_staged_calls_match_oracle_and_fused(ctx, name):
    """The solver's own call sequence: thermo.correct() with the *stored* epsilonDotEq,
    then TEqn.H's strain rate / mu / friction.correct() / cp / k."""
    case = small_cases()[name]()
    ctx.load_case(case)
    oc = oracle_case(case)
    oc.set(P.F_U, *ctx.download(P.F_U))  # identical inputs on both sides (see helpers.oracle_update)
    p = case.params
    nh = p.visc1.model == P.VISC_NORTON_HOFF
    # previous time step's epsilonDotEq (createFluidFields.H:148-164 initial value): uniform 1.0
    nTot = case.mesh.nCells + case.mesh.nBoundaryFaces
    stale = np.full(nTot, 1.0)
    ctx.upload(P.F_EPSDOT, stale[: case.mesh.nCells], stale[case.mesh.nCells:])
    oc.set(P.F_EPSDOT, stale[: case.mesh.nCells], stale[case.mesh.nCells:])
    # solveFluid.H:52  thermo.correct()
    ctx.thermo_correct(p.visc1, p.visc2, p.mix)
    oc.viscosity_correct(1, p.visc1)
    oc.viscosity_correct(2, p.visc2)
    oc.mixture("calc_nu", p.mix)
    compare_ctx_oracle(ctx, oc, [P.F_NU1, P.F_NU2, P.F_NU] + ([] if nh else [P.F_SIGMAF, P.F_SIGMAY]))
    # TEqn.H:3-15
    ctx.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    ctx.mixture("mu", p.mix)
    oc.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    oc.mixture("mu", p.mix)
    if p.stick.fricPatch >= 0 or True:
        ctx.friction_correct(p.stick)
        oc.friction_correct(p.stick)
    ctx.mixture("cp", p.mix)
    ctx.mixture("k", p.mix)
    ctx.mixture("rho", p.mix)
    ctx.solver_rho_rhocp(p.mix)
    ctx.grad_u()
    for w in ("cp", "k", "rho"):
        oc.mixture(w, p.mix)
    oc.solver_rho_rhocp(p.mix)
    fields = [P.F_EPSDOT, P.F_MUEFF, P.F_QVISC, P.F_CP, P.F_K, P.F_RHO, P.F_RHO_SOLVER, P.F_RHOCP, P.F_GRADU]
    if p.stick.fricPatch >= 0:
        fields.append(P.F_QFRIC)
    compare_ctx_oracle(ctx, oc, fields)
    # now the staged sequence in fused order must equal the fused kernel bit for bit
    ctx.thermo_correct(p.visc1, p.visc2, p.mix)
    ctx.mixture("mu", p.mix)
    ctx.friction_correct(p.stick)
    staged = {f: ctx.download(f) for f in case.outputs}
    ctx.update_fused(p, 0)
    for f in case.outputs:
        gi, gb = ctx.download(f)
        assert np.array_equal(gi, staged[f][0], equal_nan=True), P.FIELD_NAMES[f]
        assert np.array_equal(gb, staged[f][1], equal_nan=True), P.FIELD_NAMES[f] + " (boundary)"


def test_rotational_slip_bcs(ctx):
    for case in (cases.cfg1(dims=(10, 9, 8)), cases.cfg3(base=(8, 8, 8))):
        ctx.load_case(case)
        oc = oracle_case(case)
        gi, gb = ctx.download(P.F_U)
        ref = oc.field(P.F_U)
        # exp() in the exponential variant is the only non-IEEE op
        bit = case.rotslip.model == P.ROTSLIP_CONSTANT
        check_field("U_b", gb, ref[case.mesh.nCells:], bit)
        assert np.array_equal(gi, ref[: case.mesh.nCells])


def test_friction_patch_centre(ctx):
    case = cases.cfg1(dims=(12, 10, 8), affine=SHEAR)
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    oc = oracle_update(case)
    c_gpu, a_gpu = ctx.friction_patch_centre()
    c_ref, a_ref = oc.friction_patch_centre()
    assert np.array_equal(c_gpu, c_ref) and a_gpu == a_ref  # serial in-order sum: bit-exact


def test_ten_updates_cfg1(ctx):
    """cfg1: 10 updates with U, T advanced analytically per step, compared every step."""
    case = cases.cfg1(dims=(16, 16, 8))
    ctx.mesh_set(case.mesh, case.patch_u_bc)
    for step in range(10):
        cases.advance(case, step)
        ctx.upload(P.F_U, case.U_i, case.U_b)
        ctx.upload(P.F_T, case.T_i, case.T_b)
        ctx.upload(P.F_ALPHA1, case.alpha1_i, case.alpha1_b)
        ctx.upload(P.F_P, case.p_i, case.p_b)
        ctx.bc_rotational_slip(case.rotslip)
        ctx.update_fused(case.params, 0)
        oc = oracle_update(case)
        compare_ctx_oracle(ctx, oc, case.outputs)


def test_error_behaviour(ctx):
    case = cases.cfg4(dims=(6, 6, 6))
    ctx.load_case(case)
    # NortonHoff registers no sigmay: calcQfric's lookup aborts in the reference (SURVEY D-1)
    ctx.update_fused(case.params, 0)
    stick = P.constant_slip_stick(case.mesh.patch_id("zmax"), 0.9, 0.5, cases.OMEGA, 0.3)
    with pytest.raises(capi.SsamError) as e:
        ctx.friction_correct(stick)
    assert e.value.status == P.ERR_MISSING_FIELD
    bad = P.ViscosityParams(model=17)
    with pytest.raises(capi.SsamError) as e:
        ctx.viscosity_correct(1, bad)
    assert e.value.status == P.ERR_UNKNOWN_MODEL
    mix = P.mixture(1.0, 1.0, 1.0, 1.0, P.k_cubic(1, 1, 1, 1, 0, 1, units="Rankine"), P.k_constant(1.0))
    with pytest.raises(capi.SsamError) as e:
        ctx.mixture("k", mix)
    assert e.value.status == P.ERR_UNITS
    badstick = P.StickParams(model=7, fricPatch=0, betaTQ=1.0)
    with pytest.raises(capi.SsamError) as e:
        ctx.friction_correct(badstick)
    assert e.value.status == P.ERR_UNKNOWN_MODEL


def test_edge_values(ctx):
    """eDot = 0 and below eDotMin, T below Ta / above Tm, alpha outside [0,1], Celsius clamp at 0 (NaN)."""
    case = cases.cfg2(dims=(8, 8, 8), units="Celsius")
    n = case.mesh.nCells
    case.T_i[:8] = [0.0, 100.0, 292.9, 293.0, 855.0, 3000.0, 273.0, 1e-3]
    case.alpha1_i[:6] = [-0.5, 0.0, 1e-300, 1.0, 1.5, 0.5]
    case.U_i[:] = 0.0          # eDot exactly 0 in the interior
    case.U_b[:] = 0.0
    case.rotslip = None
    # Celsius fit with Tmin = 0: T = 273 K clamps to exactly 0 -> k0*(0/0) = NaN (SURVEY D-4)
    case.params.mix.k1.Tmin = 0.0
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    oc = oracle_update(case)
    compare_ctx_oracle(ctx, oc, case.outputs)
    ki, _ = ctx.download(P.F_K)
    assert np.isnan(ki[6]) and np.isnan(oc.field(P.F_K)[6])


@pytest.mark.parametrize("layout", [P.LAYOUT_CALLER, P.LAYOUT_COMPACT], ids=["caller", "compact"])
@pytest.mark.parametrize("chunks", [0, 3, 7])
def test_host_staged_update_equals_device_resident(chunks, layout):
    """ssam_update_fused_host (the e2e path): plain and pipelined over chunks, in both cell layouts, bit-identical
    to the device-resident update on the same inputs"""
    case = cases.cfg2(dims=(64, 32, 24))  # 192 tiles, bandwidth 8 tiles
    with capi.Context(0) as ctx:
        ctx.set_cell_layout(layout)
        _host_staged_check(ctx, case, chunks)
        assert ctx.layout_is_compact() == (layout == P.LAYOUT_COMPACT)


def _host_staged_check(ctx, case, chunks):
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    ref = {f: ctx.download(f) for f in case.outputs}
    U_i, U_b = ctx.download(P.F_U)
    ctx.set_host_chunks(chunks)
    try:
        # scribble over the device inputs first: the host path must bring its own
        z = np.zeros(case.mesh.nCells)
        ctx.upload(P.F_T, z + 1.0, None)
        ctx.upload(P.F_ALPHA1, z, None)
        got = ctx.host_update(case, case.outputs, U_b=U_b)
    finally:
        ctx.set_host_chunks(-1)
    for f in case.outputs:
        assert np.array_equal(got[f][0], ref[f][0], equal_nan=True), (chunks, P.FIELD_NAMES[f])
        assert np.array_equal(got[f][1], ref[f][1], equal_nan=True), (chunks, P.FIELD_NAMES[f] + " boundary")


@pytest.mark.parametrize("dims", [(1, 1, 1), (2, 1, 1), (1, 1, 3), (257, 1, 1), (3, 2, 1)])
def test_degenerate_meshes(ctx, dims):
    """one cell (no internal face), one row crossing a tile boundary, single layers"""
    case = cases.cfg1(dims=dims)
    for variant in (0, 1, 2):
        ctx.load_case(case)
        ctx.set_cell_kernel(variant)
        ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
        oc = oracle_update(case, O.FAITHFUL, ctx=ctx)
        compare_ctx_oracle(ctx, oc, case.outputs + [P.F_GRADU])
        for which in (P.ADDR_OWNER_START, P.ADDR_EXTRA_START, P.ADDR_EXTRA_FACE, P.ADDR_CELL_FACES, P.ADDR_FACE_COLOUR):
            assert ctx.addressing(which).tobytes() == oc.addressing(which).tobytes()
    ctx.set_cell_kernel(1)


@pytest.mark.parametrize("name", ["cfg1_sheared_graded", "cfg3_poly"])
def test_stress_strain_post_processing(ctx, name):
    """fluid/calculateStrain.H: Z, sigmaf, dEpsilonEq, epsilonEq += , sigmavm"""
    case = small_cases()[name]()
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    oc = oracle_update(case, O.FAITHFUL, ctx=ctx)
    n, nb = case.mesh.nCells, case.mesh.nBoundaryFaces
    prgh = 1e5 * (1.0 + 0.05 * np.random.default_rng(5).standard_normal(n + nb))
    ctx.upload(P.F_PRGH, prgh[:n], prgh[n:])
    oc.set(P.F_PRGH, prgh[:n], prgh[n:])
    # identical inputs on both sides: muEff from the fused update agrees with the oracle's only to ~1e-14,
    # and sigmavm subtracts p_rgh*I (|p| ~ 1e5 against shear stresses down to 1e-3), which would amplify it
    ctx.upload(P.F_MUEFF, oc.internal(P.F_MUEFF), oc.boundary(P.F_MUEFF))
    v = case.params.visc1
    prm = P.StressStrainParams(Q=v.Q, R=v.R, sigmaR=v.sigmaR, A=v.A, n=v.n, deltaT=1e-3)
    for _ in range(2):  # the second call accumulates epsilonEq
        ctx.stress_strain(prm)
        oc.stress_strain(prm)
    for f in (P.F_Z, P.F_SIGMAF_PP, P.F_DEPSEQ, P.F_EPSEQ, P.F_GRADU):
        gi, gb = ctx.download(f)
        check_field(P.FIELD_NAMES[f], np.concatenate([gi, gb]), oc.field(f), f in (P.F_GRADU, P.F_DEPSEQ, P.F_EPSEQ))
    gi, gb = ctx.download(P.F_SIGMAVM)
    assert np.array_equal(np.concatenate([gi, gb]), oc.field(P.F_SIGMAVM))  # pure IEEE arithmetic


@pytest.mark.parametrize("variant", [0, 1, 2], ids=["direct", "tiled", "pipe"])
@pytest.mark.parametrize("name", ["cfg1_hex", "cfg1_sheared_graded", "cfg2_fks_kelvin", "cfg3_poly"])
def test_interface_correct_matches_oracle(ctx, name, variant):
    """interfaceProperties::correct(): gradAlpha, nHatf, K -- pure IEEE arithmetic in the
    reference's summation order, so every value is bit-exact."""
    case = small_cases()[name]()
    m = case.mesh
    ctx.load_case(case)
    oc = oracle_case(case, apply_bc=False)
    dn = ctx.interface_delta_n()
    assert dn == O.interface_delta_n([oc])
    ctx.set_cell_kernel(variant)
    try:
        ctx.interface_correct(dn)
    finally:
        ctx.set_cell_kernel(1)
    oc.interface_correct(dn)
    for f in (P.F_GRADALPHA, P.F_CURVATURE):
        gi, gb = ctx.download(f)
        check_field(P.FIELD_NAMES[f], np.concatenate([gi, gb]), oc.field(f), True)
    assert np.array_equal(ctx.nhatf(m.nInternalFaces, m.nBoundaryFaces), oc.nhatf())
    assert np.count_nonzero(oc.internal(P.F_CURVATURE)) > 0


def test_interface_curvature_of_a_sphere(ctx):
    """analytic pin on the GPU itself: K = 2/R on a spherical tanh interface"""
    n = 40
    m = M.hex_block(n, n, n, origin=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0))
    R, width = 0.6, 3.0 * 2.0 / n
    xc = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    r = np.sqrt((xc ** 2).sum(1))
    a = 0.5 * (1.0 - np.tanh((r - R) / width))
    ctx.mesh_set(m)
    ctx.upload(P.F_ALPHA1, a[: m.nCells], a[m.nCells:])
    ctx.interface_correct(ctx.interface_delta_n())
    K, _ = ctx.download(P.F_CURVATURE)
    rc = r[: m.nCells]
    band = np.abs(rc - R) < 0.5 * width
    assert np.allclose(K[band] * rc[band], 2.0, rtol=0.06)


def test_interface_requires_alpha(ctx):
    ctx.mesh_set(M.hex_block(3, 3, 3))
    with pytest.raises(capi.SsamError) as e:
        ctx.interface_correct(1e-7)
    assert e.value.status == P.ERR_MISSING_FIELD
    with pytest.raises(capi.SsamError):
        ctx.nhatf(54, 54)


@pytest.mark.parametrize("name", ["cfg1_sheared_graded", "cfg2_fks_kelvin", "cfg3_poly"])
def test_muf_nuf_face_fields(ctx, name):
    """incompressibleTwoPhaseThermalMixture::muf()/nuf() (…Mixture.C:163-202): IEEE-only, bit-exact given nu1/nu2"""
    case = small_cases()[name]()
    m = case.mesh
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    oc = oracle_update(case, O.FAITHFUL, ctx=ctx)
    for f in (P.F_NU1, P.F_NU2):  # identical inputs on both sides (nu1 agrees to ~1e-15 only)
        ctx.upload(f, oc.internal(f), oc.boundary(f))
    for nuf, ff in ((False, P.FF_MUF), (True, P.FF_NUF)):
        ctx.mixture_face_visc(case.params.mix, nuf)
        ref = oc.mixture_face_visc(case.params.mix, nuf)
        got = ctx.face_field(ff, m.nInternalFaces, m.nBoundaryFaces)
        assert np.array_equal(got, ref)
        assert np.isfinite(ref).all() and (ref > 0).all()
    # kf = thermo.k() on the faces (fvm::laplacian(kf, temp), fluid/TEqn.H:22): linear interpolation, bit-exact
    ctx.upload(P.F_K, oc.internal(P.F_K), oc.boundary(P.F_K))
    ctx.face_interpolate(P.F_K, P.FF_KF)
    assert np.array_equal(ctx.face_field(P.FF_KF, m.nInternalFaces, m.nBoundaryFaces), oc.face_interpolate(P.F_K))
    with pytest.raises(capi.SsamError):
        ctx.face_field(7, 1, 1)


def test_field_min_max(ctx):
    """"Min/max T" of fluid/TEqn.H:42-43, plus odd sizes for the two-pass reduction"""
    case = small_cases()["cfg1_sheared_graded"]()
    ctx.load_case(case)
    oc = oracle_case(case, apply_bc=False)
    assert ctx.field_min_max(P.F_T) == O.field_min_max([oc], P.F_T)
    rng = np.random.default_rng(17)
    for dims in [(1, 1, 1), (7, 3, 1), (33, 9, 5), (64, 64, 20)]:
        m = M.hex_block(*dims)
        ctx.mesh_set(m)
        v = rng.standard_normal(m.nCells + m.nBoundaryFaces) * 10.0 ** rng.integers(-5, 5)
        ctx.upload(P.F_T, v[: m.nCells], v[m.nCells:])
        assert ctx.field_min_max(P.F_T) == (v.min(), v.max())
    with pytest.raises(capi.SsamError):
        ctx.field_min_max(P.F_U)
    with pytest.raises(capi.SsamError):
        ctx.field_min_max(P.F_QFRIC)  # never set on this mesh


def test_gpu_against_frozen_reference_source_outputs(ctx):
    """The CUDA path against outputs of the reference's OWN model sources (tests/golden/reference_sources.npz,
    frozen from oracle/_ref by tests/golden/gen_reference_golden.py) -- no oracle in between; <= 1e-11 relative."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from gen_reference_golden import GOLDEN_CASES, inputs_digest

    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sources.npz"))
    names = {"nu1": P.F_NU1, "nu2": P.F_NU2, "sigmaf": P.F_SIGMAF, "sigmay": P.F_SIGMAY, "nu": P.F_NU, "mu": P.F_MUEFF,
             "rho": P.F_RHO, "cp": P.F_CP, "k": P.F_K, "qvisc": P.F_QVISC, "qfric": P.F_QFRIC}
    checked = 0
    for cname, mk in GOLDEN_CASES.items():
        case = mk()
        assert inputs_digest(case) == bytes(golden[cname + "/inputs_sha256"]).decode()
        ctx.load_case(case)
        if cname + "/bc_Ub" in golden.files:  # *RotationalSlip*::updateCoeffs() of the reference itself
            _, ub = ctx.download(P.F_U)
            sl = case.mesh.patch_slice_b(case.mesh.patch_id("zmax"))
            check_field(f"{cname}/U_b vs reference sources", ub[sl].reshape(-1), golden[cname + "/bc_Ub"],
                        case.rotslip.model == P.ROTSLIP_CONSTANT)
            checked += 1
        ctx.update_fused(case.params, 0)
        for key, f in names.items():
            if cname + "/" + key not in golden.files:
                continue
            gi, gb = ctx.download(f)
            check_field(f"{cname}/{key} vs reference sources", np.concatenate([gi, gb]), golden[cname + "/" + key],
                        f in (P.F_RHO, P.F_CP, P.F_NU2))
            checked += 1
        for nuf, ff, key in ((False, P.FF_MUF, "muf"), (True, P.FF_NUF, "nuf")):
            ctx.mixture_face_visc(case.params.mix, nuf)
            got = ctx.face_field(ff, case.mesh.nInternalFaces, case.mesh.nBoundaryFaces)
            check_field(f"{cname}/{key} vs reference sources", got, golden[cname + "/" + key], False)
            checked += 1
    assert checked >= 50


@pytest.mark.parametrize("name", ["hex_natural", "hex_sheared", "poly", "nortonhoff"])
def test_tile_compact_layout_is_bit_identical(name):
    """The library's own cell order (SSAM_LAYOUT_COMPACT) is a pure data-layout change: every field, every face
    field and every exposed integer table equals the caller-order layout bit for bit."""
    mk = {"hex_natural": lambda: cases.cfg2(dims=(64, 32, 16)),
          "hex_sheared": lambda: cases.cfg1(dims=(40, 24, 20), affine=SHEAR, grading=(2.0, 0.5, 3.0)),
          "poly": lambda: cases.cfg3(base=(16, 16, 16)),
          "nortonhoff": lambda: cases.cfg4(dims=(33, 21, 17))}[name]
    case = mk()
    m = case.mesh
    res = {}
    for mode in (P.LAYOUT_CALLER, P.LAYOUT_COMPACT):
        with capi.Context(0) as c:
            c.set_cell_layout(mode)
            c.load_case(case)
            assert c.layout_is_compact() == (mode == P.LAYOUT_COMPACT)
            order = c.addressing(P.ADDR_CELL_ORDER)
            assert np.array_equal(np.sort(order), np.arange(m.nCells))
            forder = c.addressing(P.ADDR_FACE_ORDER)
            assert np.array_equal(np.sort(forder), np.arange(m.nInternalFaces))
            if mode == P.LAYOUT_COMPACT:
                assert not np.array_equal(order, np.arange(m.nCells))
                # device faces are grouped by (device position of) the caller's owner, caller order within
                inv = np.empty(m.nCells, dtype=np.int64)
                inv[order] = np.arange(m.nCells)
                own_dev = inv[m.owner[forder]]
                assert (np.diff(own_dev) >= 0).all()
                same = np.diff(own_dev) == 0
                assert (np.diff(forder)[same] > 0).all()
            else:
                assert np.array_equal(forder, np.arange(m.nInternalFaces))
            out = {}
            for variant in (0, 1, 2):
                c.set_cell_kernel(variant)
                c.update_fused(case.params, capi.FUSED_WRITE_GRAD)
                for f in case.outputs + [P.F_GRADU, P.F_U]:
                    gi, gb = c.download(f)
                    out[(variant, f)] = np.concatenate([gi, gb])
            c.set_cell_kernel(1)
            c.interface_correct(c.interface_delta_n())
            for f in (P.F_GRADALPHA, P.F_CURVATURE):
                gi, gb = c.download(f)
                out[("iface", f)] = np.concatenate([gi, gb])
            out["nhatf"] = c.face_field(P.FF_NHATF, m.nInternalFaces, m.nBoundaryFaces)
            for nuf, ff in ((False, P.FF_MUF), (True, P.FF_NUF)):
                c.mixture_face_visc(case.params.mix, nuf)
                out[("ff", ff)] = c.face_field(ff, m.nInternalFaces, m.nBoundaryFaces)
            c.strain_rate(SR_FACTOR, P.F_STRAINRATE)
            out["sr"] = np.concatenate(c.download(P.F_STRAINRATE))
            out["minmax"] = np.array(c.field_min_max(P.F_T))
            for which in range(P.ADDR_CELL_ORDER):
                out[("addr", which)] = c.addressing(which)
            # host-staged entry point (falls back to the unpipelined path under the compact layout)
            _, ub = c.download(P.F_U)  # with the BC applied, as the device-resident update saw it
            got = c.host_update(case, [P.F_NU, P.F_QVISC], U_b=ub)
            out["host_nu"], out["host_qvisc"] = np.concatenate(got[P.F_NU]), np.concatenate(got[P.F_QVISC])
            assert np.array_equal(out["host_nu"], out[(1, P.F_NU)])
            res[mode] = out
    a, b = res[P.LAYOUT_CALLER], res[P.LAYOUT_COMPACT]
    assert a.keys() == b.keys()
    for k in a:
        same = (a[k] == b[k]) | ((a[k] != a[k]) & (b[k] != b[k]))
        assert same.all(), f"{k}: {np.count_nonzero(~same)} values differ between the two layouts"


@pytest.mark.parametrize("name", ["cfg1_sheared_graded", "cfg2_fks_kelvin", "cfg3_poly"])
def test_div_dev_rho_reff_explicit(ctx, name):
    """-fvc::div((rho*nuEff)*dev2(T(grad U))) (fluid/UEqn.H:13, explicit part): IEEE-only, bit-exact"""
    case = small_cases()[name]()
    ctx.load_case(case)
    ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
    ctx.solver_rho_rhocp(case.params.mix)
    oc = oracle_update(case, O.FAITHFUL, ctx=ctx)
    oc.solver_rho_rhocp(case.params.mix)
    ctx.upload(P.F_NU, oc.internal(P.F_NU), oc.boundary(P.F_NU))  # identical operands (nu agrees to ~1e-15 only)
    ctx.div_dev_rho_reff_explicit()
    oc.div_dev_rho_reff_explicit()
    gi, gb = ctx.download(P.F_DIVDEVRHOREFF)
    check_field("divDevRhoReff (explicit)", np.concatenate([gi, gb]), oc.field(P.F_DIVDEVRHOREFF), True)
    assert np.abs(gi).max() > 0


def test_rerun_is_bitwise_deterministic(ctx):
    """No FP64 atomics anywhere on the path: repeated updates on the same inputs give identical bits (this is also
    the race detector of the tiled kernels: a missed barrier or a stale shared-memory tile would show up here)."""
    case = cases.cfg3(base=(16, 16, 16))
    ctx.load_case(case)
    first = None
    for rep in range(4):
        ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
        ctx.interface_correct(ctx.interface_delta_n())
        snap = [np.concatenate(ctx.download(f)) for f in case.outputs + [P.F_GRADU, P.F_CURVATURE, P.F_GRADALPHA]]
        if first is None:
            first = snap
        else:
            for a, b in zip(first, snap):
                assert np.array_equal(a, b, equal_nan=True), rep


@pytest.mark.parametrize("name", ["cfg1_sheared_graded", "cfg2_fks_kelvin", "cfg3_poly"])
@pytest.mark.parametrize("coeffs", [(1.0, 0.0, 0.0), (1.0, 0.25, 0.0), (0.7, 0.3, 0.6)], ids=["plain", "ic", "ic+sc"])
def test_alpha_compression_coefficient(ctx, name, coeffs):
    """multiRegionVoF/alphaEqn.H:58-87: phic = cAlpha|phi/magSf| with the optional isotropic (icAlpha) and
    shear (scAlpha) compression terms; IEEE-only arithmetic in the reference's operator order -> bit-exact"""
    case = small_cases()[name]()
    m = case.mesh
    ctx.load_case(case)
    oc = oracle_case(case)
    oc.set(P.F_U, *ctx.download(P.F_U))
    cA, ic, sc = coeffs
    phi = np.random.default_rng(17).normal(size=m.nFaces) * 1e-6
    if sc > 0:
        ctx.grad_u()
        oc.grad_u()
        gi, gb = ctx.download(P.F_GRADU)
        assert np.array_equal(np.concatenate([gi, gb]), oc.field(P.F_GRADU), equal_nan=True)
    ctx.face_field_upload(P.FF_PHI, phi)
    ctx.alpha_compression_coeff(cA, ic, sc)
    got = ctx.face_field(P.FF_PHIC, m.nInternalFaces, m.nBoundaryFaces)
    ref = oc.alpha_compression_coeff(phi, cA, ic, sc)
    assert np.array_equal(got, ref), f"{np.count_nonzero(got != ref)} of {got.size} faces differ"
    assert np.all(got[m.nInternalFaces:] == 0.0) and np.abs(got[: m.nInternalFaces]).min() > 0


def test_alpha_compression_coefficient_needs_its_inputs(ctx):
    case = small_cases()["cfg2_fks_kelvin"]()
    ctx.load_case(case)
    with pytest.raises(capi.SsamError) as e:
        ctx.alpha_compression_coeff(1.0, 0.0, 0.0)
    assert "phi" in str(e.value)
