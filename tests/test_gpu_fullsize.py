"""BASELINE.json's full sizes on the GPU: direct comparison with the oracle where it finishes in seconds
(cfg2, 8.4 M cells), plus size-independent properties (exact gradient of a linear field, zero strain of a
rigid rotation, rerun determinism, staged == fused) at cfg2 / cfg4 scale."""
import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import params as P

from helpers import EPS_FACTOR, compare_ctx_oracle, oracle_update

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


def test_cfg2_full_size_matches_oracle(ctx):
    """hex 256x256x128, FKS + cubic k(T): every cell and boundary face of every output field"""
    case = cases.cfg2()
    assert case.mesh.nCells == 8388608 and case.mesh.nInternalFaces == 25034752
    ctx.load_case(case)
    ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
    oc = oracle_update(case, O.FUSED, ctx=ctx)
    report = {}
    compare_ctx_oracle(ctx, oc, case.outputs + [P.F_GRADU], report)
    print({k: (f"{v[0]:.1e}", v[1]) for k, v in report.items()})
    # rerun determinism: no atomics anywhere on the path, so a second update is bit-identical
    first = {f: ctx.download(f) for f in (P.F_EPSDOT, P.F_NU, P.F_QVISC, P.F_QFRIC)}
    ctx.update_fused(case.params, 0)
    for f, (i, b) in first.items():
        i2, b2 = ctx.download(f)
        assert np.array_equal(i, i2, equal_nan=True) and np.array_equal(b, b2, equal_nan=True)
    # integer tables at full size
    for which in (P.ADDR_OWNER_START, P.ADDR_EXTRA_START, P.ADDR_EXTRA_FACE, P.ADDR_EXTRA_OTHER):
        assert ctx.addressing(which).tobytes() == oc.addressing(which).tobytes(), P.ADDR_NAMES[which]


def test_cfg4_scale_properties(ctx):
    """16.8 M cells, NortonHoff: linear field -> exact gradient; rigid rotation -> zero strain rate;
    staged calls == fused kernel"""
    case = cases.cfg4()
    m = case.mesh
    assert m.nCells == 16777216
    G = np.array([[1.5, -2.0, 0.25], [0.5, 3.0, -1.0], [-0.75, 0.125, 2.0]])
    X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    U = X @ G + np.array([0.1, -0.2, 0.3])
    ctx.mesh_set(m, np.full(m.nPatches, P.BC_VALUE, dtype=np.int32))
    ctx.upload(P.F_U, U[: m.nCells], U[m.nCells:])
    ctx.grad_u()
    gi, gb = ctx.download(P.F_GRADU)
    assert np.abs(gi.reshape(-1, 3, 3) - G[None]).max() <= 1e-9 * np.abs(G).max()   # coordinates ~6e-2, cells 2.5e-4
    assert np.abs(gb.reshape(-1, 3, 3) - G[None]).max() <= 1e-9 * np.abs(G).max()
    om = np.array([0.3, -0.2, 41.9])
    U = np.cross(np.broadcast_to(om, X.shape), X - X.mean(axis=0))
    ctx.upload(P.F_U, U[: m.nCells], U[m.nCells:])
    ctx.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    ei, eb = ctx.download(P.F_EPSDOT)
    assert max(ei.max(), eb.max()) <= 1e-9 * np.linalg.norm(om)
    # the configured case: fused vs staged, bit for bit
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    fused = {f: ctx.download(f) for f in case.outputs}
    p = case.params
    ctx.strain_rate(EPS_FACTOR, P.F_EPSDOT)
    ctx.thermo_correct(p.visc1, p.visc2, p.mix)
    ctx.mixture("mu", p.mix)
    ctx.friction_correct(p.stick)
    for w in ("cp", "k", "rho"):
        ctx.mixture(w, p.mix)
    for f in case.outputs:
        i, b = ctx.download(f)
        assert np.array_equal(i, fused[f][0], equal_nan=True), P.FIELD_NAMES[f]
        assert np.array_equal(b, fused[f][1], equal_nan=True), P.FIELD_NAMES[f]
    nu = fused[P.F_NU1][0]
    assert np.isfinite(nu).all() and nu.min() >= p.visc1.nuMin and nu.max() <= p.visc1.nuMax


def test_cfg3_polyhedral_mid_size(ctx):
    """~700k-cell polyhedral mesh (two refinement levels): both cell-kernel variants against the oracle"""
    case = cases.cfg3(base=(64, 64, 64))
    ctx.load_case(case)
    for variant in (0, 1, 2):
        ctx.set_cell_kernel(variant)
        ctx.update_fused(case.params, capi.FUSED_WRITE_GRAD)
        oc = oracle_update(case, O.FUSED, ctx=ctx)
        compare_ctx_oracle(ctx, oc, case.outputs + [P.F_GRADU])
    ctx.set_cell_kernel(1)
