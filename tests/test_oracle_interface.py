"""interfaceProperties::correct() restatement pinned by analytic cases (CPU only).

The class is OpenFOAM library code, absent from the reference tree and not compilable here, so -- like the
rest of the oracle -- it is pinned by closed-form answers rather than by reference output: a linear alpha
field (exact gradient, zero curvature), a spherical interface (K = 2/R), a cylinder (K = 1/R), and the
decomposed evaluation agreeing with the serial one.
"""
import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P


def _case(m, alpha_fn):
    oc = O.OracleCase(m)
    xc = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    a = alpha_fn(xc)
    oc.set(P.F_ALPHA1, a[: m.nCells], a[m.nCells:])
    return oc


def test_delta_n():
    m = M.hex_block(6, 5, 4, length=(0.6, 0.5, 0.4))
    oc = O.OracleCase(m)
    d = O.interface_delta_n([oc])
    assert d == pytest.approx(1e-8 / 0.1, rel=1e-13)  # cube root of the 0.001 cell volume


def test_linear_alpha_gives_exact_gradient_and_zero_curvature():
    m = M.hex_block(10, 8, 6, length=(1.0, 0.8, 0.6))
    a = np.array([0.3, -0.2, 0.5])
    oc = _case(m, lambda x: 0.4 + x @ a)
    dn = O.interface_delta_n([oc])
    oc.interface_correct(dn)
    g = oc.field(P.F_GRADALPHA)
    assert np.allclose(g, a, rtol=0, atol=1e-12)  # cells and (corrected) boundary values
    nh = oc.nhatf()
    assert np.allclose(nh, (m.Sf @ a) / np.linalg.norm(a), rtol=1e-6, atol=1e-15)  # deltaN/|grad alpha| ~ 2e-7
    K = oc.internal(P.F_CURVATURE)
    assert np.abs(K).max() < 1e-9
    # boundary values of K: the adjacent cell's value
    assert np.array_equal(oc.boundary(P.F_CURVATURE), K[m.owner[m.nInternalFaces:]])


@pytest.mark.parametrize("shape,expect", [("sphere", 2.0), ("cylinder", 1.0)])
def test_curvature_of_sphere_and_cylinder(shape, expect):
    n = 40
    m = M.hex_block(n, n, n, origin=(-1.0, -1.0, -1.0), length=(2.0, 2.0, 2.0))
    R, width = 0.6, 3.0 * 2.0 / n

    def radius(x):
        return np.sqrt((x[:, :2] ** 2).sum(1)) if shape == "cylinder" else np.sqrt((x ** 2).sum(1))

    oc = _case(m, lambda x: 0.5 * (1.0 - np.tanh((radius(x) - R) / width)))
    oc.interface_correct(O.interface_delta_n([oc]))
    r = radius(m.C)
    band = np.abs(r - R) < 0.5 * width
    if shape == "cylinder":
        band &= np.abs(m.C[:, 2]) < 0.7
    K = oc.internal(P.F_CURVATURE)[band]
    assert band.sum() > 500
    # K = -div(nHat), nHat = grad(alpha)/|grad(alpha)| pointing into phase 1 (the inside): K = +(d-1)/r
    assert np.allclose(K * r[band], expect, rtol=0.06)  # second-order scheme on a 40^3 grid
    assert abs((K * r[band]).mean() - expect) < 0.02 * expect
    # unit normals: |nHatf| <= |Sf|
    assert (np.abs(oc.nhatf()) <= m.magSf * (1 + 1e-12)).all()


def test_sharp_step_has_no_nan():
    m = M.hex_block(8, 8, 8)
    oc = _case(m, lambda x: (x[:, 2] < 0.5 * x[:, 2].max()).astype(float))
    oc.interface_correct(O.interface_delta_n([oc]))
    for f in (P.F_GRADALPHA, P.F_CURVATURE):
        assert np.isfinite(oc.field(f)).all()
    assert np.isfinite(oc.nhatf()).all()
    g = oc.internal(P.F_GRADALPHA)
    assert (g[:, 2] <= 0).all() and np.count_nonzero(g[:, 2]) == 2 * 64  # the two cell layers at the step


def test_decomposed_equals_serial():
    dims = (10, 8, 12)
    L = (1.0, 0.8, 1.2)
    h = M.hex_block(*dims, length=L, affine=(1.0, 0.2, 0.1, 0.0, 1.0, 0.15, 0.0, 0.0, 1.0), _keep=True)

    def alpha(x):
        return 0.5 * (1.0 - np.tanh((np.sqrt(((x - 0.55) ** 2).sum(1)) - 0.35) / 0.2))

    rng = np.random.default_rng(3)
    nc = dims[0] * dims[1] * dims[2]
    cell_proc = rng.integers(0, 3, size=nc // 8 + 1)[np.arange(nc) // 8]
    whole, parts = M.decompose(h, cell_proc, 3)
    serial = _case(whole, alpha)
    dn = O.interface_delta_n([serial])
    serial.interface_correct(dn)
    ocs = [_case(p, alpha) for p in parts]
    assert O.interface_delta_n(ocs) == pytest.approx(dn, rel=1e-14)
    O.interface_correct_multi(ocs, dn)
    Ks, Gs = serial.internal(P.F_CURVATURE), serial.internal(P.F_GRADALPHA)
    for r, (oc, p) in enumerate(zip(ocs, parts)):
        cells = np.nonzero(cell_proc == r)[0]  # decomposePar keeps the global cell order within a rank
        assert np.array_equal(p.C, whole.C[cells])
        assert np.allclose(oc.internal(P.F_GRADALPHA), Gs[cells], rtol=1e-11, atol=1e-12)
        assert np.allclose(oc.internal(P.F_CURVATURE), Ks[cells], rtol=1e-9, atol=1e-8)
        # processor-patch boundary values are the neighbour rank's cell values
        for q in range(p.nPatches):
            if p.patchKind[q] != 1:
                continue
            s = slice(p.patchStart[q] - p.nInternalFaces, p.patchStart[q] - p.nInternalFaces + p.patchSize[q])
            assert np.isfinite(oc.boundary(P.F_CURVATURE)[s]).all()


def test_muf_nuf_closed_form():
    """muf()/nuf() (…Mixture.C:163-202) on a uniform mesh: face value = mean of the two cell values"""
    from solidstateamfoam_b200 import cases

    case = cases.cfg1(dims=(6, 5, 4))
    m = case.mesh
    oc = O.OracleCase(m)
    rng = np.random.default_rng(2)
    n = m.nCells + m.nBoundaryFaces
    a, n1, n2 = rng.uniform(-0.2, 1.2, n), rng.uniform(1e-6, 1e3, n), np.full(n, 1.5e-5)
    oc.set(P.F_ALPHA1, a[: m.nCells], a[m.nCells:])
    oc.set(P.F_NU1, n1[: m.nCells], n1[m.nCells:])
    oc.set(P.F_NU2, n2[: m.nCells], n2[m.nCells:])
    mix = case.params.mix
    muf, nuf = oc.mixture_face_visc(mix, False), oc.mixture_face_visc(mix, True)
    nif = m.nInternalFaces
    o, nb = m.owner[:nif], m.neighbour

    def face(v):
        return np.concatenate([0.5 * (v[o] + v[nb]), v[m.nCells:]])

    la = np.clip(face(a), 0.0, 1.0)
    mu_ref = la * mix.rho1 * face(n1) + (1 - la) * mix.rho2 * face(n2)
    assert np.allclose(muf, mu_ref, rtol=1e-13)
    assert np.allclose(nuf, mu_ref / (la * mix.rho1 + (1 - la) * mix.rho2), rtol=1e-13)
    assert O.field_min_max([oc], P.F_NU1) == (n1.min(), n1.max())
    assert np.allclose(oc.face_interpolate(P.F_NU1), face(n1), rtol=1e-14)  # fvc::interpolate, linear


def test_div_dev_rho_reff_explicit_analytic():
    """-div(mu*dev2(T(grad U))): zero for a linear velocity field, -(2/3 mu, 0, 0) for U = (x^2, 0, 0)"""
    n = 12
    m = M.hex_block(n, n, n, length=(1.2, 1.2, 1.2))
    xc = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    nt = m.nCells + m.nBoundaryFaces
    rho, nu = 2700.0, 3.0e-3
    for kind in ("linear", "quadratic"):
        oc = O.OracleCase(m)
        if kind == "linear":
            A = np.array([[0.3, -0.2, 0.1], [0.05, 0.4, -0.3], [0.2, 0.1, -0.7]])
            U = xc @ A
        else:
            U = np.stack([xc[:, 0] ** 2, 0 * xc[:, 0], 0 * xc[:, 0]], axis=1)
        oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
        oc.grad_u()
        oc.set(P.F_RHO_SOLVER, np.full(m.nCells, rho), np.full(m.nBoundaryFaces, rho))
        oc.set(P.F_NU, np.full(m.nCells, nu), np.full(m.nBoundaryFaces, nu))
        oc.div_dev_rho_reff_explicit()
        r = oc.internal(P.F_DIVDEVRHOREFF)
        i, j, k = np.unravel_index(np.arange(m.nCells), (n, n, n), order="F")
        inner = (i > 1) & (i < n - 2) & (j > 1) & (j < n - 2) & (k > 1) & (k < n - 2)
        scale = rho * nu
        if kind == "linear":
            assert np.abs(r).max() < 1e-9 * scale  # all cells: the corrected boundary gradient is exact too
        else:
            assert np.allclose(r[inner], [-(2.0 / 3.0) * scale, 0.0, 0.0], rtol=1e-9, atol=1e-9 * scale)
        # boundary values: the adjacent cell's value
        assert np.array_equal(oc.boundary(P.F_DIVDEVRHOREFF), r[m.owner[m.nInternalFaces:]])


# ---- alphaEqn.H:58-87: the interface-compression coefficient phic (OpenFOAM semantics restated, analytic pins) ----------
def test_alpha_compression_coefficient_analytic():
    """orthogonal hex, linear U = G x with a symmetric G: fvc::grad(U) is exact in interior cells, so on interior
    faces between interior cells  phic = cAlpha|phi|/|Sf| (1 - ic) + cAlpha ic |U|_f + sc |delta . G|  in closed form;
    non-coupled boundary faces are zero."""
    import numpy as np

    from oracle import pyoracle as O
    from solidstateamfoam_b200 import mesh as M
    from solidstateamfoam_b200 import params as P

    n = 6
    h = 0.1
    m = M.hex_block(n, n, n, length=(n * h, n * h, n * h))
    G = np.array([[0.3, 0.1, -0.2], [0.1, -0.5, 0.4], [-0.2, 0.4, 0.2]])  # symmetric, U_j = sum_i x_i G_ij
    X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
    U = X @ G
    oc = O.OracleCase(m)
    oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
    oc.grad_u()
    rng = np.random.default_rng(3)
    phi = rng.normal(size=m.nFaces) * h * h
    cA, ic, sc = 1.0, 0.25, 0.5
    phic = oc.alpha_compression_coeff(phi, cA, ic, sc)
    assert np.all(phic[m.nInternalFaces:] == 0.0)
    # interior cells: no boundary face
    touches = np.zeros(m.nCells, dtype=bool)
    touches[m.owner[m.nInternalFaces:]] = True
    own, nei = m.owner[: m.nInternalFaces], m.neighbour
    inner = ~touches[own] & ~touches[nei]
    assert inner.sum() > 50
    magSf = np.linalg.norm(m.Sf[: m.nInternalFaces], axis=1)
    magU = np.linalg.norm(U[: m.nCells], axis=1)
    magUf = 0.5 * (magU[own] + magU[nei])
    delta = m.C[nei] - m.C[own]
    expect = cA * np.abs(phi[: m.nInternalFaces]) / magSf * (1 - ic) + cA * ic * magUf + sc * np.linalg.norm(delta @ G, axis=1)
    assert np.allclose(phic[: m.nInternalFaces][inner], expect[inner], rtol=1e-12, atol=0)
    # the plain coefficient (icAlpha = scAlpha = 0) needs neither U nor gradU
    oc2 = O.OracleCase(m)
    p0 = oc2.alpha_compression_coeff(phi, 1.5, 0.0, 0.0)
    assert np.allclose(p0[: m.nInternalFaces], 1.5 * np.abs(phi[: m.nInternalFaces]) / magSf, rtol=1e-15)


def test_alpha_compression_coefficient_decomposed_equals_serial():
    """processor faces: interpolation with the neighbour rank's values and mesh.delta() from the exchanged cell centres
    reproduce the serial internal-face value (weights 0.5 on this uniform mesh: bit-identical interpolation)"""
    import numpy as np

    from oracle import pyoracle as O
    from solidstateamfoam_b200 import mesh as M
    from solidstateamfoam_b200 import params as P

    n = 6
    hblk = M.hex_block(n, n, n, _keep=True)
    cp = (np.arange(n * n * n) // (n * n) >= n // 2).astype(np.int32)  # two z-slabs
    g, parts = M.decompose(hblk, cp, 2)
    rng = np.random.default_rng(5)
    Xg = np.concatenate([g.C, g.Cf[g.nInternalFaces:]])
    Ug = np.stack([np.sin(3 * Xg[:, 0]) + Xg[:, 2], np.cos(2 * Xg[:, 1]), Xg[:, 0] * Xg[:, 2]], axis=1)
    phig = rng.normal(size=g.nFaces)
    og = O.OracleCase(g)
    og.set(P.F_U, Ug[: g.nCells], Ug[g.nCells:])
    og.grad_u()
    serial = og.alpha_compression_coeff(phig, 1.0, 0.3, 0.7)
    ocs = []
    for m in parts:
        X = np.concatenate([m.C, m.Cf[m.nInternalFaces:]])
        U = np.stack([np.sin(3 * X[:, 0]) + X[:, 2], np.cos(2 * X[:, 1]), X[:, 0] * X[:, 2]], axis=1)
        oc = O.OracleCase(m)
        oc.set(P.F_U, U[: m.nCells], U[m.nCells:])
        ocs.append(oc)
    O.halo_exchange(ocs, P.F_U)
    for oc in ocs:
        oc.grad_u()
    O.halo_exchange(ocs, P.F_GRADU)
    O.exchange_cell_centres(ocs)
    for m, oc in zip(parts, ocs):
        sign = np.where(m.faceFlip == 1, -1.0, 1.0)
        phic = oc.alpha_compression_coeff(sign * phig[m.faceGlobal], 1.0, 0.3, 0.7)
        for p in range(m.nPatches):
            sl = slice(int(m.patchStart[p]), int(m.patchStart[p] + m.patchSize[p]))
            if m.patchKind[p] == 1:  # processor faces = internal faces of the serial mesh
                assert np.allclose(phic[sl], serial[m.faceGlobal[sl]], rtol=1e-12, atol=1e-15)
                assert np.abs(phic[sl]).min() > 0
            else:
                assert np.all(phic[sl] == 0.0)
        nif = m.nInternalFaces
        assert np.allclose(phic[:nif], serial[m.faceGlobal[:nif]], rtol=1e-12, atol=1e-15)
