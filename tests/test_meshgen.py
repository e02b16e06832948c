"""Invariants of the synthetic meshes (OpenFOAM conventions)."""
import numpy as np
import pytest

from solidstateamfoam_b200 import mesh as M

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)


def check_conventions(m):
    nif = m.nInternalFaces
    own, nei = m.owner[:nif], m.neighbour
    assert (own < nei).all()
    key = own.astype(np.int64) * m.nCells + nei
    assert (np.diff(key) > 0).all(), "internal faces must be in upper-triangular order"
    # patches tile the boundary faces in order
    assert m.patchStart[0] == nif if m.nPatches else True
    assert (m.patchStart[1:] == m.patchStart[:-1] + m.patchSize[:-1]).all()
    assert m.patchStart[-1] + m.patchSize[-1] == m.nFaces
    # closed cells: sum of outward area vectors vanishes
    acc = np.zeros((m.nCells, 3))
    np.add.at(acc, m.owner, m.Sf)
    np.subtract.at(acc, nei, m.Sf[:nif])
    assert np.abs(acc).max() <= 1e-12 * np.abs(m.Sf).max()
    # divergence theorem for the volume: V = 1/3 sum Sf . Cf (outward)
    v = np.zeros(m.nCells)
    d = np.einsum("fi,fi->f", m.Sf, m.Cf)
    np.add.at(v, m.owner, d)
    np.subtract.at(v, nei, d[:nif])
    assert np.allclose(v / 3.0, m.V, rtol=1e-10)
    assert (m.V > 0).all() and (m.weights > 0).all() and (m.weights <= 1).all()


def test_hex_block_counts_and_conventions():
    m = M.hex_block(5, 4, 3, length=(1.0, 0.8, 0.6), grading=(2.0, 1.0, 0.5))
    assert m.nCells == 60
    assert m.nInternalFaces == 4 * 4 * 3 + 5 * 3 * 3 + 5 * 4 * 2
    assert m.nBoundaryFaces == 2 * (4 * 3 + 5 * 3 + 5 * 4)
    assert m.patchName == ["xmin", "xmax", "ymin", "ymax", "zmin", "zmax"]
    check_conventions(m)
    # blockMesh order: cell c owns faces to c+1, c+nx, c+nx*ny in that order
    assert list(m.neighbour[:3]) == [1, 5, 20]


def test_sheared_hex_is_non_orthogonal():
    m = M.hex_block(4, 4, 4, affine=SHEAR)
    check_conventions(m)
    nif = m.nInternalFaces
    d = m.C[m.neighbour] - m.C[m.owner[:nif]]
    cosang = np.einsum("fi,fi->f", d, m.Sf[:nif]) / (np.linalg.norm(d, axis=1) * m.magSf[:nif])
    assert cosang.min() < 0.999


def test_poly_refined_face_counts():
    m = M.poly_refined(8, 8, 8, r1=0.4, d1=0.5, r2=0.2, d2=0.25)
    check_conventions(m)
    nf = np.bincount(m.owner, minlength=m.nCells) + np.bincount(m.neighbour, minlength=m.nCells)
    assert nf.min() == 6 and 6 < nf.max() <= 24
    # three cell sizes present
    vols = np.unique(np.round(m.V / m.V.min()).astype(int))
    assert set(vols) == {1, 8, 64}


def test_decompose_matches_global_and_pairs_up():
    h = M.hex_block(6, 5, 4, affine=SHEAR, _keep=True)
    nc = 6 * 5 * 4
    rng = np.random.default_rng(3)
    cell_proc = rng.integers(0, 3, size=nc)
    g, parts = M.decompose(h, cell_proc, 3)
    assert sum(p.nCells for p in parts) == g.nCells
    for r, p in enumerate(parts):
        check_conventions(p)
        # every physical patch is kept (possibly empty) in the same position on every rank
        assert p.patchName[:6] == g.patchName
        assert (p.patchKind[:6] == M.PATCH_GENERIC).all()
        # processor patches ordered by neighbour rank
        nb = p.patchNbrRank[6:]
        assert (np.diff(nb) > 0).all() and (p.patchKind[6:] == M.PATCH_PROCESSOR).all()
        assert np.array_equal(p.V, g.V[p.cellGlobal])
        for q_idx in range(6, p.nPatches):
            q = int(p.patchNbrRank[q_idx])
            other = parts[q]
            o_idx = [i for i in range(6, other.nPatches) if other.patchNbrRank[i] == r][0]
            assert p.patchSize[q_idx] == other.patchSize[o_idx]
            a = slice(p.patchStart[q_idx], p.patchStart[q_idx] + p.patchSize[q_idx])
            b = slice(other.patchStart[o_idx], other.patchStart[o_idx] + other.patchSize[o_idx])
            # same global faces in the same order, opposite orientation
            assert np.array_equal(p.faceGlobal[a], other.faceGlobal[b])
            assert np.array_equal(p.Sf[a], -other.Sf[b])
            assert np.array_equal(p.faceFlip[a] + other.faceFlip[b], np.ones(p.patchSize[q_idx], dtype=np.int32))
            # the two sides' weights are complementary up to rounding (makeWeights)
            assert np.allclose(p.weights[a] + other.weights[b], 1.0, rtol=0, atol=1e-14)


def test_slab_partition_equals_decompose_of_the_box():
    """The directly generated z-slabs (bench) equal decomposePar-style slabs of the global box."""
    nx, ny, nzr, R = 4, 3, 2, 3
    h = M.hex_block(nx, ny, nzr * R, length=(nx * 1.0, ny * 1.0, nzr * R * 1.0), _keep=True)
    cell_proc = np.repeat(np.arange(R), nx * ny * nzr)
    g, parts = M.decompose(h, cell_proc, R)
    for r in range(R):
        s = M.slab_partition(nx, ny, nzr, r, R)
        p = parts[r]
        assert s.patchName == p.patchName
        for a in ("owner", "neighbour", "patchStart", "patchSize", "patchKind", "patchNbrRank"):
            assert np.array_equal(getattr(s, a), getattr(p, a)), a
        for a in ("Sf", "Cf", "magSf", "V", "C", "weights", "deltaCoeffs"):
            assert np.allclose(getattr(s, a), getattr(p, a), rtol=1e-13, atol=1e-15), a
