"""The structured block partitions bench.py times at N > 1 (cases.block_grid / cases.cfg5_part) are generated
directly, rank by rank, without a global mesh.  Here they are checked against the decomposePar stand-in applied to the
global box (processor patches after the physical ones, ordered by neighbour rank, both sides listing their
faces in the same order): identical integer tables, geometry equal to rounding (the blocks are generated from their own
origin), and identical results of the multi-rank CPU oracle."""
import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

INT_TABLES = ("owner", "neighbour", "patchStart", "patchSize", "patchKind", "patchNbrRank")


def decomposed_reference(n_ranks, dims, cell):
    grid = cases.block_grid(n_ranks, dims)
    loc = [dims[a] // grid[a] for a in range(3)]
    h = M.hex_block(*dims, length=tuple(d * cell for d in dims), _keep=True)
    ii, jj, kk = np.meshgrid(np.arange(dims[0]), np.arange(dims[1]), np.arange(dims[2]), indexing="ij")
    r = (ii // loc[0]) + grid[0] * ((jj // loc[1]) + grid[1] * (kk // loc[2]))
    _, meshes = M.decompose(h, r.transpose(2, 1, 0).reshape(-1), n_ranks)
    return grid, meshes


def test_block_grid_of_the_headline_box():
    dims = (512, 512, 256)
    assert cases.block_grid(1, dims) == (1, 1, 1)
    assert cases.block_grid(2, dims) == (1, 2, 1)
    assert cases.block_grid(4, dims) == (2, 2, 1)
    assert cases.block_grid(8, dims) == (2, 2, 2)
    with pytest.raises(ValueError):
        cases.block_grid(7, dims)


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_direct_blocks_equal_the_decomposed_global_mesh(n_ranks):
    dims, cell = (8, 6, 4), 0.25e-3
    grid, ref = decomposed_reference(n_ranks, dims, cell)
    for rank in range(n_ranks):
        a, b = cases.cfg5_part(rank, n_ranks, "block", dims=dims).mesh, ref[rank]
        assert a.nCells == b.nCells and a.nInternalFaces == b.nInternalFaces
        for k in INT_TABLES:
            assert np.array_equal(getattr(a, k), getattr(b, k)), (rank, k)
        for k, x in vars(a).items():
            if isinstance(x, np.ndarray) and x.dtype.kind == "f":
                y = getattr(b, k)
                assert x.shape == y.shape, (rank, k)
                scale = max(float(np.abs(y).max()), 1e-300)
                assert float(np.abs(x - y).max()) <= 1e-12 * scale, (rank, k)
    # at 8 ranks every block has exactly three neighbours
    if n_ranks == 8:
        m = cases.cfg5_part(5, 8, "block", dims=dims).mesh
        assert sorted(int(r) for r in m.patchNbrRank if r >= 0) == [1, 4, 7]


@pytest.mark.parametrize("n_ranks", [2, 3, 4, 8])
def test_parity_parts_block_runs_through_the_multirank_oracle(n_ranks):
    """the small block case of the parity checks: both oracle forms agree bit for bit on every rank, halos included"""
    parts = cases.parity_parts(n_ranks, "block")
    assert sum(c.mesh.nCells for c in parts) == 12 ** 3
    res = []
    for form in (O.FAITHFUL, O.FUSED):
        ocs = []
        for c in parts:
            oc = O.OracleCase(c.mesh, c.patch_u_bc)
            oc.set(P.F_U, c.U_i, c.U_b)
            oc.set(P.F_T, c.T_i, c.T_b)
            oc.set(P.F_ALPHA1, c.alpha1_i, c.alpha1_b)
            oc.set(P.F_P, c.p_i, c.p_b)
            oc.bc_rotational_slip(c.rotslip)
            ocs.append(oc)
        O.halo_exchange(ocs, P.F_T)
        O.halo_exchange(ocs, P.F_ALPHA1)
        O.update_multi(ocs, parts[0].params, form=form)
        res.append(ocs)
    for a, b, c in zip(res[0], res[1], parts):
        for f in c.outputs + [P.F_GRADU]:
            assert np.array_equal(a.field(f), b.field(f), equal_nan=True), P.FIELD_NAMES[f]
        assert np.isfinite(a.field(P.F_EPSDOT)[0]).all()


@pytest.mark.parametrize("decomp,n_ranks", [("block", 2), ("block", 4), ("block", 8), ("slab", 2), ("slab", 4), ("slab", 8)])
def test_every_rank_names_the_friction_patch(decomp, n_ranks):
    """constantSlipStick's reduce() of the patch centroid (constantSlipStick.C:133-135) is a collective of ALL ranks:
    a partition without tool faces must still pass the (empty) patch, or the ranks that own the faces wait for ever in
    the all-reduce.  (Round 2's first 8-GPU run stalled on exactly this: the bottom blocks passed fricPatch = -1.)"""
    dims = (8, 8, 16)
    parts = [cases.cfg5_part(r, n_ranks, decomp, dims=dims) for r in range(n_ranks)]
    fric = {int(c.params.stick.fricPatch) for c in parts}
    assert len(fric) == 1 and fric.pop() >= 0
    assert {int(c.params.stick.model) for c in parts} == {P.STICK_CONSTANT}
    sizes = [int(c.mesh.patchSize[c.params.stick.fricPatch]) for c in parts]
    assert sum(sizes) == dims[0] * dims[1]
    if n_ranks == 8:
        assert 0 in sizes  # some ranks really hold no tool face
    # rotational-slip BC likewise present on every rank (applies to zero faces where the patch is empty)
    assert all(c.rotslip is not None for c in parts)
