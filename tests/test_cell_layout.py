"""Host logic of the tile-compact cell layout (no GPU): ssam_compact_cell_order is the clustering ssam_mesh_set
uses; the GPU-side equivalence of the two layouts is tests/test_gpu_parity.py::test_tile_compact_layout_is_bit_identical."""
import numpy as np
import pytest

from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import mesh as M

TILE, RUN = 256, 64


def remote_per_cell(mesh, order):
    inv = np.empty(mesh.nCells, dtype=np.int64)
    inv[order] = np.arange(mesh.nCells)
    o, n = inv[mesh.owner[: mesh.nInternalFaces]], inv[mesh.neighbour]
    return 2.0 * np.count_nonzero(o // TILE != n // TILE) / mesh.nCells


@pytest.mark.parametrize("dims", [(64, 32, 16), (256, 8, 8), (40, 24, 20), (33, 21, 17), (7, 5, 3)])
def test_order_is_a_run_preserving_permutation(dims):
    m = M.hex_block(*dims)
    order, before, after, use = capi.compact_cell_order(m)
    assert np.array_equal(np.sort(order), np.arange(m.nCells))
    # runs of RUN consecutive caller cells stay together and in order
    for s in range(0, m.nCells - m.nCells % RUN, RUN):
        run = order[s:s + RUN]
        if run[0] % RUN == 0 and run[0] + RUN <= m.nCells:
            assert np.array_equal(run, np.arange(run[0], run[0] + RUN))
    assert before == pytest.approx(remote_per_cell(m, np.arange(m.nCells)))
    assert after == pytest.approx(remote_per_cell(m, order))
    assert after <= before + 1e-12
    # deterministic
    assert np.array_equal(order, capi.compact_cell_order(m)[0])


def test_blockmesh_box_becomes_bricks():
    m = M.hex_block(256, 16, 16)  # one 256-cell tile = one x-row in the caller's order
    order, before, after, use = capi.compact_cell_order(m)
    assert use and before > 3.5 and after < 2.3
    # every full tile is a 64 x 2 x 2 brick
    i, j, k = order % 256, (order // 256) % 16, order // 4096
    for t in range(0, 64):
        sl = slice(t * TILE, (t + 1) * TILE)
        assert np.ptp(i[sl]) == 63 and np.ptp(j[sl]) == 1 and np.ptp(k[sl]) == 1


def test_auto_keeps_caller_order_when_it_does_not_pay():
    poly = cases.cfg3(base=(12, 12, 12)).mesh
    _, before, after, use = capi.compact_cell_order(poly)
    assert (before - after >= 1.0) == use
    tiny = M.hex_block(6, 6, 6)  # fewer than two tiles
    assert capi.compact_cell_order(tiny)[3] is False


def test_rejects_bad_labels():
    m = M.hex_block(8, 8, 8)
    m.neighbour = m.neighbour.copy()
    m.neighbour[5] = m.nCells + 3
    with pytest.raises(capi.SsamError):
        capi.compact_cell_order(m)
