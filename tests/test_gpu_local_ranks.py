"""Multi-rank parity on ONE GPU (VERDICT round 1 "driver-side multi-GPU parity"): the ranks are
contexts of this process linked by ssam_comm_init_local.  They run the same peer-memory schedule as one-process-per-GPU
runs (push into the neighbour's receive slot, flag, stream-ordered wait, unpack) -- no kernel ever spins on another
rank, so sharing a GPU is safe -- and are compared with the multi-rank CPU oracle: halo maps bit-exact, grad U /
strain rate bit-exact, everything else <= 1e-11."""
import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import params as P

from helpers import BIT_EXACT, check_field

pytestmark = pytest.mark.gpu


def oracle_ranks(parts):
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
    O.update_multi(ocs, parts[0].params, form=O.FAITHFUL)
    return ocs


@pytest.mark.parametrize("n_ranks,decomp", [(2, "slab"), (2, "random"), (3, "random"), (4, "random")])
def test_in_process_rank_group_matches_the_decomposed_oracle(n_ranks, decomp):
    parts = cases.parity_parts(n_ranks, decomp)
    ocs = oracle_ranks(parts)
    ctxs = [capi.Context(0) for _ in parts]
    try:
        capi.Context.comm_init_local(ctxs)
        flags = capi.FUSED_HALO_INPUTS | capi.FUSED_WRITE_GRAD
        phis = [np.random.default_rng(23 + r).normal(size=c.mesh.nFaces) * 1e-6 for r, c in enumerate(parts)]
        for ctx, case, oc, phi in zip(ctxs, parts, ocs, phis):
            ctx.load_case(case)
            for which in (P.ADDR_HALO_SEND_CELLS, P.ADDR_HALO_PATCH_OFFSETS):
                assert ctx.addressing(which).tobytes() == oc.addressing(which).tobytes(), P.ADDR_NAMES[which]
            ctx.face_field_upload(P.FF_PHI, phi)  # every host-blocking transfer before the first collective call
        for ctx, case in zip(ctxs, parts):
            for variant in (2, 1):
                ctx.set_cell_kernel(variant)
                ctx.prepare_fused(case.params, flags)  # all allocations / lazy set-up, nothing enqueued
            ctx.set_blocking(False)
        worst = 0.0
        for variant in (1, 2):
            for step in range(2):  # the second update reuses the double-buffered slots and the cached centroid
                for ctx, case in zip(ctxs, parts):  # enqueue EVERY rank before synchronising any
                    ctx.set_cell_kernel(variant)
                    ctx.update_fused(case.params, flags)
                for ctx in ctxs:
                    ctx.sync()
            for r, (ctx, case, oc) in enumerate(zip(ctxs, parts, ocs)):
                assert ctx.halo_transport() == "peer-memory"
                for f in case.outputs + [P.F_GRADU, P.F_U, P.F_T, P.F_ALPHA1]:
                    gi, gb = ctx.download(f)
                    w, _ = check_field(f"rank{r} {P.FIELD_NAMES[f]} (variant {variant})", np.concatenate([gi, gb]),
                                       oc.field(f), f in BIT_EXACT)
                    worst = max(worst, w)
                c_gpu, a_gpu = ctx.friction_patch_centre()
                c_ref, a_ref = oc.friction_patch_centre()
                assert np.allclose(c_gpu, c_ref, rtol=1e-14, atol=0) and abs(a_gpu - a_ref) <= 1e-14 * a_ref
        # staged gradient across the ranks (one exchange of the nine GRADU planes)
        for ctx in ctxs:
            ctx.grad_u()
        for ctx in ctxs:
            ctx.sync()
        for r, (ctx, oc) in enumerate(zip(ctxs, ocs)):
            gi, gb = ctx.download(P.F_GRADU)
            check_field(f"rank{r} staged gradU", np.concatenate([gi, gb]), oc.field(P.F_GRADU), True)
        # alphaEqn.H:58-87 on processor faces: interpolation with the neighbour rank's U / gradU, mesh.delta() from the
        # neighbour cell centres exchanged once per mesh
        O.halo_exchange(ocs, P.F_U)
        O.halo_exchange(ocs, P.F_GRADU)
        O.exchange_cell_centres(ocs)
        for ctx in ctxs:
            ctx.alpha_compression_coeff(0.9, 0.2, 0.5)
        for ctx in ctxs:
            ctx.sync()
        for r, (ctx, case, oc, phi) in enumerate(zip(ctxs, parts, ocs, phis)):
            m = case.mesh
            got = ctx.face_field(P.FF_PHIC, m.nInternalFaces, m.nBoundaryFaces)
            ref = oc.alpha_compression_coeff(phi, 0.9, 0.2, 0.5)
            assert np.array_equal(got, ref), f"rank {r}: {np.count_nonzero(got != ref)} faces of phic differ"
            proc = np.concatenate([np.arange(m.patchStart[p], m.patchStart[p] + m.patchSize[p])
                                   for p in range(m.nPatches) if m.patchKind[p] == 1])
            assert np.abs(got[proc]).min() > 0
        print(f"{n_ranks} in-process ranks ({decomp}): worst rel err {worst:.2e}, "
              f"{[c.halo_bytes_per_update() for c in ctxs]} halo bytes per update")
    finally:
        for ctx in ctxs:
            ctx.close()


def test_blocking_collective_on_an_in_process_group_is_refused():
    parts = cases.parity_parts(2, "slab")
    ctxs = [capi.Context(0) for _ in parts]
    try:
        capi.Context.comm_init_local(ctxs)
        for ctx, case in zip(ctxs, parts):
            ctx.load_case(case)
        with pytest.raises(capi.SsamError) as e:
            ctxs[0].update_fused(parts[0].params, capi.FUSED_HALO_INPUTS)
        assert "ssam_set_blocking" in str(e.value)
    finally:
        for ctx in ctxs:
            ctx.close()
