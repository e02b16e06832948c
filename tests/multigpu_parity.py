#!/usr/bin/env python
"""N-GPU parity check, one process per GPU (launch with torchrun / torch.distributed.run):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_parity.py [--decomp slab|random|block]

Every rank builds all partitions (small, deterministic), runs the multi-rank oracle (in-memory halo
copies) on the CPU, and compares its own GPU partition -- NCCL halos of U, temp, alpha1, then of
epsilonDotEq, and the 4-double all-reduce of the friction-patch sums -- against the oracle's partition:
integer halo maps bit-exact, grad U / strain rate bit-exact, the rest <= 1e-11.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_parts(n_ranks, decomp):
    from solidstateamfoam_b200 import cases

    return cases.parity_parts(n_ranks, decomp)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--decomp", default="slab", choices=["slab", "random", "block"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    from helpers import check_field, BIT_EXACT
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import capi
    from solidstateamfoam_b200 import params as P

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parts = build_parts(world, args.decomp)
    # ---- oracle, all ranks in this process ----
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
    # ---- GPU, this rank's partition ----
    case = parts[rank]
    ctx = capi.Context(local)
    uid = [capi.Context.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    ctx.comm_init(uid[0], rank, world)
    ctx.load_case(case)
    oc = ocs[rank]
    for which in (P.ADDR_HALO_SEND_CELLS, P.ADDR_HALO_PATCH_OFFSETS, P.ADDR_EXTRA_FACE, P.ADDR_EXTRA_OTHER):
        assert ctx.addressing(which).tobytes() == oc.addressing(which).tobytes(), P.ADDR_NAMES[which]
    worst = 0.0
    for variant in (0, 1, 2):
        ctx.set_cell_kernel(variant)
        ctx.update_fused(case.params, capi.FUSED_HALO_INPUTS | capi.FUSED_WRITE_GRAD)
        for f in case.outputs + [P.F_GRADU, P.F_U, P.F_T, P.F_ALPHA1]:
            gi, gb = ctx.download(f)
            w, _ = check_field(f"rank{rank} {P.FIELD_NAMES[f]}", np.concatenate([gi, gb]), oc.field(f), f in BIT_EXACT)
            worst = max(worst, w)
    c_gpu, a_gpu = ctx.friction_patch_centre()
    c_ref, a_ref = oc.friction_patch_centre()
    assert np.allclose(c_gpu, c_ref, rtol=1e-14, atol=0) and abs(a_gpu - a_ref) <= 1e-14 * a_ref
    # staged path across ranks too
    ctx.strain_rate(np.sqrt(2.0), P.F_STRAINRATE)
    ctx.grad_u()
    gi, gb = ctx.download(P.F_GRADU)
    check_field("staged gradU", np.concatenate([gi, gb]), oc.field(P.F_GRADU), True)
    # interfaceProperties::correct() across ranks: halos of alpha1, gradAlpha and K
    dn = ctx.interface_delta_n()
    assert abs(dn - O.interface_delta_n(ocs)) <= 1e-15 * dn
    dn = O.interface_delta_n(ocs)
    O.interface_correct_multi(ocs, dn)
    ctx.interface_correct(dn, capi.FUSED_HALO_INPUTS)
    for f in (P.F_GRADALPHA, P.F_CURVATURE):
        gi, gb = ctx.download(f)
        check_field(f"rank{rank} {P.FIELD_NAMES[f]}", np.concatenate([gi, gb]), oc.field(f), True)
    assert np.array_equal(ctx.nhatf(case.mesh.nInternalFaces, case.mesh.nBoundaryFaces), oc.nhatf())
    # explicit viscous-stress divergence across ranks (halos of gradU, rho, nu and of the result)
    ctx.update_fused(case.params, capi.FUSED_HALO_INPUTS | capi.FUSED_WRITE_GRAD)
    ctx.solver_rho_rhocp(case.params.mix)
    for o in ocs:
        o.solver_rho_rhocp(parts[0].params.mix)
    ctx.upload(P.F_NU, oc.internal(P.F_NU), oc.boundary(P.F_NU))
    for f in (P.F_GRADU, P.F_RHO_SOLVER, P.F_NU):
        O.halo_exchange(ocs, f)
    for o in ocs:
        o.div_dev_rho_reff_explicit()
    O.halo_exchange(ocs, P.F_DIVDEVRHOREFF)
    ctx.div_dev_rho_reff_explicit()
    gi, gb = ctx.download(P.F_DIVDEVRHOREFF)
    check_field(f"rank{rank} divDevRhoReff", np.concatenate([gi, gb]), oc.field(P.F_DIVDEVRHOREFF), True)
    # muf()/nuf() on processor faces, and the cross-rank "Min/max T"
    for f in (P.F_NU1, P.F_NU2):
        ctx.upload(f, oc.internal(f), oc.boundary(f))
    for nuf, ff in ((False, P.FF_MUF), (True, P.FF_NUF)):
        ctx.mixture_face_visc(case.params.mix, nuf)
        assert np.array_equal(ctx.face_field(ff, case.mesh.nInternalFaces, case.mesh.nBoundaryFaces),
                              oc.mixture_face_visc(case.params.mix, nuf))
    assert ctx.field_min_max(P.F_T) == O.field_min_max(ocs, P.F_T)
    n_proc = int(sum(case.mesh.patchSize[p] for p in range(case.mesh.nPatches) if case.mesh.patchKind[p] == 1))
    dist.barrier()
    print(f"[rank {rank}/{world}] {args.decomp}: {case.mesh.nCells} cells, {n_proc} processor faces, "
          f"halo transport {ctx.halo_transport()}, worst rel err {worst:.2e} -- OK", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
