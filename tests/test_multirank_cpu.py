"""N>1 host-side logic on the CPU: decomposition conventions, halo maps and the cross-rank patch-centre
sum, exercised (a) in-process by the multi-rank oracle and (b) by two real processes over gloo."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as O
from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from multigpu_parity import build_parts  # noqa: E402


def oracle_parts(parts):
    ocs = []
    for c in parts:
        oc = O.OracleCase(c.mesh, c.patch_u_bc)
        oc.set(P.F_U, c.U_i, c.U_b)
        oc.set(P.F_T, c.T_i, c.T_b)
        oc.set(P.F_ALPHA1, c.alpha1_i, c.alpha1_b)
        oc.set(P.F_P, c.p_i, c.p_b)
        oc.bc_rotational_slip(c.rotslip)
        ocs.append(oc)
    return ocs


@pytest.mark.parametrize("decomp", ["slab", "random"])
def test_decomposed_oracle_forms_agree(decomp):
    parts = build_parts(3, decomp)
    res = []
    for form in (O.FAITHFUL, O.FUSED):
        ocs = oracle_parts(parts)
        O.halo_exchange(ocs, P.F_T)
        O.halo_exchange(ocs, P.F_ALPHA1)
        O.update_multi(ocs, parts[0].params, form=form)
        res.append(ocs)
    for a, b, c in zip(res[0], res[1], parts):
        for f in c.outputs + [P.F_GRADU]:
            assert np.array_equal(a.field(f), b.field(f), equal_nan=True), P.FIELD_NAMES[f]


def test_decomposed_matches_serial_away_from_processor_faces():
    """Serial and decomposed runs use different (but equivalent) interpolation formulas on processor faces
    (w*U_P + (1-w)*U_N with each side's own w), so they agree to rounding, not bitwise; cells
    that touch no processor face must agree bit for bit."""
    n_ranks = 2
    cell = 0.5e-3
    dims = (10, 8, 4 * n_ranks)
    box = ((0.0, 0.0, 0.0), (dims[0] * cell, dims[1] * cell, dims[2] * cell))
    h = M.hex_block(*dims, length=box[1], _keep=True)
    cp = np.repeat(np.arange(n_ranks), dims[0] * dims[1] * 4)
    g, meshes = M.decompose(h, cp, n_ranks)
    mix = P.mixture(2700.0, 1.2, 896.0, 1005.0, P.k_constant(167.0), P.k_constant(0.026))

    def mk(m, seed):
        top = m.patch_id("zmax")
        return cases._finish("x", m, cases._fields(m, seed, "tanh", box=box), cases.sheppard_wright_aa6061(), mix,
                             P.constant_slip_stick(top, 0.9, 0.5, cases.OMEGA, 0.3),
                             P.constant_rotational_slip(top, cases.OMEGA, 0.5, cases.VT), seed)

    serial = mk(g, 5)
    parts = [mk(m, 5) for m in meshes]
    # identical inputs: restrict the serial fields to the partitions
    for c, m in zip(parts, meshes):
        c.U_i, c.T_i, c.alpha1_i, c.p_i = (serial.U_i[m.cellGlobal], serial.T_i[m.cellGlobal],
                                           serial.alpha1_i[m.cellGlobal], serial.p_i[m.cellGlobal])
        nb = m.nBoundaryFaces
        fg = m.faceGlobal[m.nInternalFaces:]
        phys = fg >= g.nInternalFaces
        for name in ("U_b", "T_b", "alpha1_b", "p_b"):
            dst = getattr(c, name)
            src = getattr(serial, name)
            dst[phys] = src[fg[phys] - g.nInternalFaces]
        assert nb == dst.shape[0]
    so = oracle_parts([serial])[0]
    so.update(serial.params)
    ocs = oracle_parts(parts)
    O.halo_exchange(ocs, P.F_T)
    O.halo_exchange(ocs, P.F_ALPHA1)
    O.update_multi(ocs, parts[0].params)
    for oc, m in zip(ocs, meshes):
        touches = np.zeros(m.nCells, dtype=bool)
        for p in range(m.nPatches):
            if m.patchKind[p] == M.PATCH_PROCESSOR:
                touches[m.face_cells(p)] = True
        for f in (P.F_EPSDOT, P.F_NU1, P.F_MUEFF, P.F_QVISC, P.F_K):
            a = oc.internal(f)
            b = so.internal(f)[m.cellGlobal]
            assert np.array_equal(a[~touches], b[~touches]), P.FIELD_NAMES[f]
            assert np.allclose(a[touches], b[touches], rtol=1e-9, atol=0), P.FIELD_NAMES[f]
        # the friction patch centre is a cross-rank sum: equal up to summation order
        c1, a1 = oc.friction_patch_centre()
        c0, a0 = so.friction_patch_centre()
        assert np.allclose(c1, c0, rtol=1e-13) and abs(a1 - a0) <= 1e-13 * a0


GLOO_WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
ROOT = os.environ["SSAM_ROOT"]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from multigpu_parity import build_parts
from oracle import pyoracle as O
from solidstateamfoam_b200 import params as P
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
parts = build_parts(world, os.environ["SSAM_DECOMP"])
case = parts[rank]
m = case.mesh
oc = O.OracleCase(m, case.patch_u_bc)
for f, i, b in ((P.F_U, case.U_i, case.U_b), (P.F_T, case.T_i, case.T_b), (P.F_ALPHA1, case.alpha1_i, case.alpha1_b), (P.F_P, case.p_i, case.p_b)):
    oc.set(f, i, b)
oc.bc_rotational_slip(case.rotslip)

def halo(field):
    """processorFvPatchField::evaluate over gloo: send field[faceCells], receive into the patch slice"""
    a = oc.field(field)
    a2 = a.reshape(a.shape[0], -1)
    reqs, recv = [], []
    for p in range(m.nPatches):
        if m.patchKind[p] != P.PATCH_PROCESSOR: continue
        peer = int(m.patchNbrRank[p])
        send = torch.from_numpy(np.ascontiguousarray(a2[m.face_cells(p)]))
        buf = torch.empty_like(send)
        reqs.append(dist.isend(send, peer)); reqs.append(dist.irecv(buf, peer)); recv.append((p, buf))
    for r in reqs: r.wait()
    for p, buf in recv:
        sl = m.patch_slice_b(p)
        a2[m.nCells + sl.start: m.nCells + sl.stop] = buf.numpy()

for f in (P.F_U, P.F_T, P.F_ALPHA1): halo(f)
oc.grad_u(O.FAITHFUL)
halo(P.F_GRADU)
oc.strain_from_grad(np.sqrt(2.0 / 3.0), P.F_EPSDOT)
sums = torch.from_numpy(oc.friction_patch_sums(case.params.stick.fricPatch))
parts_sums = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
dist.all_gather(parts_sums, sums)
tot = parts_sums[0].clone()
for s in parts_sums[1:]: tot = tot + s          # rank order, like the in-process oracle
p = case.params
oc.viscosity_correct(1, p.visc1); oc.viscosity_correct(2, p.visc2)
oc.mixture("calc_nu", p.mix); oc.mixture("mu", p.mix)
oc.friction_correct(p.stick, sums=tot.numpy())
for w in ("cp", "k", "rho"): oc.mixture(w, p.mix)
np.savez(os.path.join(os.environ["SSAM_OUT"], f"rank{rank}.npz"), **{P.FIELD_NAMES[f]: oc.field(f) for f in case.outputs + [P.F_GRADU]})
dist.destroy_process_group()
'''


@pytest.mark.parametrize("decomp", ["slab", "random", "block"])
def test_two_process_gloo_halos_equal_in_process_oracle(tmp_path, decomp):
    worker = tmp_path / "worker.py"
    worker.write_text(GLOO_WORKER)
    env = dict(os.environ, SSAM_ROOT=ROOT, SSAM_DECOMP=decomp, SSAM_OUT=str(tmp_path), OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533" if decomp == "slab" else "29534", str(worker)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    parts = build_parts(2, decomp)
    ocs = oracle_parts(parts)
    O.halo_exchange(ocs, P.F_T)
    O.halo_exchange(ocs, P.F_ALPHA1)
    O.update_multi(ocs, parts[0].params)
    for rank in range(2):
        got = np.load(tmp_path / f"rank{rank}.npz")
        for f in parts[rank].outputs + [P.F_GRADU]:
            assert np.array_equal(got[P.FIELD_NAMES[f]], ocs[rank].field(f), equal_nan=True), (rank, P.FIELD_NAMES[f])
