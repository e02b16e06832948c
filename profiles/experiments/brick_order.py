#!/usr/bin/env python
"""Experiment: how much does the tiled cell kernel gain when 256 consecutive cells form a compact brick
instead of one x-row of the blockMesh ordering?  The mesh is renumbered in numpy (as renumberMesh would,
faces back in upper-triangular order), the fields are permuted with it, and ssam_update_fused is timed on
both.  Usage: python profiles/experiments/brick_order.py [bx by bz]"""
import copy
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import params as P


def renumber(case, new_to_old):
    m = case.mesh
    nc, nif = m.nCells, m.nInternalFaces
    inv = np.empty(nc, dtype=np.int64)
    inv[new_to_old] = np.arange(nc)
    o = inv[m.owner[:nif]]
    n = inv[m.neighbour]
    flip = o > n
    lo, hi = np.where(flip, n, o), np.where(flip, o, n)
    order = np.lexsort((hi, lo))
    m2 = copy.copy(m)
    sign = np.where(flip, -1.0, 1.0)
    Sf = m.Sf.copy()
    Sf[:nif] *= sign[:, None]
    w = m.weights.copy()
    w[:nif] = np.where(flip, 1.0 - w[:nif], w[:nif])

    def perm_faces(a):
        a = a.copy()
        a[:nif] = a[:nif][order]
        return a

    m2.Sf, m2.weights = perm_faces(Sf), perm_faces(w)
    m2.Cf, m2.magSf, m2.deltaCoeffs = perm_faces(m.Cf), perm_faces(m.magSf), perm_faces(m.deltaCoeffs)
    own = m.owner.astype(np.int64).copy()
    own[:nif] = lo[order]
    own[nif:] = inv[m.owner[nif:]]
    m2.owner = own.astype(np.int32)
    m2.neighbour = hi[order].astype(np.int32)
    m2.V, m2.C = m.V[new_to_old], m.C[new_to_old]
    c2 = copy.copy(case)
    c2.mesh = m2
    for k in ("U_i", "T_i", "alpha1_i", "p_i"):
        setattr(c2, k, np.ascontiguousarray(getattr(case, k)[new_to_old]))
    return c2


def brick_perm(dims, b):
    nx, ny, nz = dims
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")  # cell = i + nx*(j + ny*k)
    key = (((k // b[2]) * (ny // b[1]) + (j // b[1])) * (nx // b[0]) + (i // b[0]))
    inner = ((k % b[2]) * b[1] + (j % b[1])) * b[0] + (i % b[0])
    full = key.astype(np.int64) * (b[0] * b[1] * b[2]) + inner
    return np.argsort(full.reshape(-1), kind="stable")


def time_case(ctx, s, case, steps=30):
    ctx.load_case(case)
    for _ in range(5):
        ctx.update_fused(case.params, 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(steps):
        ctx.update_fused(case.params, 0)
    e1.record(s)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    bricks = [tuple(int(v) for v in sys.argv[1:4])] if len(sys.argv) >= 4 else [(8, 8, 4), (16, 4, 4), (32, 4, 2), (64, 2, 2), (16, 16, 1)]
    dims = (256, 256, 128)
    case = cases.cfg2(dims=dims)
    ctx = capi.Context(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx.set_stream(s.cuda_stream)
    ctx.set_blocking(False)
    out = {"natural": time_case(ctx, s, case)}
    ref_i, _ = ctx.download(P.F_NU)
    for b in bricks:
        perm = brick_perm(dims, b)
        c2 = renumber(case, perm)
        out["brick%dx%dx%d" % b] = time_case(ctx, s, c2)
        got, _ = ctx.download(P.F_NU)
        err = np.abs(got - ref_i[perm]) / np.abs(ref_i[perm])
        out["brick%dx%dx%d_maxrel_vs_natural" % b] = float(err.max())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
