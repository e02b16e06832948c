#!/usr/bin/env python
"""Experiment: greedy clustering of 32-cell runs into 256-cell tiles (prototype of the tile-compact layout)."""
import heapq
import json
import os
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import scipy.sparse as sp


def cluster_perm(owner, neighbour, n_cells, run=32, per_tile=8):
    nb = (n_cells + run - 1) // run
    bo, bn = owner[: len(neighbour)] // run, neighbour // run
    keep = bo != bn
    A = sp.coo_matrix((np.ones(keep.sum(), dtype=np.int32), (bo[keep], bn[keep])), shape=(nb, nb)).tocsr()
    A = (A + A.T).tocsr()
    A.sum_duplicates()
    indptr, indices, data = A.indptr, A.indices, A.data
    assigned = np.zeros(nb, dtype=bool)
    order = []
    for seed in range(nb):
        if assigned[seed]:
            continue
        cluster = [seed]
        assigned[seed] = True
        w = {}
        age = {}

        def add_nbrs(b, a):
            for k in range(indptr[b], indptr[b + 1]):
                j = indices[k]
                if not assigned[j]:
                    w[j] = w.get(j, 0) + data[k]
                    age.setdefault(j, a)

        add_nbrs(seed, 0)
        while len(cluster) < per_tile and w:
            j = max(w, key=lambda q: (w[q], -age[q], -q))
            del w[j]
            assigned[j] = True
            cluster.append(j)
            add_nbrs(j, len(cluster))
        order.extend(sorted(cluster))
    order = np.asarray(order, dtype=np.int64)
    cells = (order[:, None] * run + np.arange(run)[None, :]).reshape(-1)
    return cells[cells < n_cells]


if __name__ == "__main__":
    import torch

    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "brick_order.py")).read()
    bo = types.ModuleType("bo")
    bo.__dict__["__file__"] = os.path.abspath(__file__)
    exec(compile(src.split("def main()")[0], "bo", "exec"), bo.__dict__)
    from solidstateamfoam_b200 import capi, cases

    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    case = {"cfg2": lambda: cases.cfg2(), "cfg3": lambda: cases.cfg3(base=(96, 96, 96)),
            "cfg4": lambda: cases.cfg4(dims=(200, 200, 200))}[which]()
    ctx = capi.Context(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx.set_stream(s.cuda_stream)
    ctx.set_blocking(False)
    out = {"case": which, "cells": case.mesh.nCells, "natural_ms": bo.time_case(ctx, s, case)}
    for run, per in ((32, 8), (64, 4), (16, 16)):
        t0 = time.time()
        perm = cluster_perm(case.mesh.owner.astype(np.int64), case.mesh.neighbour.astype(np.int64), case.mesh.nCells, run, per)
        out["cluster_%dx%d_build_s" % (run, per)] = round(time.time() - t0, 1)
        c2 = bo.renumber(case, perm)
        out["cluster_%dx%d_ms" % (run, per)] = bo.time_case(ctx, s, c2)
    print(json.dumps(out, indent=1))
