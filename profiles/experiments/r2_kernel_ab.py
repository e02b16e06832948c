#!/usr/bin/env python
"""Round-2 A/B of the fused cell kernel on one B200: variant 1 (one CTA per tile) against variant 2 (persistent,
self-pipelined), seed against Morton tile order, with and without the L2 prefetch, on the BASELINE workloads.

    python profiles/experiments/r2_kernel_ab.py [workloads...] > gpurun_out/r2_kernel_ab.jsonl

One JSON line per (workload, tile order, variant, prefetch): kernel ms from the library's own CUDA events around the
cell-kernel launch (ssam_profile_*), 20 timed updates after 4 warm-ups, inputs >> L2."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from solidstateamfoam_b200 import capi, cases

PEAK = 6538.9
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass


def build(name):
    if name == "cfg2":
        return cases.cfg2()
    if name == "cfg2gen":  # general exponents: no pow(x,1) shortcuts
        return cases.cfg2(visc1=cases.fks_general_exponents())
    if name == "cfg4":
        return cases.cfg4()
    if name == "cfg3":
        return cases.cfg3()
    if name == "cfg5":
        return cases.cfg5_slab(0, 1, (512, 512, 256))
    if name == "cfg5q":  # a quarter of the box (z): the same 512x512 planes at a quarter of the set-up time
        return cases.cfg5_slab(0, 1, (512, 512, 64))
    if name == "cfg2sw":  # cfg2's mesh with cfg5's constitutive model: separates model cost from mesh shape
        c = cases.cfg2()
        c.params.visc1 = cases.sheppard_wright_aa6061()
        return c
    if name == "cfg5qfks":  # cfg5's planes with cfg2's model
        c = cases.cfg5_slab(0, 1, (512, 512, 64))
        c.params.visc1 = cases.fks_placeholder()
        return c
    if name == "cfg5qnh":  # ... and with NortonHoff (needs no sigmay -> no friction patch)
        c = cases.cfg5_slab(0, 1, (512, 512, 64))
        c.params.visc1 = cases.norton_hoff_placeholder()
        c.params.stick.fricPatch = -1
        c.bytes_per_update = cases.algorithmic_bytes(c.mesh, 0, True)
        return c
    if name == "cfg5y":  # the same 16.8M cells as 256 x 512 x 128: is it the plane size?
        return cases.cfg5_part(0, 1, "block", dims=(256, 512, 128))
    raise SystemExit(name)


def main():
    names = sys.argv[1:] or ["cfg2", "cfg5q", "cfg3", "cfg4"]
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    for name in names:
        t0 = time.time()
        case = build(name)
        m = case.mesh
        n_fric = int(m.patchSize[case.params.stick.fricPatch]) if case.params.stick.fricPatch >= 0 else 0
        nbytes = case.bytes_per_update - 40 * n_fric
        print(f"# {name}: {m.nCells} cells built in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
        for order in (1, 0):
            os.environ["SSAM_TILE_ORDER"] = str(order)
            t0 = time.time()
            ctx = capi.Context(0)
            ctx.set_stream(s.cuda_stream)
            ctx.load_case(case)
            ctx.set_blocking(False)
            setup = time.time() - t0
            for variant in (2, 1):
                for pf in ("default", "0"):
                    if pf == "default":
                        os.environ.pop("SSAM_PF", None)
                    else:
                        os.environ["SSAM_PF"] = pf
                    ctx.set_cell_kernel(variant)
                    for _ in range(4):
                        ctx.update_fused(case.params, 0)
                    torch.cuda.synchronize()
                    ctx.profile_get()
                    ctx.profile_enable(True)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(s)
                    for _ in range(20):
                        ctx.update_fused(case.params, 0)
                    e1.record(s)
                    torch.cuda.synchronize()
                    ms, n = ctx.profile_get()
                    ctx.profile_enable(False)
                    k = ms / max(n, 1)
                    print(json.dumps({"workload": name, "cells": m.nCells, "compact": ctx.layout_is_compact(),
                                      "tile_order": order, "variant": variant, "pf": pf, "kernel_ms": round(k, 4),
                                      "step_ms": round(e0.elapsed_time(e1) / 20, 4),
                                      "frac": round(nbytes / (k * 1e-3) / 1e9 / PEAK, 4),
                                      "mcells_s": round(m.nCells / (e0.elapsed_time(e1) / 20 * 1e-3) / 1e6, 1),
                                      "mesh_set_s": round(setup, 1)}), flush=True)
            os.environ.pop("SSAM_PF", None)
            compact = ctx.layout_is_compact()
            ctx.close()
            if not compact:
                break  # caller-order layout: the tile order switch changes nothing


if __name__ == "__main__":
    main()
