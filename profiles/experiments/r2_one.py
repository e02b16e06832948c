#!/usr/bin/env python
"""One workload, one cell-kernel variant: `python profiles/experiments/r2_one.py cfg2 2 [steps]` prints one JSON line
(kernel ms from the library's CUDA events).  Environment switches (SSAM_PIPE_DYNAMIC, SSAM_PIPE_STAGGER_NS,
SSAM_PF_PIPE, SSAM_TILE_ORDER, SSAM_LAYOUT) are read by the library once per process, hence one process per setting.
Also the command the ncu captures of round 2 wrap."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from solidstateamfoam_b200 import capi
from profiles.experiments.r2_kernel_ab import PEAK, build


def main():
    name, variant = sys.argv[1], int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    case = build(name)
    m = case.mesh
    n_fric = int(m.patchSize[case.params.stick.fricPatch]) if case.params.stick.fricPatch >= 0 else 0
    nbytes = case.bytes_per_update - 40 * n_fric
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx = capi.Context(0)
    ctx.set_stream(s.cuda_stream)
    ctx.load_case(case)
    ctx.set_blocking(False)
    ctx.set_cell_kernel(variant)
    for _ in range(3):
        ctx.update_fused(case.params, 0)
    torch.cuda.synchronize()
    ctx.profile_get()
    ctx.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(steps):
        ctx.update_fused(case.params, 0)
    e1.record(s)
    torch.cuda.synchronize()
    ms, n = ctx.profile_get()
    k = ms / max(n, 1)
    env = {k_: v for k_, v in os.environ.items() if k_.startswith("SSAM_")}
    print(json.dumps({"workload": name, "cells": m.nCells, "variant": variant, "env": env, "kernel_ms": round(k, 4),
                      "step_ms": round(e0.elapsed_time(e1) / steps, 4),
                      "frac": round(nbytes / (k * 1e-3) / 1e9 / PEAK, 4), "compact": ctx.layout_is_compact()}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
