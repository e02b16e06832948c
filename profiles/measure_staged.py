#!/usr/bin/env python
"""Times the staged entry points and the widened ("next") rows on cfg2 with CUDA events and relates them to
their algorithmic bytes.  Run on a B200:  python profiles/measure_staged.py > gpurun_out/staged.json"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from solidstateamfoam_b200 import capi, cases
from solidstateamfoam_b200 import params as P


def main():
    case = cases.cfg2()
    m = case.mesh
    ctx = capi.Context(0)
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    ctx.set_stream(s.cuda_stream)
    ctx.load_case(case)
    ctx.update_fused(case.params, 0)
    n = m.nCells + m.nBoundaryFaces
    prgh = np.full(n, 1e5)
    ctx.upload(P.F_PRGH, prgh[: m.nCells], prgh[m.nCells:])
    ctx.set_blocking(False)
    p = case.params
    v = cases.sheppard_wright_aa6061()
    ss = P.StressStrainParams(Q=v.Q, R=v.R, sigmaR=v.sigmaR, A=v.A, n=v.n, deltaT=1e-3)
    nif, nb, nc = m.nInternalFaces, m.nBoundaryFaces, m.nCells
    grad_bytes = 40 * nif + 52 * nb + (8 + 24) * nc
    calls = {
        "ssam_strain_rate (fvc::grad + mag(symm), TEqn.H:3)": (lambda: ctx.strain_rate(math.sqrt(2 / 3), P.F_EPSDOT), grad_bytes + 8 * nc),
        "ssam_grad_u (fvc::grad(U), 9 planes out)": (lambda: ctx.grad_u(), grad_bytes + 72 * nc),
        "ssam_thermo_correct (FKS correct x2 + calcNu)": (lambda: ctx.thermo_correct(p.visc1, p.visc2, p.mix), (16 + 24 + 8 + 8 * 3 + 8) * n),
        "ssam_mixture_mu": (lambda: ctx.mixture("mu", p.mix), 32 * n),
        "ssam_mixture_k (cubic k(T))": (lambda: ctx.mixture("k", p.mix), 24 * n),
        "ssam_mixture_cp": (lambda: ctx.mixture("cp", p.mix), 16 * n),
        "ssam_friction_correct (qvisc + qfric)": (lambda: ctx.friction_correct(p.stick), 24 * n),
        "ssam_stress_strain (calculateStrain.H, incl. its fvc::grad)": (lambda: ctx.stress_strain(ss), grad_bytes + 72 * nc + (72 + 32 + 48) * n),
        "ssam_interface_correct (interfaceProperties::correct: gradAlpha, nHatf, K)": (
            lambda: ctx.interface_correct(1e-5), 92 * nif + 80 * nc + 60 * nb),
        "ssam_update_fused": (lambda: ctx.update_fused(p, 0), case.bytes_per_update),
    }
    peak = 6538.9
    out = {}
    for name, (fn, nbytes) in calls.items():
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(s)
        for _ in range(20):
            fn()
        e1.record(s)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[name] = {"ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3),
                     "GBps": round(nbytes / ms / 1e6, 1), "frac_of_measured_hbm": round(nbytes / ms / 1e6 / peak, 3)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
