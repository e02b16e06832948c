#!/usr/bin/env python
"""Freezes outputs of the reference's own model sources (oracle/_ref, see oracle/Makefile `ref`) into
tests/golden/reference_sources.npz, so that tests/test_reference_golden.py can check the oracle against the
reference on machines where /root/reference does not exist (the GPU box).

Run where the reference tree is mounted:   python tests/golden/gen_reference_golden.py
Every array is the reference's result for a deterministic synthetic case (solidstateamfoam_b200.cases, Philox
seeds); `inputs_sha256` guards against the case generators drifting away from what was frozen."""
import hashlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np

from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import params as P

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)
GOLDEN_CASES = {
    "cfg1_sheppardwright_constantslip": lambda: cases.cfg1(dims=(7, 6, 5), affine=SHEAR, grading=(2.0, 0.5, 3.0)),
    "cfg2_fks_cubic_kelvin": lambda: cases.cfg2(dims=(6, 6, 5)),
    "cfg2_fks_cubic_celsius": lambda: cases.cfg2(dims=(5, 5, 4), units="Celsius"),
    "cfg3_poly_exponentialslip": lambda: cases.cfg3(base=(4, 4, 4)),
    "cfg4_nortonhoff": lambda: cases.cfg4(dims=(5, 5, 4)),
    # other corners of the parameter space: FKS with b2, c2 != 1 (no pow shortcuts), a different cubic k(T);
    # SheppardWright with other n, A and default limiters far away; NortonHoff with another exponent
    "cfg2_fks_general_exponents": lambda: cases.cfg2(
        dims=(6, 5, 4), k_coeffs=(35.0, -0.05, 3e-4, -2e-7),
        visc1=P.fks(a1=2.1e8, a2=0.6e8, b1=0.05, b2=0.7, b3=1.3, e0=0.01, c1=3.0, c2=2.5, Tm=900.0, Ta=300.0, rho=2650.0,
                    nuMin=1e-8, nuMax=1e6, eDotMin=1e-9)),
    "cfg1_sheppardwright_other_alloy": lambda: cases.cfg1(
        dims=(6, 5, 4), visc1=P.sheppard_wright(sigmaR=3.1e7, Q=1.9e5, R=8.314, A=5.0e10, n=5.2, rho=2800.0, nuMin=1e-9,
                                              nuMax=1e7, eDotMin=1e-10)),
    "cfg4_nortonhoff_other_exponent": lambda: cases.cfg4(
        dims=(5, 4, 4), visc1=P.norton_hoff(mu0=3e7, m=0.35, rho=2700.0, nuMin=1e-7, nuMax=1e5, strainRateEffMin=1e-8)),
}


def inputs_digest(case) -> str:
    h = hashlib.sha256()
    m = case.mesh
    for a in (m.owner, m.neighbour, m.Sf, m.weights, m.V, m.C, case.U_i, case.U_b, case.T_i, case.T_b, case.alpha1_i,
              case.alpha1_b, case.p_i, case.p_b):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def main():
    import tempfile

    from oracle import pyref as R
    from ref_sequence import run_reference

    out = {}
    for name, mk in GOLDEN_CASES.items():
        case = mk()
        d = tempfile.mkdtemp(prefix="ssam_golden_")
        ref, _ = run_reference(case, d)
        if case.rotslip is not None:  # boundaryConditions/*RotationalSlip*::updateCoeffs()
            rc = R.ReferenceCase(case.mesh, d)
            ref["bc_Ub"] = rc.bc_rotational_slip("zmax").reshape(-1)
            rc.close()
        out[name + "/inputs_sha256"] = np.frombuffer(inputs_digest(case).encode(), dtype=np.uint8)
        for k, v in ref.items():
            out[name + "/" + k] = v
        print(name, case.mesh.nCells, "cells,", len(ref), "fields")
    path = os.path.join(HERE, "reference_sources.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
