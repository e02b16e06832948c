"""Shared harness: run a cases.Case through the oracle (CPU) and through the C-ABI (GPU), compare."""
from __future__ import annotations

import math

import numpy as np

from oracle import pyoracle as O
from solidstateamfoam_b200 import params as P

REL_TOL = 1e-11  # BASELINE.json north_star: every floating-point field within 1e-11 relative per cell
BIT_EXACT = (P.F_GRADU, P.F_EPSDOT, P.F_STRAINRATE, P.F_RHO, P.F_CP, P.F_NU2, P.F_RHO_SOLVER, P.F_RHOCP, P.F_U,
             P.F_GRADALPHA, P.F_CURVATURE)

EPS_FACTOR = math.sqrt(2.0 / 3.0)
SR_FACTOR = math.sqrt(2.0)


def oracle_case(case, apply_bc=True):
    oc = O.OracleCase(case.mesh, case.patch_u_bc)
    oc.set(P.F_U, case.U_i, case.U_b)
    oc.set(P.F_T, case.T_i, case.T_b)
    oc.set(P.F_ALPHA1, case.alpha1_i, case.alpha1_b)
    oc.set(P.F_P, case.p_i, case.p_b)
    if apply_bc and case.rotslip is not None:
        oc.bc_rotational_slip(case.rotslip)
    return oc


def oracle_update(case, form=O.FAITHFUL, ctx=None):
    """Full update on the oracle.  With `ctx`, U is taken from the GPU context: the exponential
    rotational-slip BC contains an exp(), so its U_b agrees with the oracle's to ~1 ulp rather than
    bit for bit, and the bit-exact gradient comparison needs identical inputs on both sides (the BC
    itself is compared separately, test_rotational_slip_bcs)."""
    oc = oracle_case(case)
    if ctx is not None:
        gi, gb = ctx.download(P.F_U)
        check_field("U (BC output)", np.concatenate([gi, gb]), oc.field(P.F_U), False)
        oc.set(P.F_U, gi, gb)
    oc.update(case.params, form=form)
    return oc


def ulp_diff(a, b):
    """distance in units in the last place between two float64 arrays (same sign assumed or tiny)."""
    ai = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    bi = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    ai = np.where(ai < 0, np.int64(-(2 ** 63)) - ai, ai)
    bi = np.where(bi < 0, np.int64(-(2 ** 63)) - bi, bi)
    return np.abs(ai - bi)


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    both_nan = np.isnan(got) & np.isnan(ref)
    d = np.abs(got - ref)
    denom = np.abs(ref)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(d == 0, 0.0, d / denom)
    r = np.where(both_nan, 0.0, r)
    r = np.where(np.isnan(r), np.inf, r)
    return r


def check_field(name, got, ref, bit_exact):
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    if bit_exact:
        same = (got == ref) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), f"{name}: {np.count_nonzero(~same)} of {same.size} values differ (must be bit-exact); " \
                           f"max rel {rel_err(got, ref).max():.3e}"
        return 0.0, 0
    r = rel_err(got, ref)
    worst = float(r.max()) if r.size else 0.0
    assert worst <= REL_TOL, f"{name}: max relative error {worst:.3e} > {REL_TOL:g} at index {int(np.argmax(r))}"
    u = int(ulp_diff(got, ref).max()) if r.size else 0
    return worst, u


def compare_ctx_oracle(ctx, oc, fields, report=None):
    """download each field from the GPU context and compare internal+boundary with the oracle."""
    for f in fields:
        gi, gb = ctx.download(f)
        got = np.concatenate([gi, gb], axis=0)
        ref = oc.field(f)
        worst, u = check_field(P.FIELD_NAMES[f], got, ref, f in BIT_EXACT)
        if report is not None:
            report[P.FIELD_NAMES[f]] = (worst, u)
