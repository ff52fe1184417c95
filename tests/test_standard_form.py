"""standard_form_qp (/root/reference/scripts/common.jl:129-288, the second half of SURVEY 8f N4): the host-side product
version (madqp.jl_b200/preprocess.py, vectorised) against the oracle's loop-for-loop restatement, entry for entry, and
the equivalence of the reformulated problem."""
import dataclasses

import numpy as np
import pytest

from madqp_jl_b200 import preprocess
from madqp_jl_b200 import problems as P
from oracle import oracle as O

FIELDS = ("nvar", "ncon", "c", "c0", "H_rows", "H_cols", "H_vals", "A_rows", "A_cols", "A_vals", "lcon", "ucon", "lvar",
          "uvar", "x0", "y0")


def mixed_bounds(seed=0):
    """Every bound class of common.jl:176-199: fixed, range, only-lower, only-upper and free variables; equality,
    range, only-upper and only-lower rows."""
    qp = P.inequality_qp(seed=seed, n=30, m=18, n_ineq=9)
    rng = np.random.default_rng(seed)
    lvar, uvar = qp.lvar.copy(), qp.uvar.copy()
    lvar[0:3], uvar[0:3] = 0.25, 0.25                 # fixed
    lvar[3:8], uvar[3:8] = 0.0, 5.0                   # range
    lvar[8:12], uvar[8:12] = -np.inf, 4.0             # only upper
    lvar[12:16], uvar[12:16] = -np.inf, np.inf        # free
    lcon, ucon = qp.lcon.copy(), qp.ucon.copy()
    ineq = np.flatnonzero(lcon < ucon)
    lcon[ineq[0:3]] = -np.inf                         # only upper
    ucon[ineq[3:5]] = np.inf                          # only lower
    return dataclasses.replace(qp, lvar=lvar, uvar=uvar, lcon=lcon, ucon=ucon)


@pytest.mark.parametrize("maker,kw", [
    (mixed_bounds, dict(seed=0)), (mixed_bounds, dict(seed=1)), (P.inequality_qp, dict(seed=2)),
    (P.maros_meszaros_like, dict(seed=0, n=200, m=100, nnz=800)), (P.simple_lp, dict()),
    (P.sparse_lp, dict(seed=4, n=1200, m=600)),
])
def test_matches_the_loop_restatement_entry_for_entry(maker, kw):
    qp = maker(**kw)
    a, b = preprocess.standard_form_qp(qp), O.standard_form_qp(qp)
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(np.asarray(x), np.asarray(y)), f


def test_structure_of_the_reformulation():
    qp = mixed_bounds()
    s = preprocess.standard_form_qp(qp)
    n, m = qp.nvar, qp.ncon
    ineq = np.flatnonzero(qp.lcon < qp.ucon)
    ns = ineq.size
    nw = s.nvar - n - ns
    n_rng_x = int(np.sum(np.isfinite(qp.lvar) & np.isfinite(qp.uvar) & (qp.lvar < qp.uvar)))
    n_rng_s = int(np.sum(np.isfinite(qp.lcon) & np.isfinite(qp.ucon) & (qp.lcon < qp.ucon)))
    assert n_rng_x >= 5 and 0 < n_rng_s < ns
    assert s.ncon == m + nw and nw == n_rng_x + n_rng_s           # one new row and one new variable per range bound
    assert (s.lcon == s.ucon).all()                               # equalities only (common.jl:146-152)
    assert (s.lcon[ineq] == 0).all()
    assert np.array_equal(s.lvar[n:n + ns], qp.lcon[ineq])        # bounds of the slacks
    assert (s.lvar[n + ns:] == 0).all() and np.isinf(s.uvar[n + ns:]).all()
    assert np.isinf(s.uvar[3:8]).all()                            # range upper bounds moved into rows
    assert (s.uvar[8:12] == 4.0).all()                            # only-upper bounds stay (common.jl:181-187)
    assert (s.uvar[0:3] == 0.25).all() and (s.lvar[0:3] == 0.25).all()   # fixed variables stay fixed
    assert s.nnzj == qp.nnzj + ns + 2 * nw and s.nnzh == qp.nnzh
    assert np.array_equal(s.A_vals[qp.nnzj:qp.nnzj + ns], -np.ones(ns))
    # the range rows read  x_i + w_k = xu_i
    A = np.zeros((s.ncon, s.nvar)); A[s.A_rows, s.A_cols] = s.A_vals
    assert (A[m:].sum(axis=1) == 2).all() and (np.count_nonzero(A[m:], axis=1) == 2).all()
    assert np.array_equal(s.lcon[m:m + 5], qp.uvar[3:8])         # the first range variables are x_4..x_8


@pytest.mark.parametrize("maker,kw", [(P.inequality_qp, dict(seed=2)), (P.maros_meszaros_like, dict(seed=0, n=200, m=100, nnz=800))])
def test_reformulated_problem_has_the_same_optimum(maker, kw):
    qp = maker(**kw)
    ref = O.MPCOracle(qp).solve()
    got = O.MPCOracle(preprocess.standard_form_qp(qp)).solve()
    assert ref.status == got.status == "SOLVE_SUCCEEDED"
    assert abs(ref.objective - got.objective) <= 1e-6 * max(1.0, abs(ref.objective))
    np.testing.assert_allclose(got.solution[:qp.nvar], ref.solution, atol=1e-5 * max(1.0, np.max(np.abs(ref.solution))))
