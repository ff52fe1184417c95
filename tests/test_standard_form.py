"""standard_form_qp (/root/reference/scripts/common.jl:129-288, the second half of SURVEY 8f N4) against the oracle's
loop-for-loop restatement, entry for entry: the device implementation `madipm_standard_form_qp` through the C ABI
(-m gpu) and the numpy variant `standard_form_qp_host`; and the equivalence of the reformulated problem."""
import dataclasses

import numpy as np
import pytest

from madqp_jl_b200 import _lib, preprocess
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
    a, b = preprocess.standard_form_qp_host(qp), O.standard_form_qp(qp)
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(np.asarray(x), np.asarray(y)), f


def test_structure_of_the_reformulation():
    qp = mixed_bounds()
    s = preprocess.standard_form_qp_host(qp)
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
    got = O.MPCOracle(preprocess.standard_form_qp_host(qp)).solve()
    assert ref.status == got.status == "SOLVE_SUCCEEDED"
    assert abs(ref.objective - got.objective) <= 1e-6 * max(1.0, abs(ref.objective))
    np.testing.assert_allclose(got.solution[:qp.nvar], ref.solution, atol=1e-5 * max(1.0, np.max(np.abs(ref.solution))))


def test_standard_form_entry_point_rejects_bad_arguments():
    lib = _lib.load()
    ns, nw = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
    none = [None] * 10
    rc = lib.madipm_standard_form_qp(0, 2, 1, 0, None, None, None, None, None, None, None, None, None, None,
                                     _lib._p32(ns), _lib._p32(nw), *none, None)
    assert rc == _lib.BAD_ARGUMENT and b"bad argument" in lib.madipm_scale_last_error()          # bounds missing
    rc = lib.madipm_standard_form_qp(0, -1, 0, 0, None, None, None, None, None, None, None, None, None, None,
                                     _lib._p32(ns), _lib._p32(nw), *none, None)
    assert rc == _lib.BAD_ARGUMENT


def test_standard_form_qp_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.MadIPMError) as e:
        preprocess.standard_form_qp(P.simple_lp())
    assert e.value.code == _lib.CUDA_ERROR and "no CPU fallback" in str(e.value)


# ---- GPU: madipm_standard_form_qp against the oracle, entry for entry ----------------------------------------------

def nasty_bounds(seed=3, n=5000, m=3100):
    """Several scan tiles (1024 flags each), every bound class at random positions, NaN-free infinities on both sides,
    rows with lcon > ucon (neither inequality nor range: copied, common.jl:244-246)."""
    rng = np.random.default_rng(seed)
    qp = P.sparse_lp(seed=seed, n=n, m=m)
    kind = rng.integers(0, 5, n)
    lvar = np.where(kind == 0, 0.5, np.where(kind >= 3, -np.inf, rng.uniform(-2, 0, n)))
    uvar = np.where(kind == 0, 0.5, np.where((kind == 2) | (kind == 4), np.inf, rng.uniform(1, 3, n)))
    rk = rng.integers(0, 6, m)
    lcon = np.where(rk == 0, 1.0, np.where((rk == 3) | (rk == 4), -np.inf, rng.uniform(-2, 0, m)))
    ucon = np.where(rk == 0, 1.0, np.where((rk == 2) | (rk == 4), np.inf, rng.uniform(1, 3, m)))
    lcon[rk == 5], ucon[rk == 5] = 2.0, 1.0
    return dataclasses.replace(qp, lvar=lvar, uvar=uvar, lcon=lcon, ucon=ucon)


def no_rows(seed=0):
    qp = mixed_bounds(seed)
    e = np.zeros(0)
    return dataclasses.replace(qp, ncon=0, A_rows=qp.A_rows[:0], A_cols=qp.A_cols[:0], A_vals=e, lcon=e, ucon=e, y0=e)


def all_equalities(seed=0):
    qp = P.sparse_lp(seed=seed, n=300, m=150)
    return dataclasses.replace(qp, lvar=np.zeros(qp.nvar), uvar=np.full(qp.nvar, np.inf))


def all_ranges(seed=0):
    qp = P.inequality_qp(seed=seed, n=1024, m=1024, n_ineq=1024)
    return dataclasses.replace(qp, lvar=np.full(qp.nvar, -1.0), uvar=np.full(qp.nvar, 2.0), lcon=np.full(qp.ncon, -3.0),
                               ucon=np.full(qp.ncon, 4.0))


@pytest.mark.gpu
@pytest.mark.parametrize("maker,kw", [
    (mixed_bounds, dict(seed=0)), (mixed_bounds, dict(seed=1)), (P.inequality_qp, dict(seed=2)),
    (P.maros_meszaros_like, dict(seed=0, n=200, m=100, nnz=800)), (P.simple_lp, dict()),
    (P.sparse_lp, dict(seed=4, n=1200, m=600)), (nasty_bounds, dict()), (nasty_bounds, dict(seed=9, n=1024, m=1024)),
    (no_rows, dict()), (all_equalities, dict()), (all_ranges, dict()),
])
def test_device_standard_form_matches_the_loop_restatement_entry_for_entry(maker, kw):
    qp = maker(**kw)
    info = {}
    a, b = preprocess.standard_form_qp(qp, info=info), O.standard_form_qp(qp)
    for f in FIELDS:
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(np.asarray(x), np.asarray(y)), f
    assert info["launches"] == (3 if qp.nvar + qp.ncon else 0)
    assert info["ns"] == int(np.sum(qp.lcon < qp.ucon)) and a.nvar == qp.nvar + info["ns"] + info["nw"]


@pytest.mark.gpu
def test_device_standard_form_then_solve():
    """The reference's pipeline (scripts/test_gpu.jl:24-29): scale_qp, standard_form_qp, solve -- same optimum as the
    original model."""
    from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
    qp = P.inequality_qp(seed=2)
    ref = O.MPCOracle(qp).solve()
    scaled, _, _, _ = preprocess.scale_qp(qp)
    std = preprocess.standard_form_qp(scaled)
    assert (std.lcon == std.ucon).all()
    got = MPCSolver(std, regularization=FixedRegularization(1e-8, -1e-8)).solve()
    assert got.status[0] == "SOLVE_SUCCEEDED"
    assert abs(ref.objective - got.objective[0]) <= 1e-6 * max(1.0, abs(ref.objective))
