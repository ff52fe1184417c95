"""scale_qp (/root/reference/scripts/common.jl:46-100): the oracle's restatement of the Ruiz
equilibration + model rescaling (CPU tests), and the device implementation `madipm_scale_qp` against it through the
C ABI (GPU tests: bit-exact, a maximum does not depend on the order of its operands)."""
import dataclasses

import numpy as np
import pytest
import scipy.sparse as sp

from madqp_jl_b200 import _lib, preprocess
from madqp_jl_b200 import problems as P
from oracle import oracle as O


def badly_scaled(qp, seed=0, row_decades=3, col_decades=2):
    """The same QP with rows scaled by 10^U(-r, r) and variables substituted x = cs * x'."""
    rng = np.random.default_rng(seed)
    rs = 10.0 ** rng.uniform(-row_decades, row_decades, qp.ncon)
    cs = 10.0 ** rng.uniform(-col_decades, col_decades, qp.nvar)
    return dataclasses.replace(
        qp, A_vals=qp.A_vals * rs[qp.A_rows] * cs[qp.A_cols], lcon=qp.lcon * rs, ucon=qp.ucon * rs, lvar=qp.lvar / cs,
        uvar=qp.uvar / cs, x0=qp.x0 / cs, c=qp.c * cs, H_vals=qp.H_vals * cs[qp.H_rows] * cs[qp.H_cols])


def norms(qp):
    A = sp.coo_matrix((np.abs(qp.A_vals), (qp.A_rows, qp.A_cols)), shape=(qp.ncon, qp.nvar)).tocsr()
    r = np.asarray(A.max(axis=1).todense()).ravel()
    c = np.asarray(A.max(axis=0).todense()).ravel()
    return r, c


# ---- CPU: the oracle ---------------------------------------------------------------------------------------------

def test_oracle_equilibrates_rows_and_columns():
    bad = badly_scaled(P.maros_meszaros_like(seed=0, n=400, m=200, nnz=2000))
    r0, c0 = norms(bad)
    assert r0.max() / r0[r0 > 0].min() > 1e4
    prev = np.inf
    for sweeps in (1, 5, 10, 30):
        s, Dr, Dc, done, err = O.scale_qp(bad, max_iter=sweeps)
        r, c = norms(s)
        dev = max(np.abs(1 - r[r > 0]).max(), np.abs(1 - c[c > 0]).max())
        assert done == sweeps and dev < prev
        prev = dev
        assert (Dc[c0 == 0] == 1.0).all() and (Dr[r0 == 0] == 1.0).all()      # empty rows / columns keep factor 1
    assert prev < 1e-6                                                        # linear convergence, 30 sweeps
    s, _, _, done, err = O.scale_qp(bad, max_iter=100, tol=1e-3)
    assert done < 100 and err <= 1e-3


def test_oracle_scaling_formulas_follow_the_reference():
    """common.jl:38-44, 70-97: A ./ (Dr_i Dc_j), H ./ (Dc_i Dc_j), c ./ Dc, lvar/uvar/x0 .* Dc, lcon/ucon/y0 ./ Dr."""
    qp = badly_scaled(P.inequality_qp(seed=2))
    s, Dr, Dc, _, _ = O.scale_qp(qp)
    assert np.array_equal(s.A_vals, qp.A_vals / (Dr[qp.A_rows] * Dc[qp.A_cols]))
    assert np.array_equal(s.H_vals, qp.H_vals / (Dc[qp.H_rows] * Dc[qp.H_cols]))
    assert np.array_equal(s.c, qp.c / Dc) and np.array_equal(s.x0, qp.x0 * Dc) and np.array_equal(s.y0, qp.y0 / Dr)
    for a, b, D, div in ((s.lvar, qp.lvar, Dc, False), (s.uvar, qp.uvar, Dc, False), (s.lcon, qp.lcon, Dr, True),
                         (s.ucon, qp.ucon, Dr, True)):
        assert np.array_equal(a, b / D if div else b * D)
        assert np.array_equal(np.isinf(a), np.isinf(b))                       # infinite bounds stay infinite
    assert s.c0 == qp.c0 and s.nvar == qp.nvar and s.ncon == qp.ncon


def test_scaled_problem_has_the_same_optimum():
    """x = xs ./ Dc maps solutions onto each other, so the objective value is unchanged (benchmarks_cpu.jl:29-30 solves
    the scaled model and reports its objective)."""
    qp = P.maros_meszaros_like(seed=0, n=400, m=200, nnz=2000)
    bad = badly_scaled(qp)
    s, Dr, Dc, _, _ = O.scale_qp(bad)
    ref = O.MPCOracle(qp).solve()
    got = O.MPCOracle(s).solve()
    assert got.status == ref.status == "SOLVE_SUCCEEDED"
    assert abs(got.objective - ref.objective) <= 1e-6 * max(1.0, abs(ref.objective))
    # and back in the variables of `bad`
    direct = O.MPCOracle(bad).solve()
    np.testing.assert_allclose(got.solution / Dc, direct.solution, rtol=1e-4, atol=1e-5 * np.max(np.abs(direct.solution)))


def test_scale_entry_point_rejects_bad_arguments():
    lib = _lib.load()
    opt = _lib.ScaleOptions()
    lib.madipm_default_scale_options(opt)
    assert (opt.max_iter, opt.tol, opt.device) == (10, 0.0, 0)
    I = np.array([3], dtype=np.int32); J = np.array([1], dtype=np.int32); V = np.array([1.0])
    rc = lib.madipm_scale_qp(opt, 2, 2, 1, _lib._p32(I), _lib._p32(J), _lib._pd(V), 0, None, None, None,
                             None, None, None, None, None, None, None, None, None, None, None)
    assert rc == _lib.BAD_ARGUMENT and b"out of range" in lib.madipm_scale_last_error()
    assert V[0] == 1.0                                                         # untouched on error


def test_scale_qp_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.MadIPMError) as e:
        preprocess.scale_qp(P.simple_lp())
    assert e.value.code == _lib.CUDA_ERROR and "no CPU fallback" in str(e.value)


# ---- GPU: madipm_scale_qp against the oracle ---------------------------------------------------------------------

def same_model(a, b):
    for f in ("A_vals", "H_vals", "c", "lvar", "uvar", "lcon", "ucon", "x0", "y0"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.gpu
@pytest.mark.parametrize("maker,kw", [
    (P.maros_meszaros_like, dict(seed=0, n=400, m=200, nnz=2000)),
    (P.inequality_qp, dict(seed=2)),
    (P.sparse_lp, dict(seed=4, n=1200, m=600)),
    (P.transportation_lp, dict(n_src=60, n_dst=80, n_arcs=1500)),
    (P.simple_lp, dict()),
])
@pytest.mark.parametrize("sweeps,tol", [(10, 0.0), (1, 0.0), (0, 0.0), (100, 1e-3)])
def test_device_scaling_is_bit_exact(maker, kw, sweeps, tol):
    bad = badly_scaled(maker(**kw))
    ref, Dr, Dc, done, err = O.scale_qp(bad, max_iter=sweeps, tol=tol)
    got, gDr, gDc, info = preprocess.scale_qp(bad, max_iter=sweeps, tol=tol)
    assert info["sweeps"] == done and info["launches"] > 0
    assert np.array_equal(gDr, Dr) and np.array_equal(gDc, Dc)
    same_model(got, ref)
    if sweeps:
        assert info["err"] == err


@pytest.mark.gpu
def test_device_scaling_ragged_and_empty():
    """Empty rows and columns keep factor 1; a model without constraints or without a Hessian passes through."""
    rng = np.random.default_rng(5)
    n, m = 50, 30
    rows = rng.integers(0, m // 2, 80) * 2            # odd rows empty
    cols = rng.integers(0, n // 2, 80)                # upper half of the columns empty
    key = np.unique(rows * n + cols)
    rows, cols = key // n, key % n
    qp = P.QPData(name="ragged", nvar=n, ncon=m, c=rng.standard_normal(n), c0=0.0, H_rows=np.arange(n), H_cols=np.arange(n),
                  H_vals=1.0 + rng.random(n), A_rows=rows, A_cols=cols, A_vals=10.0 ** rng.uniform(-4, 4, rows.size),
                  lcon=np.full(m, -np.inf), ucon=rng.random(m), lvar=np.zeros(n), uvar=np.full(n, np.inf),
                  x0=np.ones(n), y0=np.zeros(m))
    ref, Dr, Dc, _, _ = O.scale_qp(qp)
    got, gDr, gDc, _ = preprocess.scale_qp(qp)
    assert np.array_equal(gDr, Dr) and np.array_equal(gDc, Dc)
    assert (gDr[1::2] == 1.0).all() and (gDc[n // 2:] == 1.0).all()
    same_model(got, ref)
    empty = dataclasses.replace(qp, ncon=0, A_rows=np.zeros(0, int), A_cols=np.zeros(0, int), A_vals=np.zeros(0),
                                lcon=np.zeros(0), ucon=np.zeros(0), y0=np.zeros(0))
    got, gDr, gDc, _ = preprocess.scale_qp(empty)
    assert gDr.size == 0 and (gDc == 1.0).all() and np.array_equal(got.H_vals, qp.H_vals) and np.array_equal(got.c, qp.c)


@pytest.mark.gpu
def test_device_scaling_full_size_properties():
    """C2-sized matrix (1M nonzeros): after 30 sweeps every non-empty row and column has infinity norm 1 to 1e-6,
    scaling the scaled model again changes nothing beyond that, and the solver reaches the optimum of the original."""
    qp = P.sparse_lp(seed=0, n=200_000, m=100_000)
    bad = badly_scaled(qp, seed=1)
    got, Dr, Dc, info = preprocess.scale_qp(bad, max_iter=30)
    r, c = norms(got)
    assert np.abs(1 - r[r > 0]).max() < 1e-6 and np.abs(1 - c[c > 0]).max() < 1e-6
    again, Dr2, Dc2, _ = preprocess.scale_qp(got, max_iter=3)
    assert np.abs(Dr2 - 1).max() < 1e-6 and np.abs(Dc2 - 1).max() < 1e-6
    assert np.array_equal(got.A_vals, bad.A_vals / (Dr[bad.A_rows] * Dc[bad.A_cols]))   # _scale_coo! with the returned factors


@pytest.mark.gpu
def test_solver_on_device_scaled_model_matches_original():
    from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
    qp = P.maros_meszaros_like(seed=0, n=400, m=200, nnz=2000)
    bad = badly_scaled(qp)
    scaled, Dr, Dc, _ = preprocess.scale_qp(bad)
    a = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8)); sa = a.solve(); a.close()
    b = MPCSolver(scaled, regularization=FixedRegularization(1e-8, -1e-8)); sb = b.solve(); b.close()
    assert sa.status[0] == sb.status[0] == "SOLVE_SUCCEEDED"
    assert abs(sa.objective[0] - sb.objective[0]) <= 1e-6 * max(1.0, abs(sa.objective[0]))
