"""fixed_variable_treatment = MakeParameter (/root/reference/src/utils.jl:81; the reference's own test names the case,
/root/reference/test/runtests.jl:69-73): the host-side reduction to the free variables (madqp.jl_b200/preprocess.py,
vectorised) against the oracle's loop restatement entry for entry, the reduced model against independent solutions of
the full one, and the whole solve on the GPU."""
import dataclasses

import numpy as np
import pytest

from madqp_jl_b200 import preprocess
from madqp_jl_b200 import problems as P
from oracle import oracle as O

FIELDS = ("nvar", "ncon", "c", "H_rows", "H_cols", "H_vals", "A_rows", "A_cols", "A_vals", "lcon", "ucon", "lvar", "uvar", "x0", "y0")


_OPT = {}


def with_fixed(qp, seed=0, frac=0.2):
    """A random fifth of the variables fixed at their values in the optimum of `qp` (so the model stays feasible, and
    its optimum is known: the same point)."""
    key = (qp.name, qp.nvar, qp.ncon, float(np.sum(qp.c)))
    if key not in _OPT:
        ref = O.MPCOracle(qp).solve()
        assert ref.status == "SOLVE_SUCCEEDED"
        _OPT[key] = ref
    val = _OPT[key].solution
    fix = np.random.default_rng(seed).random(qp.nvar) < frac
    return dataclasses.replace(qp, lvar=np.where(fix, val, qp.lvar), uvar=np.where(fix, val, qp.uvar)), _OPT[key]


CASES = [(P.maros_meszaros_like, dict(seed=0, n=60, m=30, nnz=240)), (P.inequality_qp, dict(seed=2)),
         (P.sparse_lp, dict(seed=4, n=300, m=150)), (P.maros_meszaros_like, dict(seed=5, n=200, m=100, nnz=800))]


@pytest.mark.parametrize("maker,kw", CASES)
def test_matches_the_loop_restatement(maker, kw):
    qp, _ = with_fixed(maker(**kw), seed=1)
    a, fa, ra, xa = preprocess.make_parameter(qp)
    b, fb, rb, xb = O.make_parameter(qp)
    assert fa.size > 0 and np.array_equal(fa, fb) and np.array_equal(ra, rb) and np.array_equal(xa, xb)
    for f in FIELDS:
        x, y = np.asarray(getattr(a, f)), np.asarray(getattr(b, f))
        if f in ("c", "lcon", "ucon"):                               # sums over the fixed entries: order of additions only
            np.testing.assert_allclose(x, y, rtol=1e-13, atol=1e-13, err_msg=f)
        else:
            assert np.array_equal(x, y), f
    assert abs(a.c0 - b.c0) <= 1e-12 * max(1.0, abs(b.c0))


def test_nothing_fixed_is_the_identity():
    qp = P.inequality_qp(seed=2)
    red, fixed, free, xf = preprocess.make_parameter(qp)
    assert red is qp and fixed.size == 0 and np.array_equal(free, np.arange(qp.nvar)) and xf.size == 0


def test_objective_and_constraints_agree_at_any_point():
    """q_reduced(x_free) == q(x) and A_reduced x_free - (lcon', ucon') == A x - (lcon, ucon) with x_fixed in place."""
    qp, _ = with_fixed(P.maros_meszaros_like(seed=3, n=80, m=40, nnz=400), seed=7, frac=0.3)
    red, fixed, free, xf = preprocess.make_parameter(qp)
    rng = np.random.default_rng(0)
    x = rng.normal(size=qp.nvar); x[fixed] = xf

    def evaluate(q, v):
        H = np.zeros((q.nvar, q.nvar)); H[q.H_rows, q.H_cols] = q.H_vals
        H = H + np.tril(H, -1).T
        A = np.zeros((q.ncon, q.nvar)); np.add.at(A, (q.A_rows, q.A_cols), q.A_vals)
        return q.c0 + q.c @ v + 0.5 * v @ H @ v, A @ v
    f_full, ax_full = evaluate(qp, x)
    f_red, ax_red = evaluate(red, x[free])
    assert abs(f_full - f_red) <= 1e-10 * max(1.0, abs(f_full))
    np.testing.assert_allclose(ax_full - qp.lcon, ax_red - red.lcon, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(ax_full - qp.ucon, ax_red - red.ucon, rtol=1e-10, atol=1e-10)


def test_lp_with_fixed_variables_against_highs():
    from scipy.optimize import linprog
    from scipy.sparse import coo_matrix
    qp, _ = with_fixed(P.sparse_lp(seed=4, n=300, m=150), seed=2)
    assert np.any(qp.lvar == qp.uvar)
    got = O.MPCOracle(qp).solve()
    assert got.status == "SOLVE_SUCCEEDED"
    A = coo_matrix((qp.A_vals, (qp.A_rows, qp.A_cols)), shape=(qp.ncon, qp.nvar)).tocsr()
    eq = qp.lcon == qp.ucon
    assert eq.all()
    ref = linprog(qp.c, A_eq=A, b_eq=qp.lcon, bounds=list(zip(np.where(np.isfinite(qp.lvar), qp.lvar, None),
                                                             np.where(np.isfinite(qp.uvar), qp.uvar, None))), method="highs")
    assert ref.status == 0
    assert abs(got.objective - (ref.fun + qp.c0)) <= 1e-6 * max(1.0, abs(ref.fun))
    fixed = np.flatnonzero(qp.lvar == qp.uvar)
    assert np.array_equal(got.solution[fixed], qp.lvar[fixed]) and got.solution.size == qp.nvar


@pytest.mark.parametrize("maker,kw", CASES)
def test_fixing_variables_at_the_optimum_keeps_the_optimum(maker, kw):
    """Independent of the reduction: the model with a fifth of its variables fixed at their optimal values has the same
    optimum, and the reduced solve must find it (objective incl. the constant term that moved into c0)."""
    qp, ref = with_fixed(maker(**kw), seed=1)
    fixed = np.flatnonzero(qp.lvar == qp.uvar)
    got = O.MPCOracle(qp).solve()
    assert got.status == "SOLVE_SUCCEEDED" and fixed.size > 0 and got.nx == qp.nvar - fixed.size
    assert abs(got.objective - ref.objective) <= 1e-6 * max(1.0, abs(ref.objective))
    assert got.solution.size == qp.nvar and np.array_equal(got.solution[fixed], qp.lvar[fixed])
    if qp.nnzh:                                                      # (an LP's optimal face need not be a point)
        np.testing.assert_allclose(got.solution, ref.solution, atol=1e-4 * max(1.0, np.max(np.abs(ref.solution))))


# ---- GPU: the whole solve with fixed variables against the oracle ---------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("maker,kw", CASES[:3])
def test_gpu_solve_with_fixed_variables(maker, kw):
    from madqp_jl_b200.mpc import MPCSolver
    qp, _ = with_fixed(maker(**kw), seed=1)
    s = MPCSolver(qp)
    st = s.solve()
    ref = O.MPCOracle(qp, perm=s.h.symbolic("PERM")).solve()
    assert st.status[0] == ref.status == "SOLVE_SUCCEEDED"
    assert abs(int(st.iter[0]) - ref.k) <= 1
    assert abs(st.objective[0] - ref.objective) <= 1e-8 * max(1.0, abs(ref.objective))
    assert st.solution.shape == (1, qp.nvar)
    fixed = np.flatnonzero(qp.lvar == qp.uvar)
    assert np.array_equal(st.solution[0][fixed], qp.lvar[fixed])
    np.testing.assert_allclose(st.solution[0], ref.solution, atol=1e-5 * max(1.0, np.max(np.abs(ref.solution))))
    s.close()


@pytest.mark.gpu
def test_gpu_batch_with_fixed_variables():
    """Three instances with one pattern and the same fixed variables (feasible by construction, different costs)."""
    from madqp_jl_b200.mpc import MPCSolver
    qpf, _ = with_fixed(P.maros_meszaros_like(seed=0, n=60, m=30, nnz=240), seed=1)
    qps = [dataclasses.replace(qpf, c=qpf.c * (1.0 + 0.1 * b)) for b in range(3)]
    s = MPCSolver(qps)
    st = s.solve()
    assert st.solution.shape == (3, qpf.nvar)
    fixed = np.flatnonzero(qpf.lvar == qpf.uvar)
    for b, q in enumerate(qps):
        ref = O.MPCOracle(q, perm=s.h.symbolic("PERM")).solve()
        assert st.status[b] == ref.status == "SOLVE_SUCCEEDED"
        assert abs(st.objective[b] - ref.objective) <= 1e-7 * max(1.0, abs(ref.objective))
        assert np.array_equal(st.solution[b][fixed], q.lvar[fixed])
    s.close()
