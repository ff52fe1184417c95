"""CPU tests of the oracle itself: pinned against the golden fixtures (definitions of the reference's
in-repo functions), the reference's one known answer, and independent dense / HiGHS solves."""
import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from madqp_jl_b200 import problems as P
from oracle import oracle as O


def test_coo_to_csr_golden(golden):
    for c in golden["sparse_cases"]:
        Bp, Bj, Bx = O.coo_to_csr(c["m"], c["n"], np.array(c["I"], dtype=np.int32), np.array(c["J"], dtype=np.int32),
                                  np.array(c["V"], dtype=float))
        assert Bp.tolist() == c["Bp"] and Bj.tolist() == c["Bj"] and Bx.tolist() == c["Bx"]
        # the A_csr_map trick of normalkkt.jl:83-88: values 1:nnz come back as the permutation
        _, _, order = O.coo_to_csr(c["m"], c["n"], np.array(c["I"], dtype=np.int32), np.array(c["J"], dtype=np.int32),
                                   np.arange(1, len(c["I"]) + 1, dtype=float))
        assert order.astype(int).tolist() == c["order"]


def test_build_and_assemble_normal_golden(golden):
    for c in golden["sparse_cases"]:
        if "Cp" not in c:
            continue
        Bp, Bj, Bx = O.coo_to_csr(c["m"], c["n"], np.array(c["I"], dtype=np.int32), np.array(c["J"], dtype=np.int32),
                                  np.array(c["V"], dtype=float))
        Cp, Cj = O.build_normal_system(c["m"], c["n"], Bp, Bj)
        assert Cp.tolist() == c["Cp"] and Cj.tolist() == c["Cj"]
        Cp2, Cj2 = O.build_normal_system_fast(c["m"], c["n"], Bp, Bj)
        assert Cp2.tolist() == c["Cp"] and Cj2.tolist() == c["Cj"]
        Cx = np.zeros(len(c["Cj"]))
        O.assemble_normal_system(c["m"], c["n"], Bp, Bj, Bx, Cp, Cj, Cx, np.array(c["D"], dtype=float))
        assert Cx.tolist() == c["Cx"]          # exact: integer / dyadic data


def test_transfer_golden(golden):
    t = golden["transfer"]
    dest = np.full(t["ndest"], 99.0)
    O.transfer(dest, np.array(t["src"], dtype=float), np.array(t["map"], dtype=np.int64))
    assert dest.tolist() == t["dest"]


def test_argmin_fold_golden(golden):
    for f in golden["argmin_fold"]:
        vals = np.array([np.inf if v == "inf" else v for v in f["vals"]], dtype=float)
        assert O.MPCOracle._argmin_fold(vals) == (f["value"], f["index"])


def test_coo_to_csc_merges_duplicates():
    I = np.array([1, 3, 2, 3, 3]); J = np.array([1, 1, 2, 1, 3])
    colptr, rowval, m = O.coo_to_csc(3, 3, I, J)
    assert colptr.tolist() == [1, 3, 4, 5] and rowval.tolist() == [1, 3, 2, 3] and m.tolist() == [1, 2, 3, 2, 4]


@pytest.mark.parametrize("n,perm_seed", [(1, None), (30, None), (30, 3), (80, 5)])
def test_ldl_against_dense(n, perm_seed):
    rng = np.random.default_rng(n)
    M = sp.random(n, n, density=0.15, random_state=1, format="csr")
    K = (M + M.T).toarray() + np.diag(np.where(np.arange(n) < n // 2, 10.0, -10.0))   # quasi-definite-like
    L = sp.tril(sp.csc_matrix(K), format="csc")
    L.sort_indices()
    perm = None if perm_seed is None else np.random.default_rng(perm_seed).permutation(n)
    s = O.LDLSolver(n, (L.indptr + 1).astype(np.int32), (L.indices + 1).astype(np.int32), L.data.copy(), perm)
    assert s.factorize()
    b = rng.standard_normal(n)
    x = s.solve(b.copy())
    assert np.max(np.abs(K @ x - b)) <= 1e-10 * max(1.0, np.max(np.abs(b)))
    pos, zero, neg = s.inertia()
    ev = np.linalg.eigvalsh(K)
    assert (pos, zero, neg) == (int((ev > 0).sum()), 0, int((ev < 0).sum()))


def test_ldl_zero_pivot_reports_failure():
    s = O.LDLSolver(2, np.array([1, 3, 4], dtype=np.int32), np.array([1, 2, 2], dtype=np.int32), np.array([0.0, 1.0, 1.0]))
    assert not s.factorize()


def test_simple_lp_known_answer(golden):
    """/root/reference/test/runtests.jl:24-55, 139-180: objective 1.0; NormalKKT agrees with K2 to 1e-6."""
    qp = P.simple_lp()
    ref = O.MPCOracle(qp, regularization=("none",)).solve()
    assert ref.status == O.SOLVE_SUCCEEDED
    assert abs(ref.objective - golden["simple_lp"]["objective"]) <= 1e-8
    nrm = O.MPCOracle(qp, kkt_system="NORMAL").solve()
    assert nrm.status == O.SOLVE_SUCCEEDED
    assert abs(nrm.objective - ref.objective) <= 1e-6
    assert np.allclose(nrm.solution, ref.solution, atol=1e-6)
    assert np.allclose(nrm.multipliers, ref.multipliers, atol=1e-6)


def test_regularised_matches_unregularised():
    """/root/reference/test/runtests.jl:117-135."""
    qp = P.inequality_qp(seed=1)
    ref = O.MPCOracle(qp, regularization=("none",)).solve()
    for reg in [("fixed", 1e-8, -1e-9), ("adaptive", 1e-8, -1e-9, 1e-9)]:
        s = O.MPCOracle(qp, regularization=reg).solve()
        assert s.status == O.SOLVE_SUCCEEDED
        assert abs(s.objective - ref.objective) <= 1e-6
        assert np.allclose(s.solution, ref.solution, atol=1e-6)


@pytest.mark.parametrize("rule", [("adaptive", 0.99), ("conservative", 0.99), ("mehrotra", 0.99)])
def test_step_rules(rule):
    """/root/reference/test/runtests.jl:80-92: status only."""
    assert O.MPCOracle(P.inequality_qp(seed=2), step_rule=rule).solve().status == O.SOLVE_SUCCEEDED


def test_lp_against_highs():
    qp = P.sparse_lp(n=600, m=300)
    A = sp.csr_matrix((qp.A_vals, (qp.A_rows, qp.A_cols)), shape=(qp.ncon, qp.nvar))
    r = linprog(qp.c, A_eq=A, b_eq=qp.lcon, bounds=[(0, None)] * qp.nvar, method="highs")
    for kkt in ("K2", "NORMAL"):
        s = O.MPCOracle(qp, kkt_system=kkt, regularization=("fixed", 1e-8, -1e-8)).solve()
        assert s.status == O.SOLVE_SUCCEEDED
        assert abs(s.objective - r.fun) <= 1e-6 * max(1.0, abs(r.fun))


def test_qp_reaches_constructed_kkt_point():
    qp = P.maros_meszaros_like(seed=3, n=300, m=150, nnz=1200)
    s = O.MPCOracle(qp, regularization=("fixed", 1e-8, -1e-8), max_ncorr=3).solve()
    assert s.status == O.SOLVE_SUCCEEDED
    assert abs(s.objective - qp.meta["obj_star"]) <= 1e-5 * max(1.0, abs(qp.meta["obj_star"]))


def test_normal_kkt_system_consistency():
    """MadNLPTests.test_kkt_system analogue (/root/reference/test/runtests.jl:149-163):
    solve-then-multiply round trip on the oracle's NormalKKTSystem."""
    qp = P.sparse_lp(n=200, m=100)
    s = O.MPCOracle(qp, kkt_system="NORMAL", regularization=("fixed", 1e-8, 0.0))
    s.initialize()
    s.update_regularization()
    s.factorize_regularized_system()
    s.set_predictive_rhs()
    s.solve_system(s.d, s.p)
    assert s.last_residual_ratio <= 1e-10


def test_scale_augmented_system_coo_golden(golden):
    """K2.5 build, /root/reference/scripts/cuda_wrapper.jl:90-111: the oracle against the kernel's branch list applied
    entry by entry (dyadic values, exact products -> bit-exact)."""
    g = golden["scale_aug_coo"]
    dest = O.scale_augmented_system_coo(np.array(g["I"]), np.array(g["J"]), np.array(g["V"]), np.array(g["scaling"]), g["n"], g["m"])
    assert None not in g["dest"]                       # every entry of the fixture falls in one of the four blocks
    assert np.array_equal(dest, np.array(g["dest"]))
    # blocks: first n copied, last m copied, something in between actually scaled
    n, m = g["n"], g["m"]
    assert np.array_equal(dest[:n], np.array(g["V"])[:n]) and np.array_equal(dest[-m:], np.array(g["V"])[-m:])
    assert not np.array_equal(dest[n:-m], np.array(g["V"])[n:-m])
