"""K2.5 = MadNLP.ScaledSparseKKTSystem (next row N3 of SURVEY 8f), CPU side: the oracle's restatement against the
reference's own test (/root/reference/test/runtests.jl:102-115) and against the algebra that defines the system
(/root/reference/src/kernels.jl:149-165, /root/reference/scripts/cuda_wrapper.jl:90-116), plus the host side of the
library (same symbolic analysis as K2, loud failure without a GPU).  The GPU parity tests of the K2.5 kernels are the
`*_k25` cases of tests/test_gpu_kernels.py and tests/test_gpu_solve.py."""
import numpy as np
import pytest
import scipy.sparse as sp

from madqp_jl_b200 import _lib
from madqp_jl_b200 import problems as P
from oracle import oracle as O
from conftest import analyze_handle


def dense_kkt(o):
    N = o.N
    K = sp.csc_matrix((o.aug_nzval, o.aug_rowval.astype(np.int64) - 1, o.aug_colptr.astype(np.int64) - 1), shape=(N, N))
    return (K + sp.tril(K, -1).T).toarray()


def at_iterate(qp, kkt, iters=3, reg=("fixed", 1e-8, -1e-8)):
    o = O.MPCOracle(qp, kkt_system=kkt, regularization=reg, max_iter=iters)
    o.initialize()
    o.mpc()
    o.update_regularization()
    o.set_aug_diagonal_reg()
    o.build_kkt()
    return o


@pytest.mark.parametrize("maker,kw", [
    (P.inequality_qp, dict(seed=4)),
    (P.maros_meszaros_like, dict(seed=0, n=400, m=200, nnz=2000)),
    (P.simple_lp, dict()),
])
@pytest.mark.parametrize("reg", [("none",), ("fixed", 1e-8, -1e-8)])
def test_k25_reaches_the_k2_solution_in_the_same_iterations(maker, kw, reg):
    """runtests.jl:102-115 (there on DenseDummyQP with NoRegularization as the K2 reference)."""
    qp = maker(**kw)
    ref = O.MPCOracle(qp, kkt_system="K2", regularization=reg).solve()
    k25 = O.MPCOracle(qp, kkt_system="K2.5", regularization=reg).solve()
    assert k25.status == ref.status == "SOLVE_SUCCEEDED"
    assert k25.k == ref.k
    assert abs(k25.objective - ref.objective) <= 1e-6
    np.testing.assert_allclose(k25.solution, ref.solution, atol=1e-6)
    np.testing.assert_allclose(k25.multipliers, ref.multipliers, atol=1e-6)


@pytest.mark.parametrize("maker,kw", [
    (P.inequality_qp, dict(seed=3, n=60, m=40, n_ineq=14)),
    (P.maros_meszaros_like, dict(seed=1, n=300, m=150, nnz=1500)),
])
def test_k25_matrix_is_the_symmetrically_scaled_k2_matrix(maker, kw):
    """cuda_wrapper.jl:90-116: Hessian entries * s_i * s_j, Jacobian entries * s_j, diagonals copied, with
    pr_diag = s^2 * (K2 pr_diag)  =>  K2.5 = T * K2 * T,  T = diag(s, I)."""
    qp = maker(**kw)
    a = at_iterate(qp, "K2")
    b = at_iterate(qp, "K2.5")
    np.testing.assert_allclose(b.x, a.x, rtol=1e-9, atol=1e-12)          # same trajectory so far
    # take the K2 matrix at exactly b's iterate
    for name in ("x", "xl", "xu", "zl", "zu"):
        getattr(a, name)[:] = getattr(b, name)
    a.del_w, a.del_c = b.del_w, b.del_c
    a.set_aug_diagonal_reg(); a.build_kkt()
    assert (b.l_diag > 0).all() and (b.u_diag > 0).all()                 # kernels.jl:148 "ensure coefficients are positive"
    np.testing.assert_array_equal(b.l_diag, -a.l_diag)
    np.testing.assert_array_equal(b.u_diag, -a.u_diag)
    t = np.concatenate([b.scaling_factor, np.ones(b.m)])
    np.testing.assert_allclose(dense_kkt(b), t[:, None] * dense_kkt(a) * t[None, :], rtol=1e-13, atol=1e-300)
    # variables without bounds are not scaled
    free = np.ones(b.n, bool); free[b.ind_lb] = False; free[b.ind_ub] = False
    assert (b.scaling_factor[free] == 1.0).all()


def test_k25_solve_inverts_the_unreduced_system():
    """solve!(kkt, w) followed by mul!(., kkt, ., 1, 0) gives w back, and K2 / K2.5 return the same direction."""
    qp = P.inequality_qp(seed=3, n=60, m=40, n_ineq=14)
    a = at_iterate(qp, "K2")
    b = at_iterate(qp, "K2.5")
    for name in ("x", "xl", "xu", "zl", "zu"):
        getattr(a, name)[:] = getattr(b, name)
    a.del_w, a.del_c = b.del_w, b.del_c
    a.set_aug_diagonal_reg(); a.build_kkt()
    assert a.linear_solver.factorize() and b.linear_solver.factorize()
    rng = np.random.default_rng(0)
    v = rng.standard_normal(b.d.size)
    db = b.kkt_solve(v.copy())
    da = a.kkt_solve(v.copy())
    back = b.kkt_mul(np.zeros_like(v), db.copy(), 1.0, 0.0)
    assert np.max(np.abs(back - v)) <= 1e-9 * max(1.0, np.max(np.abs(v)))
    np.testing.assert_allclose(db, da, rtol=1e-6, atol=1e-8 * np.max(np.abs(da)))
    # the two unreduced operators are the same operator (only the stored sign of l_diag / u_diag differs)
    w = rng.standard_normal(b.d.size)
    np.testing.assert_allclose(b.kkt_mul(w.copy(), v.copy(), -0.5, 2.0), a.kkt_mul(w.copy(), v.copy(), -0.5, 2.0),
                               rtol=1e-13, atol=1e-13)


def test_k25_initial_system_is_unscaled():
    """MadNLP.initialize!(kkt): scaling_factor = 1, so the starting-point solves (solver.jl:16-33) see the K2 matrix."""
    qp = P.inequality_qp(seed=1)
    a = O.MPCOracle(qp, kkt_system="K2"); a.initialize()
    b = O.MPCOracle(qp, kkt_system="K2.5"); b.initialize()
    np.testing.assert_allclose(b.x, a.x, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(b.y, a.y, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(b.zl, a.zl, rtol=1e-10, atol=1e-12)


def test_unknown_kkt_system_is_rejected():
    with pytest.raises(ValueError):
        O.MPCOracle(P.simple_lp(), kkt_system="K3")


def test_scaled_handle_shares_the_k2_symbolic_analysis():
    """MADIPM_KKT_K2_SCALED: same pattern entry point, permutation, supernodes and assembly map as MADIPM_KKT_K2."""
    qp = P.maros_meszaros_like(seed=2, n=300, m=150, nnz=1500)
    h2, _ = analyze_handle(qp, "K2")
    h25, _ = analyze_handle(qp, "K2.5")
    assert h25.info()["kkt_kind"] == _lib.KKT_K2_SCALED and h2.info()["kkt_kind"] == _lib.KKT_K2
    for what in ("PERM", "ETREE_PARENT", "COLCOUNT", "SUPER_FIRST", "SUPER_ROWPTR", "ASM_MAP", "KKT_COLPTR", "KKT_ROWVAL",
                 "AUG_CSC_MAP"):
        assert np.array_equal(h25.symbolic(what), h2.symbolic(what)), what
    assert h25.vec_len("SCALING") == h25.info()["n_tot"]
    h2.close(); h25.close()


def test_scaled_handle_rejects_the_normal_pattern_and_bad_kinds():
    h = _lib.Handle(kkt_kind=_lib.KKT_K2_SCALED)
    with pytest.raises(_lib.MadIPMError) as e:
        h.set_pattern_normal(2, 1, np.array([1]), np.array([1]), np.zeros(0, dtype=np.int64))
    assert e.value.code == _lib.BAD_ARGUMENT
    h.close()
    with pytest.raises(_lib.MadIPMError):
        _lib.Handle(kkt_kind=7)


def test_scaled_numeric_calls_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    qp = P.simple_lp()
    h, _ = analyze_handle(qp, "K2.5")
    with pytest.raises(_lib.MadIPMError) as e:
        h.set_aug_diagonal_reg(1e-8, -1e-8)
    assert e.value.code == _lib.CUDA_ERROR and "no CPU fallback" in str(e.value)
    h.close()
