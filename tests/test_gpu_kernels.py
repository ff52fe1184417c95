"""GPU parity tests, kernel by kernel, through the C ABI: the device state is loaded from the
oracle's state at a real iterate, one entry point is called, and the result is compared with the
oracle's restatement of the same reference function.  Bit-exact where the operation order is
fixed by the reference (index work, divisions, the normal-equations assembly); otherwise within
the floating-point tolerance written in each test (FMA contraction / summation order only)."""
import numpy as np
import pytest
import scipy.sparse as sp

from madqp_jl_b200 import _lib
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
from oracle import oracle as O

pytestmark = pytest.mark.gpu

REG = ("fixed", 1e-8, -1e-8)
CASES = {
    "qp_k2": (lambda: P.maros_meszaros_like(seed=1, n=300, m=150, nnz=1500), "K2"),
    "ineq_k2": (lambda: P.inequality_qp(seed=3, n=60, m=40, n_ineq=14), "K2"),
    "banded_k2": (lambda: P.banded_qp(seed=2, n=4000, m=2000), "K2"),
    "lp_normal": (lambda: P.sparse_lp(seed=4, n=1200, m=600), "NORMAL"),
    "lp_k2": (lambda: P.sparse_lp(seed=5, n=800, m=400), "K2"),
    # K2.5 = MadNLP.ScaledSparseKKTSystem (/root/reference/src/kernels.jl:149-165, scripts/cuda_wrapper.jl:90-116)
    "qp_k25": (lambda: P.maros_meszaros_like(seed=1, n=300, m=150, nnz=1500), "K2.5"),
    "ineq_k25": (lambda: P.inequality_qp(seed=3, n=60, m=40, n_ineq=14), "K2.5"),
    "banded_k25": (lambda: P.banded_qp(seed=2, n=4000, m=2000), "K2.5"),
}


def oracle_at_iterate(qp, kkt, iters=2, perm=None):
    """Oracle advanced `iters` MPC iterations, stopped right after factorize_regularized_system.
    `perm`: the library's fill-reducing permutation, so both sides eliminate in the same order
    (LDL' without pivoting on a quasi-definite matrix is only as accurate as its pivot order)."""
    o = O.MPCOracle(qp, perm=perm, kkt_system=kkt, regularization=REG, max_iter=iters)
    o.initialize()
    o.status = o.mpc()            # stops with MAXIMUM_ITERATIONS_EXCEEDED at k == iters
    o.update_regularization()
    o.set_aug_diagonal_reg()
    return o


def new_solver(qp, kkt, **kw):
    return MPCSolver(qp, kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8), **kw)


def device_like(o, qp, kkt, s=None, **kw):
    s = s or new_solver(qp, kkt, **kw)
    h = s.h
    for name, arr in (("X", o.x), ("XL", o.xl), ("XU", o.xu), ("ZL", o.zl), ("ZU", o.zu), ("F", o.f), ("JACL", o.jacl),
                      ("COST", o._cost), ("Y", o.y), ("C", o.c), ("RHS", o.rhs), ("D", o.d), ("P", o.p),
                      ("CORR_LB", o.correction_lb), ("CORR_UB", o.correction_ub)):
        h.vec_set(name, arr)
    h.set_values(o.hess_V, o.jac_V[:qp.nnzj])
    return s, h


@pytest.fixture(scope="module", params=list(CASES))
def state(request):
    maker, kkt = CASES[request.param]
    qp = maker()
    s = new_solver(qp, kkt)
    o = oracle_at_iterate(qp, kkt, perm=s.h.symbolic("PERM"))
    s, h = device_like(o, qp, kkt, s)
    yield qp, kkt, o, s, h
    s.close()


def test_spmv_matches_scipy(state):
    qp, kkt, o, s, h = state
    rng = np.random.default_rng(0)
    A = o._Amat
    x, y = rng.standard_normal(o.n), rng.standard_normal(o.m)
    h.vec_set("X", x); h.vec_set("Y", y)
    h.spmv(_lib.OP_A, "C", "X")
    np.testing.assert_allclose(h.vec_get("C"), A @ x, rtol=1e-13, atol=1e-13)
    h.spmv(_lib.OP_AT, "JACL", "Y")
    np.testing.assert_allclose(h.vec_get("JACL"), A.T @ y, rtol=1e-13, atol=1e-13)
    if kkt != "NORMAL":
        h.spmv(_lib.OP_Q, "F", "X")
        np.testing.assert_allclose(h.vec_get("F"), o._Hsym() @ x, rtol=1e-13, atol=1e-13)
    h.vec_set("X", o.x); h.vec_set("Y", o.y); h.vec_set("C", o.c); h.vec_set("JACL", o.jacl); h.vec_set("F", o.f)


def test_set_aug_diagonal_reg_bit_exact(state):
    """K1, /root/reference/src/kernels.jl:128-146 (K2.5: 149-165 + _set_aug_diagonal!): same operations in the same
    order -> identical bits, including the K2.5 scaling factors (IEEE square roots on both sides)."""
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c)
    for name, ref in (("REG", o.reg), ("PR_DIAG", o.pr_diag), ("DU_DIAG", o.du_diag), ("L_DIAG", o.l_diag),
                      ("U_DIAG", o.u_diag), ("L_LOWER", o.l_lower), ("U_LOWER", o.u_lower)):
        assert np.array_equal(h.vec_get(name), ref), name
    if kkt == "K2.5":
        assert np.array_equal(h.vec_get("SCALING"), o.scaling_factor)
        assert (h.vec_get("L_DIAG") > 0).all() and (h.vec_get("U_DIAG") > 0).all()      # kernels.jl:148
        assert (h.vec_get("PR_DIAG") > 0).all()


def test_build_kkt_bit_exact(state):
    """K2a (transfer!), K2b (assemble_normal_system!) and the K2.5 scaled build
    (_build_scale_augmented_system_coo! + transfer!) produce the reference's nzval bit for bit."""
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c)
    h.build_kkt()
    o.build_kkt()
    assert np.array_equal(h.vec_get("KKT_NZVAL"), o.aug_nzval)


def test_factor_and_solve_residual(state):
    """(b): K x = b with relative residual <= 1e-10 (north star), inertia as the KKT structure dictates."""
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c)
    h.build_kkt()
    assert h.factorize() == 0 and h.is_factorized().all()
    o.build_kkt()
    N = o.N
    cp = o.aug_colptr.astype(np.int64) - 1
    K = sp.csc_matrix((o.aug_nzval, o.aug_rowval.astype(np.int64) - 1, cp), shape=(N, N))
    K = K + sp.tril(K, -1).T
    rng = np.random.default_rng(1)
    b = rng.standard_normal(N)
    h.vec_fill("W2", 0.0)
    w2 = np.zeros(h.vec_len("W2")); w2[:N] = b
    h.vec_set("W2", w2)
    h.linear_solve_ptr(h.vec_device_ptr("W2"), h.vec_len("W2"))
    x = h.vec_get("W2")[:N]
    res = np.max(np.abs(K @ x - b)) / max(1.0, np.max(np.abs(b)))
    # scale-aware bound: the KKT matrix at an interior iterate is badly scaled; compare with the oracle too
    assert o.linear_solver.factorize()
    xo = o.linear_solver.solve(b.copy())
    res_o = np.max(np.abs(K @ xo - b)) / max(1.0, np.max(np.abs(b)))
    assert res <= max(1e-10, 10 * res_o), (res, res_o)
    pos, zero, neg = h.inertia()
    assert (pos, zero, neg) == o.linear_solver.inertia()
    if kkt != "NORMAL":
        assert (pos, zero, neg) == (o.n, 0, o.m)
    else:
        assert (pos, zero, neg) == (o.m, 0, 0)


def test_kkt_solve_and_mul_match_oracle(state):
    """K4/K4' and K5 on an UnreducedKKTVector."""
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c); h.build_kkt(); assert h.factorize() == 0
    o.build_kkt(); assert o.linear_solver.factorize()
    rng = np.random.default_rng(2)
    v = rng.standard_normal(o.d.size)
    w = rng.standard_normal(o.d.size)
    h.vec_set("W2", v); h.vec_set("D", w)
    h.kkt_mul("D", "W2", -0.5, 2.0)
    ref = o.kkt_mul(w.copy(), v.copy(), -0.5, 2.0)
    np.testing.assert_allclose(h.vec_get("D"), ref, rtol=1e-12, atol=1e-12 * np.max(np.abs(ref)))
    h.vec_set("D", v)
    h.kkt_solve("D")
    got = h.vec_get("D")
    ref = o.kkt_solve(v.copy())
    # both are backward-stable solves of the same ill-conditioned system: compare through the residual
    back_g = o.kkt_mul(np.zeros_like(v), got.copy(), 1.0, 0.0)
    back_o = o.kkt_mul(np.zeros_like(v), ref.copy(), 1.0, 0.0)
    rg = np.max(np.abs(back_g - v)) / max(1.0, np.max(np.abs(v)))
    ro = np.max(np.abs(back_o - v)) / max(1.0, np.max(np.abs(v)))
    assert rg <= max(1e-10, 10 * ro), (rg, ro)
    np.testing.assert_allclose(got, ref, rtol=1e-6, atol=1e-6 * np.max(np.abs(ref)))
    h.vec_set("D", o.d)


def test_solve_system_norms(state):
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c); h.build_kkt(); assert h.factorize() == 0
    o.build_kkt(); assert o.linear_solver.factorize()
    o.set_predictive_rhs()
    h.set_predictive_rhs()
    np.testing.assert_allclose(h.vec_get("P"), o.p, rtol=1e-15, atol=0)
    rc, norms = h.solve_system("D", "P")
    o.solve_system(o.d, o.p)
    assert rc == _lib.OK
    assert norms[0, 1] == np.max(np.abs(o.p))
    np.testing.assert_allclose(norms[0, 2], np.max(np.abs(o.d)), rtol=1e-6)
    assert norms[0, 0] / max(1.0, norms[0, 1]) <= max(1e-10, 10 * o.last_residual_ratio)


def test_rhs_builders(state):
    """K9, /root/reference/src/kernels.jl:1-61."""
    qp, kkt, o, s, h = state
    for dev, orc in ((h.set_initial_primal_rhs, o.set_initial_primal_rhs), (h.set_initial_dual_rhs, o.set_initial_dual_rhs),
                     (h.set_predictive_rhs, o.set_predictive_rhs)):
        dev(); orc()
        np.testing.assert_allclose(h.vec_get("P"), o.p, rtol=1e-15, atol=0)
    mu = 0.37
    h.set_correction_rhs(mu); o.set_correction_rhs(mu)
    np.testing.assert_allclose(h.vec_get("P"), o.p, rtol=1e-12, atol=1e-14)   # FMA contraction only


def test_ratio_test_bit_exact(state):
    """K12, /root/reference/src/kernels.jl:242-305: values and blocking indices identical."""
    qp, kkt, o, s, h = state
    h.set_aug_diagonal_reg(o.del_w, o.del_c); h.build_kkt(); h.factorize()
    o.build_kkt(); o.linear_solver.factorize()
    o.set_predictive_rhs(); o.solve_system(o.d, o.p)
    h.vec_set("D", o.d)
    for tau in (1.0, 0.995, 0.5):
        r = h.ratio_test(tau)[0]
        axl, axu, ixl, ixu = o.get_alpha_max_primal(tau)
        azl, azu, izl, izu = o.get_alpha_max_dual(tau)
        assert r[:4].tolist() == [axl, axu, azl, azu]
        assert r[4:].astype(int).tolist() == [ixl, ixu, izl, izu]


def test_ratio_test_ties_and_empty():
    """Fold semantics of kernels.jl:248 (later index wins ties; index 0 when nothing is below 1)."""
    qp = P.simple_lp()
    s = MPCSolver(qp)
    h = s.h
    h.set_values([], qp.A_vals)
    h.vec_set("X", [1.0, 1.0]); h.vec_set("XL", [0.0, 0.0]); h.vec_set("XU", [np.inf, np.inf])
    h.vec_set("ZL", [2.0, 2.0]); h.vec_set("ZU", [0.0, 0.0])
    d = np.zeros(h.vec_len("D"))
    d[0:2] = [-2.0, -2.0]       # both primal ratios are 0.5: tie -> index 2
    d[3:5] = [1.0, 1.0]         # dual directions positive: nothing blocks -> (1.0, 0)
    h.vec_set("D", d)
    r = h.ratio_test(1.0)[0]
    assert r[0] == 0.5 and int(r[4]) == 2
    assert r[2] == 1.0 and int(r[6]) == 0
    assert r[1] == 1.0 and int(r[5]) == 0 and r[3] == 1.0 and int(r[7]) == 0    # no upper bounds at all
    d[0:2] = [-1.0, -0.5]       # ratio exactly 1.0 at index 1: ties with the init -> index 1 (the element wins)
    h.vec_set("D", d)
    r = h.ratio_test(1.0)[0]
    assert r[0] == 1.0 and int(r[4]) == 1
    s.close()


def test_affine_phase_and_corrections(state):
    """K10 + K11 + K12 fused (solver.jl:295-297) and Gondzio's set_extra_correction! (kernels.jl:78-126)."""
    qp, kkt, o, s, h = state
    h.vec_set("D", o.d)
    aff = h.phase_affine()[0]
    ap, ad = o.get_fraction_to_boundary_step(1.0)
    assert aff[0] == ap and aff[1] == ad
    np.testing.assert_allclose(aff[2], o.get_affine_complementarity_measure(ap, ad), rtol=1e-12)
    np.testing.assert_allclose(aff[3], o.get_complementarity_measure(), rtol=1e-12)
    o.get_correction()
    np.testing.assert_allclose(h.vec_get("CORR_LB"), o.correction_lb, rtol=1e-15)
    np.testing.assert_allclose(h.vec_get("CORR_UB"), o.correction_ub, rtol=1e-15)
    c = h.complementarity(0.3, 0.6)[0]
    np.testing.assert_allclose(c[0], o.get_affine_complementarity_measure(0.3, 0.6), rtol=1e-12)
    mu = 0.05
    h.set_extra_correction(0.4, 0.7, 0.1, 10.0, mu)
    o.set_extra_correction(0.4, 0.7, 0.1, 10.0, mu)
    np.testing.assert_allclose(h.vec_get("CORR_LB"), o.correction_lb, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(h.vec_get("CORR_UB"), o.correction_ub, rtol=1e-13, atol=1e-15)


def test_mehrotra_step(state):
    """MehrotraAdaptiveStep without scalar device reads (kernels.jl:325-374)."""
    qp, kkt, o, s, h = state
    h.vec_set("D", o.d)
    o.opt.step_rule = ("mehrotra", 0.99)
    o.update_step()
    got = h.mehrotra_step(0.99)[0]
    np.testing.assert_allclose(got, [o.alpha_p, o.alpha_d], rtol=1e-10)


def test_update_eval_residuals(state):
    """K14 (+adjust_boundary!), K8, K7+K15."""
    qp, kkt, o, s, h = state
    for name, arr in (("X", o.x), ("XL", o.xl), ("XU", o.xu), ("ZL", o.zl), ("ZU", o.zu), ("Y", o.y), ("D", o.d)):
        h.vec_set(name, arr)
    ap, ad, mu = 0.6, 0.8, 1e-3
    h.update_iterate(ap, ad, mu)
    dx, dy, dzl, dzu = o._views(o.d)
    o.x += ap * dx; o.y += ad * dy; o.zl[o.ind_lb] += ad * dzl; o.zu[o.ind_ub] += ad * dzu
    o.mu = mu
    obj = h.eval_model(float(o.qp.c0 * o.obj_scale))[0]
    o.obj_val = o.eval_f(); o.eval_cons(); o.eval_grad_f(); o.adjust_boundary()
    for name, ref in (("X", o.x), ("Y", o.y), ("ZL", o.zl), ("ZU", o.zu), ("XL", o.xl), ("XU", o.xu)):
        np.testing.assert_allclose(h.vec_get(name), ref, rtol=1e-13, atol=1e-15, err_msg=name)   # FMA contraction only
    np.testing.assert_allclose(obj, o.obj_val, rtol=1e-12)
    np.testing.assert_allclose(h.vec_get("F"), o.f, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(h.vec_get("C"), o.c, rtol=1e-12, atol=1e-12)
    res = h.phase_residuals()[0]
    jacl = o._Amat.T @ o.y
    np.testing.assert_allclose(h.vec_get("JACL"), jacl, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(res[0], np.max(np.abs(o.c)) if o.m else 0.0, rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(res[1], np.max(np.abs(o.f - o.zl + o.zu + jacl)), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(res[2], o.get_optimality_gap(), rtol=1e-12)
    np.testing.assert_allclose(res[3], np.max(np.abs(dx)), rtol=0)


def test_adjust_boundary_moves_a_touched_bound():
    qp = P.simple_lp()
    s = MPCSolver(qp)
    h = s.h
    h.set_values([], qp.A_vals)
    h.vec_set("X", [1e-30, 1.0]); h.vec_set("XL", [0.0, 0.0]); h.vec_set("XU", [np.inf, np.inf])
    h.vec_fill("D", 0.0)
    h.update_iterate(0.0, 0.0, 1e-3)
    xl = h.vec_get("XL")
    eps = np.finfo(float).eps
    assert xl[0] == 0.0 - eps ** 0.75 * 1.0 and xl[1] == 0.0
    s.close()


def test_nan_in_solve_is_reported():
    """linear_solver.jl:41-43: NaN residual -> status NONFINITE at the ABI."""
    qp = P.maros_meszaros_like(seed=1, n=60, m=30, nnz=200)
    o = oracle_at_iterate(qp, "K2", iters=1)
    s, h = device_like(o, qp, "K2")
    h.set_aug_diagonal_reg(o.del_w, o.del_c); h.build_kkt(); assert h.factorize() == 0
    p = o.p.copy(); p[3] = np.nan
    h.vec_set("P", p)
    rc, norms = h.solve_system("D", "P")
    assert rc == _lib.NONFINITE and np.isnan(norms[0, 0])
    s.close()


def test_zero_pivot_is_reported_and_retry_recovers():
    """is_factorized + the x100 retry of linear_solver.jl:6-17."""
    qp = P.simple_lp()
    s = MPCSolver(qp, regularization=FixedRegularization(0.0, 0.0))
    s.initialize()
    h = s.h
    h.vec_fill("ZL", 0.0)                       # Sigma = 0 + delta_w = 0 -> zero pivot in the (1,1) block
    h.set_aug_diagonal_reg(0.0, 0.0)
    h.build_kkt()
    assert h.factorize() == 1 and not h.is_factorized()[0]
    s.del_w, s.del_c = 1e-8, 0.0
    h.vec_fill("ZL", 1.0)
    s.factorize_regularized_system()
    assert h.is_factorized()[0]
    s.close()


def test_full_size_c3_factor_solve_properties():
    """BASELINE config C3 at full size (n=500k, m=250k): assembly + factor + solve through size-independent
    properties: residual of K x = b, linearity of the solve, inertia (n, 0, m)."""
    qp = P.banded_qp()
    s = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8))
    s.initialize()
    h = s.h
    s._update_regularization()
    h.set_aug_diagonal_reg(s.del_w, s.del_c)
    h.build_kkt()
    assert h.factorize() == 0
    assert h.inertia() == (s.n, 0, s.m)
    # (i) the first real predictor solve: the reference's own residual measure (linear_solver.jl:28-36)
    h.set_predictive_rhs()
    rc, nr = h.solve_system("D", "P")
    assert rc == _lib.OK
    ratio = nr[0, 0] / max(1.0, nr[0, 1])
    # same inputs on the CPU oracle (reference's AMD-style ordering through the library's host code)
    import bench
    perm, _ = bench.cpu_perm(qp, "K2")
    o = O.MPCOracle(qp, perm=perm, regularization=REG)
    o.initialize(); o.update_regularization(); o.factorize_regularized_system(); o.set_predictive_rhs()
    o.solve_system(o.d, o.p)
    print("C3 full-size predictor solve: residual ratio gpu", ratio, "oracle", o.last_residual_ratio)
    assert ratio <= max(1e-10, 10 * o.last_residual_ratio)
    # (ii) random right-hand sides: normwise backward error and linearity of the solve
    N = s.n + s.m
    cp, ri = h.symbolic("KKT_COLPTR"), h.symbolic("KKT_ROWVAL")
    nz = h.vec_get("KKT_NZVAL")
    K = sp.csc_matrix((nz, ri, cp), shape=(N, N))
    K = K + sp.tril(K, -1).T
    normK = abs(K).sum(axis=1).max()
    rng = np.random.default_rng(0)
    L = h.vec_len("D")
    sols = []
    for v in (rng.standard_normal(N), rng.standard_normal(N)):
        w = np.zeros(L); w[:N] = v
        h.vec_set("W2", w)
        h.linear_solve_ptr(h.vec_device_ptr("W2"), L)
        x = h.vec_get("W2")[:N]
        back = np.max(np.abs(K @ x - v)) / (normK * np.max(np.abs(x)) + np.max(np.abs(v)))
        print("C3 full-size random rhs: backward error", back)
        assert back <= 1e-14
        sols.append((v, x))
    (v1, x1), (v2, x2) = sols
    w = np.zeros(L); w[:N] = 2.0 * v1 - 3.0 * v2
    h.vec_set("W2", w)
    h.linear_solve_ptr(h.vec_device_ptr("W2"), L)
    x3 = h.vec_get("W2")[:N]
    np.testing.assert_allclose(x3, 2.0 * x1 - 3.0 * x2, rtol=0, atol=1e-6 * np.max(np.abs(x3)))
    s.close()


@pytest.mark.parametrize("maker,kw,kkt", [
    (P.inequality_qp, dict(seed=1, n=60, m=40, n_ineq=15), "K2"),
    (P.maros_meszaros_like, dict(seed=3, n=300, m=150, nnz=1500), "K2"),
    (P.sparse_lp, dict(n=3000, m=1500), "NORMAL"),
    (P.inequality_qp, dict(seed=1, n=60, m=40, n_ineq=15), "K2.5"),
])
@pytest.mark.parametrize("scaling", [True, False])
def test_load_problem_matches_host_preparation(maker, kw, kkt, scaling):
    """madipm_load_problem (initialize! + set_scaling! + value compression on the device) against the numpy
    restatement MPCSolver.prepare_host + upload: same iterate, bounds, rhs, cost, scalings and KKT values."""
    from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
    qp = maker(**kw)
    a = MPCSolver(qp, kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8), scaling=scaling, device_prepare=False)
    b = MPCSolver(qp, kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8), scaling=scaling, device_prepare=True)
    a.prepare_host(); a.upload()
    b.load_device()
    assert np.allclose(b.obj_scale, a.obj_scale, rtol=1e-14, atol=0)
    assert np.allclose(b.norm_b, a.norm_b, rtol=1e-14, atol=0)
    assert np.allclose(b.h.vec_get("CON_SCALE"), a.con_scale.ravel(), rtol=1e-15, atol=0)
    for v in ("X", "XL", "XU", "Y", "RHS", "COST"):
        x, y = a.h.vec_get(v), b.h.vec_get(v)
        assert np.array_equal(np.isinf(x), np.isinf(y)), v
        f = np.isfinite(x)
        assert np.allclose(y[f], x[f], rtol=1e-14, atol=1e-300), v
    if kkt != "NORMAL":
        assert np.allclose(b.h.vec_get("KKT_NZVAL"), a.h.vec_get("KKT_NZVAL"), rtol=1e-14, atol=1e-300)
    sa, sb = a.solve(), b.solve()
    assert sa.status == sb.status and abs(int(sa.iter[0]) - int(sb.iter[0])) <= 1
    assert abs(sa.objective[0] - sb.objective[0]) <= 1e-8 * max(1.0, abs(sa.objective[0]))
    assert np.allclose(sa.multipliers, sb.multipliers, rtol=1e-5, atol=1e-7)
    a.close(); b.close()
