"""The persistent tree kernels synchronise their grid by hand; these tests make sure that path can neither drift nor hang
silently (the driver's benchmark sat at 100 % GPU for 30 minutes without output).

  * a deliberately broken barrier counter surfaces as MADIPM_INTERNAL with the barrier's state in last_error, within the
    device-side spin budget, and the handle keeps working afterwards;
  * repeated full-size C3 solves (the headline workload) on one handle converge identically every time.
"""
import os
import time

import numpy as np
import pytest

from madqp_jl_b200 import _lib
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import FixedRegularization, MPCSolver

pytestmark = pytest.mark.gpu


def _factor_ready(qp):
    s = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8))
    s.initialize()
    s._update_regularization()
    s.h.set_aug_diagonal_reg(s.del_w, s.del_c)
    s.h.build_kkt()
    return s


@pytest.mark.parametrize("where", ["factor", "solve"])
def test_broken_grid_barrier_is_an_error_not_a_hang(where, monkeypatch):
    monkeypatch.setenv("MADIPM_BARRIER_TIMEOUT_MS", "300")
    s = _factor_ready(P.banded_qp(seed=1, n=40000, m=20000))     # several heights -> the launches have barriers
    h = s.h
    assert h.info()["n_levels"] > 3
    h.set_tree_dataflow(False)                                    # the barrier schedule (the dataflow one has its own test below)
    assert h.factorize() == 0
    good = h.inertia()
    assert h.debug_barrier_state()["aborted"] == 0
    t0 = time.perf_counter()
    h.debug_inject_barrier_fault()
    if where == "factor":
        with pytest.raises(_lib.MadIPMError) as ei:
            h.factorize()
    else:
        h.set_predictive_rhs()
        with pytest.raises(_lib.MadIPMError) as ei:
            h.kkt_solve("P")
            h.synchronize()
    dt = time.perf_counter() - t0
    assert ei.value.code == _lib.INTERNAL
    assert "grid barrier timed out" in str(ei.value)
    st = h.debug_barrier_state()
    assert st["aborted"] == 1 and st["seen"] >= 0x80000000 and st["grid"] > 0
    assert dt < 30.0, dt                                          # a handful of barriers x 0.3 s, not for ever
    # the handle is usable again: same factorisation as before the fault
    h.build_kkt()
    assert h.factorize() == 0
    assert h.inertia() == good
    assert h.debug_barrier_state()["aborted"] == 1
    s.close()


def test_three_full_size_c3_solves_back_to_back():
    """The driver's benchmark repeats solve() on one handle; every repetition must terminate and agree."""
    s = MPCSolver(P.banded_qp(), regularization=FixedRegularization(1e-8, -1e-8), max_ncorr=0, tol=1e-8)
    s.keep_snapshot = True
    out = []
    for k in range(3):
        st = s.solve(resident=(k > 0))
        assert st.status[0] == "SOLVE_SUCCEEDED"
        out.append((int(st.iter[0]), float(st.objective[0])))
    assert s.h.debug_barrier_state()["aborted"] == 0
    assert len({o[0] for o in out}) == 1, out
    assert max(abs(o[1] - out[0][1]) for o in out) <= 1e-10 * max(1.0, abs(out[0][1])), out
    s.close()


def test_bench_supervisor_retries_a_hung_attempt(tmp_path):
    """bench.py at N = 1 runs the benchmark in a child process; an attempt that never returns is killed and repeated once,
    and the line says so."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BENCH_TEST_HANG_ATTEMPT="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--scale", "0.02", "--steps", "1", "--warmup", "1", "--no-cpu",
                          "--no-c5", "--guard-s", "8"], env=env, capture_output=True, text=True, timeout=300)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout + out.stderr[-2000:]
    line = json.loads(lines[0])
    assert "error" not in line and line["attempts"] == 2 and len(line["failed_attempts"]) == 1, line
    assert line["value"] > 0 and line["converged"]


# ---- dataflow scheduling (k_tree_df): dependency counters instead of height barriers ---------------------------------

@pytest.mark.parametrize("where", ["factor", "solve"])
def test_broken_dependency_counter_is_an_error_not_a_hang(where, monkeypatch):
    monkeypatch.setenv("MADIPM_BARRIER_TIMEOUT_MS", "300")
    s = _factor_ready(P.banded_qp(seed=1, n=40000, m=20000))
    h = s.h
    assert h.set_tree_dataflow(True)
    assert h.factorize() == 0
    good = h.inertia()
    t0 = time.perf_counter()
    h.debug_inject_barrier_fault()                                # every counter starts negative: no task with children can start
    if where == "factor":
        with pytest.raises(_lib.MadIPMError) as ei:
            h.factorize()
    else:
        h.set_predictive_rhs()
        with pytest.raises(_lib.MadIPMError) as ei:
            h.kkt_solve("P")
            h.synchronize()
    dt = time.perf_counter() - t0
    assert ei.value.code == _lib.INTERNAL
    assert "dependency wait timed out" in str(ei.value)
    assert dt < 30.0, dt
    h.build_kkt()                                                 # the handle is usable again
    assert h.factorize() == 0
    assert h.inertia() == good
    s.close()


@pytest.mark.parametrize("maker,kw,opts", [
    (P.banded_qp, dict(seed=1, n=40000, m=20000), dict()),
    (P.banded_qp, dict(seed=2, n=20000, m=10000), dict(max_ncorr=3)),
    (P.maros_meszaros_like, dict(seed=0, n=2000, m=1000, nnz=10000), dict(kkt_system="K2.5")),
    (P.sparse_lp, dict(seed=4, n=6000, m=3000), dict()),
])
def test_dataflow_schedule_is_bit_identical_to_the_barrier_schedule(maker, kw, opts):
    """Same tasks, same summation order: whole solves agree bit for bit wherever the dataflow launches are in effect."""
    qp = maker(**kw)
    ref = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8), **opts)
    ref.h.set_tree_dataflow(False)
    a = ref.solve()
    ref.close()
    s = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8), **opts)
    if not s.h.set_tree_dataflow(True):
        s.close()
        pytest.skip("multi-block fronts: the dataflow launches are not in effect")
    b = s.solve()
    s.close()
    assert a.status == b.status and np.array_equal(a.iter, b.iter)
    assert np.array_equal(a.objective, b.objective) and np.array_equal(a.solution, b.solution)
    assert np.array_equal(a.multipliers, b.multipliers)


def test_dataflow_schedule_batch_is_bit_identical():
    qps = [P.small_qp(instance=i) for i in range(64)]
    outs = []
    for on in (False, True):
        s = MPCSolver(qps, regularization=FixedRegularization(1e-8, -1e-8))
        assert s.h.set_tree_dataflow(on) == on
        outs.append(s.solve())
        s.close()
    a, b = outs
    assert a.status == b.status and np.array_equal(a.iter, b.iter)
    assert np.array_equal(a.objective, b.objective) and np.array_equal(a.solution, b.solution)
