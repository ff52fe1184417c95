"""Dataflow schedule (k_tree_df) against the barrier schedule over many tree shapes: whole solves must agree bit for bit.
One line per case on stdout (flushed), a summary line at the end.  A hang surfaces as MADIPM_INTERNAL after
MADIPM_SYNC_TIMEOUT_S (set it small).  usage: python scripts/dataflow_stress.py [seconds]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import FixedRegularization, MPCSolver

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 35.0
t_start = time.perf_counter()
cases = []
for k, n in enumerate([300, 1000, 2000, 5000, 12000, 20000, 40000, 100000, 250000]):
    cases.append((f"banded n={n}", lambda n=n, k=k: P.banded_qp(seed=k, n=n, m=n // 2), dict()))
cases.append(("banded n=30000 ncorr3", lambda: P.banded_qp(seed=11, n=30000, m=15000), dict(max_ncorr=3)))
cases.append(("banded n=30000 K2.5", lambda: P.banded_qp(seed=12, n=30000, m=15000), dict(kkt_system="K2.5")))
for k, n in enumerate([60, 200, 600, 1500]):
    cases.append((f"mm n={n}", lambda n=n, k=k: P.maros_meszaros_like(seed=k, n=n, m=n // 2, nnz=4 * n), dict()))
    cases.append((f"mm n={n} K2.5", lambda n=n, k=k: P.maros_meszaros_like(seed=k, n=n, m=n // 2, nnz=4 * n), dict(kkt_system="K2.5")))
for k in range(3):
    cases.append((f"ineq seed={k}", lambda k=k: P.inequality_qp(seed=k), dict()))
cases.append(("simple_lp", P.simple_lp, dict()))
cases.append(("sparse_lp n=3000", lambda: P.sparse_lp(seed=4, n=3000, m=1500), dict()))
for B in (1, 7, 64, 300, 1500):
    cases.append((f"batch B={B}", lambda B=B: [P.small_qp(instance=i) for i in range(B)], dict()))

ok = bad = skipped = 0
for name, make, opts in cases:
    if time.perf_counter() - t_start > budget:
        print(f"budget reached before {name}", flush=True)
        break
    try:
        qp = make()
        outs, times, eff = [], [], None
        for on in (False, True):
            s = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8), **opts)
            e = s.h.set_tree_dataflow(on)
            if on:
                eff = e
            t0 = time.perf_counter()
            st = s.solve()
            if on and eff:                      # a few more solves on the same handle (counters are re-zeroed per launch)
                for _ in range(2):
                    st2 = s.solve()
                    assert np.array_equal(st2.solution, st.solution)
            times.append(time.perf_counter() - t0)
            info = s.info
            outs.append(st)
            s.close()
        a, b = outs
        if not eff:
            skipped += 1
            print(f"SKIP {name}: dataflow not in effect (n_big={info['n_big']})", flush=True)
            continue
        same = (a.status == b.status and np.array_equal(a.iter, b.iter) and np.array_equal(a.objective, b.objective)
                and np.array_equal(a.solution, b.solution) and np.array_equal(a.multipliers, b.multipliers))
        ok += same
        bad += not same
        print(f"{'OK  ' if same else 'DIFF'} {name}: {sorted(set(b.status))} it={int(b.iter.max())} heights={info['n_levels']} "
              f"nsuper={info['nsuper']} barrier {1e3 * times[0]:.1f} ms dataflow(x3) {1e3 * times[1]:.1f} ms", flush=True)
    except Exception as ex:  # noqa: BLE001
        bad += 1
        print(f"FAIL {name}: {type(ex).__name__}: {ex}", flush=True)
print(f"summary ok={ok} bad={bad} skipped={skipped} in {time.perf_counter() - t_start:.1f} s", flush=True)
