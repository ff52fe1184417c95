"""Quick end-to-end sanity run on a B200 (development aid; the real checks live in tests/)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization, NoRegularization
from oracle import oracle as O

def run(qp, name, ok=True, **kw):
    t = time.time()
    s = MPCSolver(qp, **kw)
    st = s.solve()
    dt = time.time() - t
    okw = dict(kkt_system=kw.get("kkt_system", "K2"))
    r = kw.get("regularization")
    if isinstance(r, NoRegularization): okw["regularization"] = ("none",)
    elif r is not None: okw["regularization"] = ("fixed", r.delta_p, r.delta_d)
    if "max_ncorr" in kw: okw["max_ncorr"] = kw["max_ncorr"]
    ref = O.MPCOracle(qp, perm=s.h.symbolic("PERM"), **okw).solve() if ok else None
    print(f"{name}: gpu {st.status[0]} it={int(st.iter[0])} obj={st.objective[0]:.10e} t={dt:.2f}s "
          f"solve={st.total_time:.3f}s analysis={s.analysis_time:.2f}s launches={s.h.launch_count()}"
          + (f" | oracle {ref.status} it={ref.k} obj={ref.objective:.10e} t={ref.total_time:.2f}s" if ref else ""), flush=True)
    print("   info", {k: s.info[k] for k in ("kkt_order", "nnz_L", "factor_flops", "nsuper", "n_block_tasks", "n_warp_tasks", "n_big", "n_levels", "max_height")}, flush=True)
    if len(s.trace) > 1:
        print("   last", {k: (float(v[0]) if hasattr(v, '__len__') else v) for k, v in s.trace[-1].items()}, "resid", s.last_residual_ratio)
    s.close()
    return st

which = sys.argv[1:] or ["small"]
if "small" in which:
    run(P.simple_lp(), "simple_lp K2", regularization=NoRegularization())
    run(P.simple_lp(), "simple_lp NORMAL", kkt_system="NORMAL")
    run(P.maros_meszaros_like(n=400, m=200, nnz=2000), "mm400", regularization=FixedRegularization(1e-8, -1e-8))
    run(P.maros_meszaros_like(), "C1", regularization=FixedRegularization(1e-8, -1e-8))
    run(P.inequality_qp(), "ineq", regularization=FixedRegularization(1e-8, -1e-8))
    run(P.sparse_lp(n=2000, m=1000), "lp2000 NORMAL", kkt_system="NORMAL", regularization=FixedRegularization(1e-8, -1e-8))
    run(P.banded_qp(n=20000, m=10000), "qp20k", regularization=FixedRegularization(1e-8, -1e-8))
if "c3" in which:
    run(P.banded_qp(), "C3", ok="oracle" in which, regularization=FixedRegularization(1e-8, -1e-8))
if "prof" in which:
    from madqp_jl_b200.mpc import AdaptiveStep
    qp = P.banded_qp()
    s = MPCSolver(qp, regularization=FixedRegularization(1e-8, -1e-8), step_rule=AdaptiveStep(0.995), max_iter=300)
    s.solve()
    s.h.profile(1); s.h.profile_read()
    t = time.time(); st = s.solve(); dt = time.time() - t
    pr = s.h.profile_read()
    print("C3 profiled solve: %.3fs, %d iters" % (dt, int(st.iter[0])))
    for k, (ms, calls) in pr.items():
        print("   %-16s %9.2f ms  %6d calls  %8.3f ms/call  %6.2f ms/iter" % (k, ms, calls, ms / max(calls, 1), ms / st.iter[0]))
    print("   sum classes %.1f ms" % sum(v[0] for k, v in pr.items() if k != "factor_tree"))
    s.close()
if "c2" in which:
    run(P.sparse_lp(), "C2", ok=False, kkt_system="NORMAL", regularization=FixedRegularization(1e-8, -1e-8))
if "c4" in which:
    run(P.transportation_lp(), "C4", ok=False, kkt_system="NORMAL", regularization=FixedRegularization(1e-8, -1e-8))
if "c4h" in which:
    run(P.transportation_lp(n_src=12500, n_dst=12500, n_arcs=1000000), "C4half", ok=False, kkt_system="NORMAL", regularization=FixedRegularization(1e-8, -1e-8))
