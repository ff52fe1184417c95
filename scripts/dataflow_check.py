"""Dataflow scheduling of the tree kernels (k_tree_df: dependency counters instead of height barriers) against the barrier
schedule on the same handle: results must be bit-identical (same tasks, same summation order), times are reported.
usage: python scripts/dataflow_check.py [C3|C5] [scale] [reps]   -> one JSON line per workload"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from madqp_jl_b200.mpc import MPCSolver

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def kernels(s, tag):
    """factorize + one full KKT solve of the predictor system at the starting point: hashes of the factor-dependent results."""
    h = s.h
    s._update_regularization()
    h.set_aug_diagonal_reg(s.del_w, s.del_c); h.build_kkt()
    out = {}
    for _ in range(2):
        assert h.factorize() == 0
        h.set_predictive_rhs()
        h.solve_system("D", "P")
    out["d"] = digest(h.vec_get("D"))
    h.profile(1); h.profile_read()
    for _ in range(reps):
        assert h.factorize() == 0
        h.set_predictive_rhs()
        h.kkt_solve("P")
    pr = h.profile_read()
    h.profile(0)
    out["factor_ms"] = pr["factor"][0] / reps
    out["solve_ms"] = pr["solve"][0] / reps
    out["d2"] = digest(h.vec_get("P"))
    return out


def solves(s):
    ts, st = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        st = s.solve(resident=True)
        ts.append(time.perf_counter() - t0)
    return dict(ms=1e3 * min(ts), iters=int(st.iter.max()), status=sorted(set(st.status)), obj=digest(st.objective),
                x=digest(st.solution), launches=int(s.h.launch_count()))


qps, kkt, name = bench.make_workload(wl, scale)
s = MPCSolver(qps if len(qps) > 1 else qps[0], kkt_system=kkt, **bench.solver_options())
s.keep_snapshot = True
s.solve()                                            # uploads, snapshot, warm-up (barrier schedule)
res = {"workload": name, "instances": len(qps)}
for tag, on in (("barrier", False), ("dataflow", True), ("barrier2", False)):
    eff = s.h.set_tree_dataflow(on)
    if on and not eff:
        res["dataflow"] = "not in effect for this workload"
        continue
    s.upload_resident(); s.start()
    r = kernels(s, tag)
    r.update(solve=solves(s))
    res[tag] = r
b, d = res.get("barrier"), res.get("dataflow")
if isinstance(d, dict):
    res["bit_identical"] = bool(b["d"] == d["d"] and b["d2"] == d["d2"] and b["solve"]["obj"] == d["solve"]["obj"]
                                and b["solve"]["x"] == d["solve"]["x"] and b["solve"]["iters"] == d["solve"]["iters"])
    res["speedup_factor"] = b["factor_ms"] / d["factor_ms"]
    res["speedup_solve"] = b["solve_ms"] / d["solve_ms"]
    res["speedup_whole"] = b["solve"]["ms"] / d["solve"]["ms"]
print(json.dumps(res), flush=True)
s.close()
