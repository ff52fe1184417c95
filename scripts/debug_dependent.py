import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
import test_gpu_solve as T
for kw in (dict(), dict(regularization=FixedRegularization(1e-8, -1e-8))):
    reg = kw.get("regularization", FixedRegularization(1e-8, 0.0))
    s = MPCSolver(T.dependent_rows_qp(), regularization=reg, max_iter=40, fused_iteration=False)
    st = s.solve()
    print("reg", reg, "->", st.status[0], int(st.iter[0]), st.objective[0], "del", s.del_w, s.del_c)
    for tr in s.trace[:40:3]:
        print("  k", tr["k"], "inf_pr %.2e inf_du %.2e compl %.2e mu %.2e ap %.2e ad %.2e" % (tr["inf_pr"][0], tr["inf_du"][0], tr["inf_compl"][0], tr["mu"][0], tr["alpha_p"][0], tr["alpha_d"][0]))
    print("  last residual ratio", s.last_residual_ratio)
    s.close()
base = MPCSolver(T.dependent_rows_qp(dup=False), regularization=FixedRegularization(1e-8, 0.0))
st = base.solve(); print("no dup:", st.status[0], int(st.iter[0]), st.objective[0])
