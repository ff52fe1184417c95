"""Times factorize / kkt_solve alone on a workload for several schedule parameters (development aid)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200 import problems as P
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
budgets = [int(float(x)) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
extra = {}
for a in sys.argv[4:]:
    k, v = a.split("=")
    extra[k] = tuple(float(x) for x in v.split(",")) if "," in v else int(float(v))
qps, kkt, name = bench.make_workload(wl, scale)
for b in budgets:
    s = MPCSolver(qps if len(qps) > 1 else qps[0], kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8), block_panel=b, **extra)
    s.prepare_host(); s.upload(); s.start()
    s._update_regularization()
    h = s.h
    h.set_aug_diagonal_reg(s.del_w, s.del_c); h.build_kkt()
    for _ in range(3):
        h.factorize(); h.set_predictive_rhs(); h.solve_system("D", "P")
    h.profile(1); h.profile_read()
    for _ in range(10):
        assert h.factorize() == 0
        h.set_predictive_rhs()
        h.kkt_solve("P")
    pr = h.profile_read()
    i = s.info
    print(f"{wl} block_panel={b} {extra}: factor {pr['factor'][0]/10:.3f} ms  solve {pr['solve'][0]/10:.3f} ms | nsuper={i['nsuper']} block/warp/big tasks={i['n_block_tasks']}/{i['n_warp_tasks']}/{i['n_big']} "
          f"heights={i['n_levels']} nnzL={i['nnz_L']} "
          f"flops={i['factor_flops']:.3g} GF/s={i['factor_flops']*len(qps)/(pr['factor'][0]/10)/1e6:.1f}", flush=True)
    s.close()
