"""Stage clocks of the factor tasks near the root (development aid; needs a library built with -DMADIPM_TREE_TIMING, pointed to by
MADIPM_LIB).  Prints, for the largest front of each of the last heights, the cycles per stage with k_front and with k_tree_top."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import bench
    from madqp_jl_b200.mpc import MPCSolver
    qps, kkt, _ = bench.make_workload("C3")
    s = MPCSolver(qps[0], kkt_system=kkt, **bench.solver_options())
    s.initialize(); s._update_regularization()
    h = s.h
    for _ in range(2):
        h.phase_factor_predictor(s.del_w, s.del_c)
    h.synchronize()
    s.close()
    sys.exit(0)

import bench
from madqp_jl_b200.mpc import MPCSolver
qps, kkt, _ = bench.make_workload("C3")
s = MPCSolver(qps[0], kkt_system=kkt, **bench.solver_options())
lev = s.h.symbolic("SUPER_LEVEL"); rp = s.h.symbolic("SUPER_ROWPTR"); first = s.h.symbolic("SUPER_FIRST")
hh = np.diff(rp); ww = np.diff(first)
pick = []
for l in sorted(set(lev))[-9:-1][::2] + [sorted(set(lev))[7]]:
    idx = np.where(lev == l)[0]
    pick.append(int(idx[np.argmax(hh[idx] * ww[idx])]))
print("supernodes", pick, [(int(lev[p]), int(hh[p]), int(ww[p])) for p in pick], flush=True)
for nofront in ("0", "1"):
    env = dict(os.environ, MADIPM_DBG_SN=",".join(map(str, pick)), MADIPM_NO_FRONT=nofront)
    print(f"--- MADIPM_NO_FRONT={nofront}", flush=True)
    out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
    lines = [l for l in out.stdout.splitlines() if l.startswith(("front", "factor", "forward"))]
    print("\n".join(lines[-3 * len(pick):]), flush=True)
    if out.returncode:
        print(out.stderr[-2000:])
