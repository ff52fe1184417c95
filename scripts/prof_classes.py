"""Per-class device time of one solve (madipm_profile level 1: a stream synchronisation per call; development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from madqp_jl_b200.mpc import MPCSolver
wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
qps, kkt, name = bench.make_workload(wl)
s = MPCSolver(qps if len(qps) > 1 else qps[0], kkt_system=kkt, **bench.solver_options())
s.solve()
t = time.time(); st = s.solve(); dt0 = time.time() - t
s.h.profile(1); s.h.profile_read()
t = time.time(); st = s.solve(); dt = time.time() - t
pr = s.h.profile_read()
it = float(st.iter[0])
print(f"{wl}: solve {dt0:.3f} s unprofiled, {dt:.3f} s profiled, {int(it)} iterations, init {st.init_time:.3f} s")
for k, (ms, calls) in pr.items():
    print("   %-16s %9.2f ms  %6d calls  %8.3f ms/call  %7.3f ms/iter" % (k, ms, calls, ms / max(calls, 1), ms / it))
s.close()
