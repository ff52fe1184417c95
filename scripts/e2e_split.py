import sys,time; sys.path.insert(0,'/root/repo')
import bench
from madqp_jl_b200.mpc import MPCSolver
qps,kkt,_=bench.make_workload("C3",1.0)
s=MPCSolver(qps[0],kkt_system=kkt,**bench.solver_options())
s.solve()
for _ in range(3):
    t0=time.perf_counter(); s.prepare_host(); t1=time.perf_counter(); s.upload(); s.h.synchronize(); t2=time.perf_counter(); s.start(); s.h.synchronize(); t3=time.perf_counter(); s.mpc(); s.h.synchronize(); t4=time.perf_counter()
    print("prepare_host %.1f ms  upload %.1f ms  start %.1f ms  mpc %.1f ms" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3))
