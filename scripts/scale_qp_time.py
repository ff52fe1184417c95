"""Times madipm_scale_qp (device Ruiz equilibration + rescaling, host buffers in and out) on the C2-sized matrix against
the numpy oracle of the same algorithm.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200 import preprocess, problems  # noqa: E402
from oracle import oracle as O  # noqa: E402

qp = problems.sparse_lp(seed=0, n=200_000, m=100_000)
preprocess.scale_qp(qp)                      # warm-up (context, allocations)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    s, Dr, Dc, info = preprocess.scale_qp(qp)
    ts.append(time.perf_counter() - t0)
t0 = time.perf_counter()
r, oDr, oDc, _, _ = O.scale_qp(qp)
t_cpu = time.perf_counter() - t0
print(json.dumps(dict(what="scale_qp, C2 matrix (1M nnz, 10 sweeps), host buffers in/out", device_ms=1e3 * min(ts),
                      oracle_numpy_ms=1e3 * t_cpu, launches=info["launches"], bit_exact=bool(np.array_equal(Dr, oDr) and np.array_equal(s.A_vals, r.A_vals)))))
