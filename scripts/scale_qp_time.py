"""Times madipm_scale_qp (device Ruiz equilibration + rescaling, host buffers in and out) on the C2-sized matrix and
checks the equilibration property on the result.  (The bit-exact comparison with the oracle lives in
tests/test_scale_qp.py; the numpy oracle of the same algorithm took 97 ms on the round-1 box.)  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madqp_jl_b200 import preprocess, problems  # noqa: E402

qp = problems.sparse_lp(seed=0, n=200_000, m=100_000)
preprocess.scale_qp(qp)                      # warm-up (context, allocations)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    s, Dr, Dc, info = preprocess.scale_qp(qp)
    ts.append(time.perf_counter() - t0)
rmax = np.zeros(qp.ncon); cmax = np.zeros(qp.nvar)
np.maximum.at(rmax, s.A_rows, np.abs(s.A_vals)); np.maximum.at(cmax, s.A_cols, np.abs(s.A_vals))
print(json.dumps(dict(what="scale_qp, C2 matrix (1M nnz, 10 sweeps), host buffers in/out", device_ms=1e3 * min(ts),
                      launches=info["launches"], max_row_norm_dev=float(np.abs(1 - rmax[rmax > 0]).max()),
                      max_col_norm_dev=float(np.abs(1 - cmax[cmax > 0]).max()))))
