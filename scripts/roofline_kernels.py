"""Isolated roofline measurement of the streaming kernels (SpMV, assembly, vector phases) on config C3.

Each kernel is timed with CUDA events on the solver's stream as a back-to-back series of 20 launches (no host
gaps inside the timed region): warm = series / 20 (operands stay in the 126 MB L2), cold = a 512 MB memset
between launches evicts L2 and the same series without the kernel is subtracted.  Bytes are the algorithmic
bytes of SURVEY 8d, the peak is the measured HBM copy bandwidth of MEASURED_PEAKS.json.  Output: one JSON line per kernel (kept under profiles/).
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from madqp_jl_b200 import _lib
from madqp_jl_b200.mpc import MPCSolver, FixedRegularization

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0      # > 1: the same generator at a multiple of the size (launch overhead amortised)
qps, kkt, name = bench.make_workload(wl, scale)
qp = qps[0]
s = MPCSolver(qp, kkt_system=kkt, regularization=FixedRegularization(1e-8, -1e-8))
s.prepare_host(); s.upload(); s.start(); s._update_regularization()
h = s.h
h.set_aug_diagonal_reg(s.del_w, s.del_c); h.build_kkt(); h.factorize(); h.set_predictive_rhs(); h.solve_system("D", "P")
info = s.info
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
stream = torch.cuda.ExternalStream(h.stream())
flush = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device="cuda")
n, m, nlb, nub = s.n, s.m, s.nlb, s.nub
nnzA, nnzQ = qp.nnzj + s.ns, (2 * qp.nnzh - int(np.sum(qp.H_rows == qp.H_cols))) if qp.nnzh else 0
L = n + m + nlb + nub

def spmv_bytes(nnz, rows, cols):
    return 12 * nnz + 4 * (rows + 1) + 8 * cols + 8 * rows

KERNELS = {
    "spmv_A   (k_spmv, y = A x)": (lambda: h.spmv(_lib.OP_A, "C", "X"), spmv_bytes(nnzA, m, n)),
    "spmv_AT  (k_spmv, y = A'x)": (lambda: h.spmv(_lib.OP_AT, "JACL", "Y"), spmv_bytes(nnzA, n, m)),
    "spmv_Q   (k_spmv, y = Q x)": (lambda: h.spmv(_lib.OP_Q, "F", "X"), spmv_bytes(nnzQ, n, n)) if nnzQ else None,
    "assembly K1 (k_set_aug_diagonal_reg)": (lambda: h.set_aug_diagonal_reg(1e-8, -1e-8), 8 * (5 * (nlb + nub) + 2 * n + m) + 8 * (nlb + nub) + 8 * n),
    "assembly K2a (k_build_kkt_k2_perm)": (lambda: h.build_kkt(), 20 * (n + m)) if kkt == "K2" else (lambda: h.build_kkt(), 12 * nnzA + 8 * n + 12 * info["kkt_nnz"] + 12 * info["n_products"]),
    "rhs K9 (k_set_rhs predictive)": (lambda: h.set_predictive_rhs(), 8 * (5 * n + 2 * m + 4 * (nlb + nub)) + 4 * (nlb + nub)),
    "kkt_mul K5 (3 spmv + tail)": (lambda: h.kkt_mul("W1", "D", -1.0, 1.0), spmv_bytes(nnzA, m, n) + spmv_bytes(nnzA, n, m) + (spmv_bytes(nnzQ, n, n) if nnzQ else 0) + 8 * 4 * L),
    "eval_model K8 (k_eval_model)": (lambda: h.eval_model(0.0, want_obj=False), spmv_bytes(nnzA, m, n) + spmv_bytes(nnzQ, n, n) + 8 * (2 * n + m)),
    "update K14 (k_update_iterate)": (lambda: h.update_iterate(0.0, 0.0, 1e-3), 8 * (9 * n + 2 * m) + 8 * (nlb + nub)),
}

def run_series(fn, cold, reps):
    """reps x [flush?; kernel] back to back on the solver's stream between two events (no host gaps)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps):
            if cold:
                flush.zero_()
            if fn is not None:
                fn()
        e1.record(stream)
    h.synchronize()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def timed(fn, cold, reps=20):
    """Device time of one launch: the series with the kernel minus the series without it (cold: a 512 MB
    memset between launches evicts L2, and its own time is what the second series measures)."""
    base = min(run_series(None, cold, reps) for _ in range(3)) if cold else 0.0
    full = min(run_series(fn, cold, reps) for _ in range(3))
    return max(full - base, 1e-9) / reps


def cusparse_baselines():
    """cusparseSpMV (CSR, Int32 indices, FP64, CUSPARSE_SPMV_ALG_DEFAULT) on the same three operators: what the reference's GPU
    path calls for every product (/root/reference/scripts/cuda_wrapper.jl:60-87, which also rebuilds the descriptors per call --
    not timed here).  A stated baseline for k_spmv; library code, never on the product path."""
    import ctypes as C
    import scipy.sparse as sp
    try:
        cs = C.CDLL("libcusparse.so.12")
    except OSError:
        return {}
    hdl = C.c_void_p()
    assert cs.cusparseCreate(C.byref(hdl)) == 0
    assert cs.cusparseSetStream(hdl, C.c_void_p(h.stream())) == 0
    A = sp.csr_matrix((qp.A_vals, (qp.A_rows, qp.A_cols)), shape=(qp.ncon, qp.nvar))
    ops = {"A": A, "AT": A.T.tocsr()}
    if qp.nnzh:
        Hl = sp.coo_matrix((qp.H_vals, (qp.H_rows, qp.H_cols)), shape=(qp.nvar, qp.nvar))
        ops["Q"] = (Hl + sp.tril(Hl, -1).T).tocsr()
    out, keep = {}, []
    one, zero = C.c_double(1.0), C.c_double(0.0)
    for name, M in ops.items():
        M.sort_indices()
        rp = torch.as_tensor(M.indptr.astype(np.int32), device="cuda")
        ci = torch.as_tensor(M.indices.astype(np.int32), device="cuda")
        va = torch.as_tensor(M.data.astype(np.float64), device="cuda")
        x = torch.ones(M.shape[1], dtype=torch.float64, device="cuda")
        y = torch.zeros(M.shape[0], dtype=torch.float64, device="cuda")
        mat, vx, vy = C.c_void_p(), C.c_void_p(), C.c_void_p()
        assert cs.cusparseCreateCsr(C.byref(mat), C.c_int64(M.shape[0]), C.c_int64(M.shape[1]), C.c_int64(M.nnz), C.c_void_p(rp.data_ptr()),
                                    C.c_void_p(ci.data_ptr()), C.c_void_p(va.data_ptr()), 2, 2, 0, 1) == 0      # 32I, 32I, base 0, CUDA_R_64F
        assert cs.cusparseCreateDnVec(C.byref(vx), C.c_int64(M.shape[1]), C.c_void_p(x.data_ptr()), 1) == 0
        assert cs.cusparseCreateDnVec(C.byref(vy), C.c_int64(M.shape[0]), C.c_void_p(y.data_ptr()), 1) == 0
        bs = C.c_size_t(0)
        assert cs.cusparseSpMV_bufferSize(hdl, 0, C.byref(one), mat, vx, C.byref(zero), vy, 1, 0, C.byref(bs)) == 0
        buf = torch.empty(max(bs.value, 8), dtype=torch.uint8, device="cuda")
        keep += [rp, ci, va, x, y, buf]

        def fn(mat=mat, vx=vx, vy=vy, buf=buf):
            rc = cs.cusparseSpMV(hdl, 0, C.byref(one), mat, vx, C.byref(zero), vy, 1, 0, C.c_void_p(buf.data_ptr()))
            assert rc == 0, rc
        # check against scipy once
        fn(); torch.cuda.synchronize()
        assert np.allclose(y.cpu().numpy(), M @ np.ones(M.shape[1]), rtol=1e-12, atol=1e-12)
        out[f"cusparseSpMV {name:2s} (CSR, f64, ALG_DEFAULT; baseline)"] = (fn, spmv_bytes(M.nnz, M.shape[0], M.shape[1]))
    out["_keep"] = keep
    return out


if scale == 1.0 or True:
    base = cusparse_baselines()
    _keep = base.pop("_keep", None)
    KERNELS.update(base)

for kname, spec in KERNELS.items():
    if spec is None:
        continue
    fn, nbytes = spec
    for _ in range(3):
        fn()
    h.synchronize()
    tc, tw = timed(fn, True), timed(fn, False)
    print(json.dumps(dict(workload=wl if scale == 1.0 else f"{wl} x{scale:g}", kernel=kname, algorithmic_MB=nbytes / 1e6, cold_us=tc * 1e6, warm_us=tw * 1e6,
                          cold_GBs=nbytes / tc / 1e9, warm_GBs=nbytes / tw / 1e9, peak_GBs=peak,
                          frac_cold=nbytes / tc / 1e9 / peak, frac_warm=nbytes / tw / 1e9 / peak)), flush=True)
s.close()
