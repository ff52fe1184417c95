"""ctypes binding of libmadipm_b200.so (declarations: include/madipm_b200.h).

This is the Python twin of the Julia ``ccall`` shim shown in INTEGRATION.md: the
same entry points, the same argument meaning.  There is no fallback: if the
shared library is missing the import fails loudly, and numeric calls fail with
``MadIPMError`` when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MADIPM_LIB: another build of the same library (instrumented development builds), never a different backend
LIB_PATH = os.environ.get("MADIPM_LIB") or os.path.join(_HERE, "libmadipm_b200.so")

OK, FACTOR_BREAKDOWN, NONFINITE, CUDA_ERROR, BAD_ARGUMENT, INTERNAL = range(6)
KKT_K2, KKT_NORMAL, KKT_K2_SCALED = 0, 1, 2
ORDER_AMD, ORDER_NATURAL, ORDER_ND = 0, 1, 2
OP_A, OP_AT, OP_Q = 0, 1, 2

VEC_NAMES = [
    "X", "XL", "XU", "ZL", "ZU", "F", "JACL", "COST", "Y", "C", "RHS", "D", "P", "W1", "W2",
    "CORR_LB", "CORR_UB", "REG", "PR_DIAG", "DU_DIAG", "L_DIAG", "L_LOWER", "U_DIAG", "U_LOWER", "KKT_NZVAL", "CON_SCALE",
    "SCALING",
]
VEC = {name: i for i, name in enumerate(VEC_NAMES)}

SYM = dict(PERM=0, ETREE_PARENT=1, COLCOUNT=2, SUPER_FIRST=3, SUPER_PARENT=4, SUPER_ROWPTR=5, SUPER_ROWS=6,
           SUPER_XOFF=7, ASM_MAP=8, KKT_COLPTR=9, KKT_ROWVAL=10, AUG_CSC_MAP=11, A_ROWPTR=12, A_COLIND=13,
           A_CSR_MAP=14, SUPER_LEVEL=15, PROD_PTR=16, PROD_A=17, PROD_B=18, PROD_K=19, CHILD_PTR=20, CHILD_IDX=21,
           SUPER_KIND=22, PAR_POS=23, SUPER_COFF=24, TASK_SN=25, PHASE_PTR=26, PHASE_NBLOCK=27, PHASE_CLS=28)

T_CLASSES = ["assembly", "factor", "solve", "spmv", "vector", "factor_tree", "factor_top"]


class MadIPMError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"madipm_b200 status {code}: {message}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("kkt_kind", C.c_int32), ("ordering", C.c_int32), ("max_width", C.c_int32), ("relax_small", C.c_int32),
                ("relax_zeros", C.c_double * 3), ("block_panel", C.c_int64), ("device", C.c_int32), ("batch", C.c_int32),
                ("nd_leaf", C.c_int32), ("reserved", C.c_int32)]


class LoadOptions(C.Structure):
    _fields_ = [("tol", C.c_double), ("bound_push", C.c_double), ("bound_fac", C.c_double), ("max_gradient", C.c_double),
                ("scaling", C.c_int32), ("reserved", C.c_int32)]


class IterOptions(C.Structure):
    _fields_ = [("del_w", C.c_double), ("del_c", C.c_double), ("mu_min", C.c_double), ("tau", C.c_double), ("c0", C.c_double),
                ("res_tol", C.c_double), ("adaptive_step", C.c_int32), ("has_inequalities", C.c_int32)]


class ScaleOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("max_iter", C.c_int32), ("tol", C.c_double), ("launches", C.c_int64)]


class Info(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in "kkt_kind nvar n_slack n_tot ncon nlb nub batch kkt_order".split()]
                + [(n, C.c_int64) for n in "nnzh nnzj kkt_nnz aug_coo_len nnz_L nnz_stored".split()]
                + [("factor_flops", C.c_double)]
                + [(n, C.c_int32) for n in "nsuper max_height max_width n_warp_tasks n_levels".split()]
                + [("n_products", C.c_int64), ("kkt_vec_len", C.c_int64), ("n_block_tasks", C.c_int32),
                   ("n_big", C.c_int32)])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_P32, _P64, _PD, _H = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_void_p

# name -> argument types after the handle (None = no handle argument)
_SIGNATURES = {
    "madipm_default_config": (None, [C.POINTER(Config)], None),
    "madipm_create": (None, [C.POINTER(Config), C.POINTER(_H)], C.c_int32),
    "madipm_destroy": ([], C.c_int32),
    "madipm_last_error": ([], C.c_char_p),
    "madipm_version": (None, [], C.c_char_p),
    "madipm_set_pattern_k2": ([C.c_int32, C.c_int32, C.c_int64, _P32, _P32, C.c_int64, _P32, _P32, C.c_int32, _P64], C.c_int32),
    "madipm_set_pattern_normal": ([C.c_int32, C.c_int32, C.c_int64, _P32, _P32, C.c_int32, _P64], C.c_int32),
    "madipm_set_bounds_index": ([C.c_int32, _P64, C.c_int32, _P64], C.c_int32),
    "madipm_analyze": ([], C.c_int32),
    "madipm_get_info": ([C.POINTER(Info)], C.c_int32),
    "madipm_get_symbolic": ([C.c_int32, _P64, C.c_int64, _P64], C.c_int32),
    "madipm_set_values": ([_PD, _PD], C.c_int32),
    "madipm_load_problem": ([_PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD, C.POINTER(LoadOptions), _PD, _PD], C.c_int32),
    "madipm_vec_len": ([C.c_int32, _P64], C.c_int32),
    "madipm_vec_set": ([C.c_int32, _PD], C.c_int32),
    "madipm_vec_get": ([C.c_int32, _PD], C.c_int32),
    "madipm_vec_fill": ([C.c_int32, C.c_double], C.c_int32),
    "madipm_vec_device_ptr": ([C.c_int32, C.POINTER(C.c_void_p)], C.c_int32),
    "madipm_synchronize": ([], C.c_int32),
    "madipm_stream": ([], C.c_void_p),
    "madipm_kkt_initialize": ([], C.c_int32),
    "madipm_set_aug_diagonal_reg": ([C.c_double, C.c_double], C.c_int32),
    "madipm_set_initial_regularization": ([C.c_double, C.c_double], C.c_int32),
    "madipm_build_kkt": ([], C.c_int32),
    "madipm_factorize": ([_P32], C.c_int32),
    "madipm_is_factorized": ([_P32], C.c_int32),
    "madipm_set_pivot_checks": ([C.c_double, C.c_double], C.c_int32),
    "madipm_inertia": ([C.c_int32, _P64, _P64, _P64], C.c_int32),
    "madipm_linear_solve": ([C.c_void_p, C.c_int64], C.c_int32),
    "madipm_kkt_solve": ([C.c_int32], C.c_int32),
    "madipm_kkt_mul": ([C.c_int32, C.c_int32, C.c_double, C.c_double], C.c_int32),
    "madipm_solve_system": ([C.c_int32, C.c_int32, _PD], C.c_int32),
    "madipm_spmv": ([C.c_int32, C.c_int32, C.c_int32], C.c_int32),
    "madipm_spmv_ptr": ([C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64], C.c_int32),
    "madipm_eval_model": ([C.c_double, _PD], C.c_int32),
    "madipm_phase_residuals": ([_PD], C.c_int32),
    "madipm_set_initial_primal_rhs": ([], C.c_int32),
    "madipm_set_initial_dual_rhs": ([], C.c_int32),
    "madipm_set_predictive_rhs": ([], C.c_int32),
    "madipm_set_correction_rhs": ([_PD], C.c_int32),
    "madipm_ratio_test": ([_PD, _PD], C.c_int32),
    "madipm_complementarity": ([_PD, _PD, _PD], C.c_int32),
    "madipm_phase_affine": ([_PD], C.c_int32),
    "madipm_get_correction": ([], C.c_int32),
    "madipm_set_extra_correction": ([_PD, _PD, C.c_double, C.c_double, _PD], C.c_int32),
    "madipm_update_iterate": ([_PD, _PD, _PD, _P32], C.c_int32),
    "madipm_mehrotra_step": ([C.c_double, _PD], C.c_int32),
    "madipm_init_multipliers_and_shift": ([C.c_double], C.c_int32),
    "madipm_vec_copy": ([C.c_int32, C.c_int32], C.c_int32),
    "madipm_vec_axpy": ([C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64, C.c_double], C.c_int32),
    "madipm_vec_norm_inf": ([C.c_int32, _PD], C.c_int32),
    "madipm_vec_get_range": ([C.c_int32, C.c_int64, C.c_int64, _PD], C.c_int32),
    "madipm_vec_set_range": ([C.c_int32, C.c_int64, C.c_int64, _PD], C.c_int32),
    "madipm_vec_fill_range": ([C.c_int32, C.c_int64, C.c_int64, C.c_double], C.c_int32),
    "madipm_vec_copy_range": ([C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int64], C.c_int32),
    "madipm_vec_norm_inf_range": ([C.c_int32, C.c_int64, C.c_int64, _PD], C.c_int32),
    "madipm_vec_axpy_at": ([C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_double], C.c_int32),
    "madipm_adjust_boundary": ([_PD], C.c_int32),
    "madipm_pinned_alloc": (None, [C.c_int64, C.POINTER(C.c_void_p)], C.c_int32),
    "madipm_pinned_free": (None, [C.c_void_p], C.c_int32),
    "madipm_snapshot_inputs": ([], C.c_int32),
    "madipm_restore_inputs": ([], C.c_int32),
    "madipm_launch_count": ([], C.c_int64),
    "madipm_profile": ([C.c_int32], C.c_int32),
    "madipm_profile_read": ([_PD, _P64, C.c_int32], C.c_int32),
    "madipm_debug_barrier_state": ([_P64], C.c_int32),
    "madipm_debug_inject_barrier_fault": ([], C.c_int32),
    "madipm_debug_describe": ([], C.c_char_p),
    "madipm_phase_residuals_obj": ([C.c_double, _PD, _PD], C.c_int32),
    "madipm_phase_predictor": ([_PD, _PD], C.c_int32),
    "madipm_phase_factor_predictor": ([C.c_double, C.c_double, _PD, _PD], C.c_int32),
    "madipm_phase_corrector": ([_PD, _PD, _PD, _PD], C.c_int32),
    "madipm_phase_iteration": ([C.POINTER(IterOptions), _P32, _PD, _P32], C.c_int32),
    "madipm_set_tree_dataflow": ([C.c_int32, _P32], C.c_int32),
    "madipm_default_scale_options": (None, [C.POINTER(ScaleOptions)], None),
    "madipm_scale_qp": (None, [C.POINTER(ScaleOptions), C.c_int32, C.c_int32, C.c_int64, _P32, _P32, _PD, C.c_int64, _P32, _P32, _PD,
                               _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _P32, _PD], C.c_int32),
    "madipm_scale_last_error": (None, [], C.c_char_p),
    "madipm_standard_form_qp": (None, [C.c_int32, C.c_int32, C.c_int32, C.c_int64, _P32, _P32, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD,
                                       _P32, _P32, _P32, _P32, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _PD, _P64], C.c_int32),
}

EXPORTED_SYMBOLS = sorted(_SIGNATURES)

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, sig in _SIGNATURES.items():
        fn = getattr(lib, name)
        if sig[0] is None:
            fn.argtypes, fn.restype = sig[1], sig[2]
        else:
            fn.argtypes, fn.restype = [_H] + sig[0], sig[1]
    _lib = lib
    return lib


class PinnedArray:
    """A float64 numpy array over page-locked host memory (madipm_pinned_alloc).  `.a` is the array; freed with the object.
    Falls back to ordinary memory when no CUDA device is present (host-only use of the package)."""

    def __init__(self, shape):
        lib = load()
        n = int(np.prod(shape))
        self._ptr = C.c_void_p()
        self.pinned = lib.madipm_pinned_alloc(8 * n, C.byref(self._ptr)) == OK and bool(self._ptr.value)
        if self.pinned:
            buf = (C.c_double * max(n, 1)).from_address(self._ptr.value)
            self.a = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)
        else:
            self.a = np.zeros(shape)

    def __del__(self):
        try:
            if getattr(self, "pinned", False) and self._ptr.value:
                load().madipm_pinned_free(self._ptr)
                self._ptr = C.c_void_p()
        except Exception:  # noqa: BLE001
            pass


def _pd(a):
    return None if a is None else a.ctypes.data_as(_PD)


def _p32(a):
    return None if a is None else a.ctypes.data_as(_P32)


def _p64(a):
    return None if a is None else a.ctypes.data_as(_P64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Handle:
    """Thin object wrapper: one method per C entry point, numpy in / numpy out."""

    def __init__(self, kkt_kind=KKT_K2, batch=1, device=0, **cfg):
        self.lib = load()
        c = Config()
        self.lib.madipm_default_config(C.byref(c))
        c.kkt_kind, c.batch, c.device = kkt_kind, batch, device
        for k, v in cfg.items():
            if k == "relax_zeros":
                for i in range(3):
                    c.relax_zeros[i] = v[i]
            elif not hasattr(c, k):
                raise TypeError(f"unknown config field {k}")
            else:
                setattr(c, k, v)
        self.cfg = c
        self.batch = batch
        self._h = _H()
        rc = self.lib.madipm_create(C.byref(c), C.byref(self._h))
        if rc != OK:
            raise MadIPMError(rc, "madipm_create failed (bad configuration)")

    # -- plumbing
    def _check(self, rc, allow=()):
        if rc != OK and rc not in allow:
            raise MadIPMError(rc, self.lib.madipm_last_error(self._h).decode())
        return rc

    def close(self):
        if self._h:
            self.lib.madipm_destroy(self._h)
            self._h = _H()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- symbolic
    def set_pattern_k2(self, nvar, ncon, hess_I, hess_J, jac_I, jac_J, ind_ineq):
        """1-based Int32 triplets, 1-based Int64 ind_ineq (as Julia holds them)."""
        hI, hJ = np.ascontiguousarray(hess_I, np.int32), np.ascontiguousarray(hess_J, np.int32)
        jI, jJ = np.ascontiguousarray(jac_I, np.int32), np.ascontiguousarray(jac_J, np.int32)
        iq = np.ascontiguousarray(ind_ineq, np.int64)
        self._check(self.lib.madipm_set_pattern_k2(self._h, nvar, ncon, hI.size, _p32(hI), _p32(hJ), jI.size, _p32(jI),
                                                   _p32(jJ), iq.size, _p64(iq)))

    def set_pattern_normal(self, nvar, ncon, jac_I, jac_J, ind_ineq):
        jI, jJ = np.ascontiguousarray(jac_I, np.int32), np.ascontiguousarray(jac_J, np.int32)
        iq = np.ascontiguousarray(ind_ineq, np.int64)
        self._check(self.lib.madipm_set_pattern_normal(self._h, nvar, ncon, jI.size, _p32(jI), _p32(jJ), iq.size, _p64(iq)))

    def set_bounds_index(self, ind_lb, ind_ub):
        lb, ub = np.ascontiguousarray(ind_lb, np.int64), np.ascontiguousarray(ind_ub, np.int64)
        self._check(self.lib.madipm_set_bounds_index(self._h, lb.size, _p64(lb), ub.size, _p64(ub)))

    def analyze(self):
        self._check(self.lib.madipm_analyze(self._h))
        return self.info()

    def info(self):
        i = Info()
        self._check(self.lib.madipm_get_info(self._h, C.byref(i)))
        return i.as_dict()

    def symbolic(self, which):
        which = SYM[which] if isinstance(which, str) else which
        n = C.c_int64()
        self._check(self.lib.madipm_get_symbolic(self._h, which, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int64)
        self._check(self.lib.madipm_get_symbolic(self._h, which, _p64(out), n.value, C.byref(n)))
        return out

    # -- values / vectors
    def set_values(self, hess_V, jac_V):
        hv, jv = _f64(hess_V), _f64(jac_V)
        self._check(self.lib.madipm_set_values(self._h, _pd(hv), _pd(jv)))

    def load_problem(self, x0, lvar, uvar, y0, lcon, ucon, c, hess_V, jac_V, tol, bound_push, bound_fac, scaling,
                     max_gradient=100.0):
        """initialize! + set_scaling! + value compression on the device; returns (obj_scale, norm_b) per instance."""
        arrs = [_f64(a) for a in (x0, lvar, uvar, y0, lcon, ucon, c, hess_V, jac_V)]
        opt = LoadOptions(tol, bound_push, bound_fac, max_gradient, 1 if scaling else 0, 0)
        os_, nb = np.zeros(self.batch), np.zeros(self.batch)
        self._check(self.lib.madipm_load_problem(self._h, *[_pd(a) for a in arrs], C.byref(opt), _pd(os_), _pd(nb)))
        return os_, nb, sum(a.nbytes for a in arrs)

    def vec_len(self, vec):
        n = C.c_int64()
        self._check(self.lib.madipm_vec_len(self._h, VEC[vec], C.byref(n)))
        return n.value

    def vec_set(self, vec, host):
        a = _f64(host)
        assert a.size == self.batch * self.vec_len(vec), (vec, a.size)
        self._check(self.lib.madipm_vec_set(self._h, VEC[vec], _pd(a)))

    def vec_get(self, vec, out=None):
        """out: a preallocated (ideally page-locked, PinnedArray) array of the right size to read into."""
        if out is None:
            out = np.zeros(self.batch * self.vec_len(vec))
        else:
            assert out.size == self.batch * self.vec_len(vec) and out.flags.c_contiguous and out.dtype == np.float64
        self._check(self.lib.madipm_vec_get(self._h, VEC[vec], _pd(out)))
        out = out.reshape(-1)
        return out.reshape(self.batch, -1) if self.batch > 1 else out

    def vec_fill(self, vec, value):
        self._check(self.lib.madipm_vec_fill(self._h, VEC[vec], value))

    def vec_device_ptr(self, vec):
        p = C.c_void_p()
        self._check(self.lib.madipm_vec_device_ptr(self._h, VEC[vec], C.byref(p)))
        return p.value

    def vec_copy(self, dst, src):
        self._check(self.lib.madipm_vec_copy(self._h, VEC[dst], VEC[src]))

    def vec_axpy(self, dst, off_dst, src, off_src, length, alpha):
        self._check(self.lib.madipm_vec_axpy(self._h, VEC[dst], off_dst, VEC[src], off_src, length, alpha))

    # -- sub-range / index-view operations (what julia/MadIPMB200.jl's vector type calls)
    def vec_get_range(self, vec, off, length):
        out = np.zeros(self.batch * length)
        self._check(self.lib.madipm_vec_get_range(self._h, VEC[vec], off, length, _pd(out)))
        return out.reshape(self.batch, -1) if self.batch > 1 else out

    def vec_set_range(self, vec, off, host):
        a = _f64(host)
        self._check(self.lib.madipm_vec_set_range(self._h, VEC[vec], off, a.size // self.batch, _pd(a)))

    def vec_fill_range(self, vec, off, length, value):
        self._check(self.lib.madipm_vec_fill_range(self._h, VEC[vec], off, length, value))

    def vec_copy_range(self, dst, off_dst, src, off_src, length):
        self._check(self.lib.madipm_vec_copy_range(self._h, VEC[dst], off_dst, VEC[src], off_src, length))

    def vec_norm_inf_range(self, vec, off, length):
        out = np.zeros(self.batch)
        self._check(self.lib.madipm_vec_norm_inf_range(self._h, VEC[vec], off, length, _pd(out)))
        return out

    def vec_axpy_at(self, dst, which, src, off_src, alpha):
        self._check(self.lib.madipm_vec_axpy_at(self._h, VEC[dst], which, VEC[src], off_src, alpha))

    def adjust_boundary(self, mu):
        m = _f64(np.broadcast_to(mu, (self.batch,)))
        self._check(self.lib.madipm_adjust_boundary(self._h, _pd(m)))

    def vec_norm_inf(self, vec):
        out = np.zeros(self.batch)
        self._check(self.lib.madipm_vec_norm_inf(self._h, VEC[vec], _pd(out)))
        return out

    def synchronize(self):
        self._check(self.lib.madipm_synchronize(self._h))

    def stream(self):
        return self.lib.madipm_stream(self._h)

    # -- KKT system / linear solver
    def kkt_initialize(self):
        self._check(self.lib.madipm_kkt_initialize(self._h))

    def set_aug_diagonal_reg(self, del_w, del_c):
        self._check(self.lib.madipm_set_aug_diagonal_reg(self._h, del_w, del_c))

    def set_initial_regularization(self, del_w, del_c):
        self._check(self.lib.madipm_set_initial_regularization(self._h, del_w, del_c))

    def build_kkt(self):
        self._check(self.lib.madipm_build_kkt(self._h))

    def factorize(self):
        """Returns the number of instances whose factorisation broke down."""
        n = C.c_int32()
        self._check(self.lib.madipm_factorize(self._h, C.byref(n)), allow=(FACTOR_BREAKDOWN,))
        return n.value

    def factorize_async(self):
        """madipm_factorize(h, NULL): no synchronisation; read the outcome with is_factorized() after the next call that
        returns scalars."""
        self._check(self.lib.madipm_factorize(self._h, None))

    def set_pivot_checks(self, tiny, growth):
        self._check(self.lib.madipm_set_pivot_checks(self._h, float(tiny), float(growth)))

    def set_tree_dataflow(self, on=True):
        """Dependency counters instead of height barriers in the small-front tree kernels; returns whether it is in effect."""
        eff = C.c_int32(0)
        self._check(self.lib.madipm_set_tree_dataflow(self._h, int(bool(on)), C.byref(eff)))
        return bool(eff.value)

    def is_factorized(self):
        f = np.zeros(self.batch, dtype=np.int32)
        self._check(self.lib.madipm_is_factorized(self._h, _p32(f)))
        return f.astype(bool)

    def inertia(self, inst=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.madipm_inertia(self._h, inst, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def linear_solve_ptr(self, dev_ptr, stride):
        self._check(self.lib.madipm_linear_solve(self._h, dev_ptr, stride))

    def kkt_solve(self, vec):
        self._check(self.lib.madipm_kkt_solve(self._h, VEC[vec]))

    def kkt_mul(self, w, v, alpha=1.0, beta=0.0):
        self._check(self.lib.madipm_kkt_mul(self._h, VEC[w], VEC[v], alpha, beta))

    def solve_system(self, d="D", p="P"):
        norms = np.zeros(3 * self.batch)
        rc = self._check(self.lib.madipm_solve_system(self._h, VEC[d], VEC[p], _pd(norms)), allow=(NONFINITE,))
        return rc, norms.reshape(self.batch, 3)

    def spmv(self, op, out, inp):
        self._check(self.lib.madipm_spmv(self._h, op, VEC[out], VEC[inp]))

    # -- fused phases
    def eval_model(self, c0=0.0, want_obj=True):
        obj = np.zeros(self.batch) if want_obj else None
        self._check(self.lib.madipm_eval_model(self._h, c0, _pd(obj)))
        return obj

    def phase_residuals(self):
        out = np.zeros(4 * self.batch)
        self._check(self.lib.madipm_phase_residuals(self._h, _pd(out)))
        return out.reshape(self.batch, 4)

    def phase_residuals_obj(self, c0=0.0):
        out, obj = np.zeros(4 * self.batch), np.zeros(self.batch)
        self._check(self.lib.madipm_phase_residuals_obj(self._h, c0, _pd(out), _pd(obj)))
        return out.reshape(self.batch, 4), obj

    def phase_predictor(self):
        norms, out = np.zeros(3 * self.batch), np.zeros(4 * self.batch)
        rc = self._check(self.lib.madipm_phase_predictor(self._h, _pd(norms), _pd(out)), allow=(NONFINITE,))
        return rc, norms.reshape(self.batch, 3), out.reshape(self.batch, 4)

    def phase_factor_predictor(self, del_w, del_c):
        """set_aug_diagonal_reg! + build_kkt! + factorize! (asynchronous flags) + phase_predictor, the predictor's forward
        sweep inside the factor launch."""
        norms, out = np.zeros(3 * self.batch), np.zeros(4 * self.batch)
        rc = self._check(self.lib.madipm_phase_factor_predictor(self._h, del_w, del_c, _pd(norms), _pd(out)), allow=(NONFINITE,))
        return rc, norms.reshape(self.batch, 3), out.reshape(self.batch, 4)

    def phase_iteration(self, del_w, del_c, mu_min, tau, adaptive, has_inequalities, c0=0.0, active=None, res_tol=0.0):
        """One mpc! iteration with one synchronisation (madipm_phase_iteration).  Returns (ok, out[batch, 28])."""
        opt = IterOptions(del_w, del_c, mu_min, tau, c0, res_tol, 1 if adaptive else 0, 1 if has_inequalities else 0)
        out = np.zeros(28 * self.batch)
        ok = C.c_int32(0)
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        self._check(self.lib.madipm_phase_iteration(self._h, C.byref(opt), _p32(act), _pd(out), C.byref(ok)))
        return bool(ok.value), out.reshape(self.batch, 28)

    def phase_corrector(self, mu, tau):
        mu = _f64(np.broadcast_to(mu, (self.batch,)))
        tau = _f64(np.broadcast_to(tau, (self.batch,)))
        norms, ratio = np.zeros(3 * self.batch), np.zeros(8 * self.batch)
        rc = self._check(self.lib.madipm_phase_corrector(self._h, _pd(mu), _pd(tau), _pd(norms), _pd(ratio)), allow=(NONFINITE,))
        return rc, norms.reshape(self.batch, 3), ratio.reshape(self.batch, 8)

    def set_initial_primal_rhs(self):
        self._check(self.lib.madipm_set_initial_primal_rhs(self._h))

    def set_initial_dual_rhs(self):
        self._check(self.lib.madipm_set_initial_dual_rhs(self._h))

    def set_predictive_rhs(self):
        self._check(self.lib.madipm_set_predictive_rhs(self._h))

    def set_correction_rhs(self, mu):
        mu = _f64(np.broadcast_to(mu, (self.batch,)))
        self._check(self.lib.madipm_set_correction_rhs(self._h, _pd(mu)))

    def ratio_test(self, tau):
        tau = _f64(np.broadcast_to(tau, (self.batch,)))
        out = np.zeros(8 * self.batch)
        self._check(self.lib.madipm_ratio_test(self._h, _pd(tau), _pd(out)))
        return out.reshape(self.batch, 8)

    def complementarity(self, alpha_p=None, alpha_d=None):
        ap = None if alpha_p is None else _f64(np.broadcast_to(alpha_p, (self.batch,)))
        ad = None if alpha_d is None else _f64(np.broadcast_to(alpha_d, (self.batch,)))
        out = np.zeros(2 * self.batch)
        self._check(self.lib.madipm_complementarity(self._h, _pd(ap), _pd(ad), _pd(out)))
        return out.reshape(self.batch, 2)

    def phase_affine(self):
        out = np.zeros(4 * self.batch)
        self._check(self.lib.madipm_phase_affine(self._h, _pd(out)))
        return out.reshape(self.batch, 4)

    def get_correction(self):
        self._check(self.lib.madipm_get_correction(self._h))

    def set_extra_correction(self, alpha_p, alpha_d, beta_min, beta_max, mu):
        ap, ad = _f64(np.broadcast_to(alpha_p, (self.batch,))), _f64(np.broadcast_to(alpha_d, (self.batch,)))
        mu = _f64(np.broadcast_to(mu, (self.batch,)))
        self._check(self.lib.madipm_set_extra_correction(self._h, _pd(ap), _pd(ad), beta_min, beta_max, _pd(mu)))

    def update_iterate(self, alpha_p, alpha_d, mu, active=None):
        ap, ad = _f64(np.broadcast_to(alpha_p, (self.batch,))), _f64(np.broadcast_to(alpha_d, (self.batch,)))
        mu = _f64(np.broadcast_to(mu, (self.batch,)))
        act = None if active is None else np.ascontiguousarray(active, dtype=np.int32)
        self._check(self.lib.madipm_update_iterate(self._h, _pd(ap), _pd(ad), _pd(mu), _p32(act)))

    def mehrotra_step(self, gamma_f):
        out = np.zeros(2 * self.batch)
        self._check(self.lib.madipm_mehrotra_step(self._h, gamma_f, _pd(out)))
        return out.reshape(self.batch, 2)

    def init_multipliers_and_shift(self, bound_fac):
        self._check(self.lib.madipm_init_multipliers_and_shift(self._h, bound_fac))

    def snapshot_inputs(self):
        self._check(self.lib.madipm_snapshot_inputs(self._h))

    def restore_inputs(self):
        self._check(self.lib.madipm_restore_inputs(self._h))

    # -- measurement
    def launch_count(self):
        return int(self.lib.madipm_launch_count(self._h))

    def last_error(self):
        return self.lib.madipm_last_error(self._h).decode()

    def debug_barrier_state(self):
        """What the first grid barrier that gave up saw (host-visible memory: readable while a kernel is stuck)."""
        out = np.zeros(8, dtype=np.int64)
        self._check(self.lib.madipm_debug_barrier_state(self._h, _p64(out)))
        return dict(zip("aborted kernel op height block grid seen target".split(), (int(v) for v in out)))

    def debug_describe(self):
        """Host-only snapshot (no CUDA call): safe from a watchdog thread while the owning thread is blocked."""
        return self.lib.madipm_debug_describe(self._h).decode()

    def debug_inject_barrier_fault(self):
        self._check(self.lib.madipm_debug_inject_barrier_fault(self._h))

    def profile(self, level=1):
        """1: every class with a sync per call (diagnostics); 2: factorisation only, no extra sync; 0: off."""
        self._check(self.lib.madipm_profile(self._h, int(level)))

    def profile_read(self, reset=True):
        ms = np.zeros(len(T_CLASSES))
        calls = np.zeros(len(T_CLASSES), dtype=np.int64)
        self._check(self.lib.madipm_profile_read(self._h, _pd(ms), _p64(calls), int(reset)))
        return {k: (float(ms[i]), int(calls[i])) for i, k in enumerate(T_CLASSES)}
